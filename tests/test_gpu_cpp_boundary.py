"""The C++ host layer above the C ABI, executed on the GPU: the reference-signature template shim
(cpp/int8conv/conv2DForward3x3.hpp) driven by the check.cu analogue, the libbenchmark sweep app with the reference's
config.json schema, and the benchmark/int8.cu harness.  These are the files a maintainer of the reference would bind
(INTEGRATION.md); each binary is built by __graft_entry__.build() (make -C lowbitdnn-project_b200/cpp all check)."""
import json
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
CPP = os.path.join(ROOT, "lowbitdnn-project_b200", "cpp")


def _run(args, timeout=600):
    r = subprocess.run(args, capture_output=True, text=True, timeout=timeout, cwd=CPP)
    assert r.returncode == 0, f"{args} exited {r.returncode}\n{r.stdout[-2000:]}\n{r.stderr[-2000:]}"
    return r.stdout


def test_check_driver_exact_against_library_conv():
    """checkForward3x3 (cpp/int8conv/check.cu:62-155): 16x128x130x130 (*) 128x128x3x3, {0,1}-valued draws, the new
    operator behind the reference's template signature must equal the fp32 library convolution exactly."""
    exe = os.path.join(CPP, "build", "check")
    if not os.path.exists(exe):
        pytest.skip("cpp/build/check not built (needs the torch C++ headers)")
    out = _run([exe, "4", "5"])
    assert "CHECK OK (4 exact comparisons)" in out, out[-1500:]
    assert "MISMATCH" not in out


def test_int8_bench_harness_runs_config1_and_a_resnet_layer():
    """benchmark/int8.cu (the file the reference left empty): per-layer table + network line."""
    exe = os.path.join(CPP, "build", "int8_bench")
    assert os.path.exists(exe), "cpp/build/int8_bench missing: run __graft_entry__.build()"
    out = _run([exe, "--network", "single_3x3", "--repeats", "3"])
    assert "conv3x3_56_64" in out and "images/s" in out, out[-1500:]
    out = _run([exe, "--network", "resnet50", "--batch", "32", "--layer", "l3.1.conv2", "--repeats", "3"])
    assert "l3.1.conv2" in out and "igemm_tc" in out, out[-1500:]


def test_benchmark_app_sweep_writes_reference_schema(tmp_path):
    """cpp/apps/benchmark.cpp:109-168: config.json sweep -> output.json {repeats, configs, benchmarks:[...]}."""
    exe = os.path.join(CPP, "build", "benchmark_app")
    assert os.path.exists(exe), "cpp/build/benchmark_app missing: run __graft_entry__.build()"
    out_json = str(tmp_path / "output.json")
    _run([exe, os.path.join(CPP, "apps", "config.json"), out_json, "--limit", "6"])
    with open(out_json) as fh:
        res = json.load(fh)
    assert {"repeats", "configs", "benchmarks"} <= set(res)
    assert len(res["benchmarks"]) == 6
    for b in res["benchmarks"]:
        assert {"B", "C", "H", "W", "filters", "filter_width", "filter_height", "config", "name", "timing"} <= set(b)
        assert float(b["timing"]) > 0, b
