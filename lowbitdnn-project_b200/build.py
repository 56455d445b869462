"""
build.py — compiles liblowbit-cnn (the C-ABI shared library) for sm_100a with nvcc, in tree.

    python lowbitdnn-project_b200/build.py [--force] [--verbose]

Output: lowbitdnn-project_b200/lib/liblowbit_cnn.so (git-ignored; travels to the GPU box with the snapshot).
nvcc cross-compiles without a GPU.  cudart is linked statically and the driver API (cuTensorMapEncode*) is
resolved at run time through cudaGetDriverEntryPoint, so the library loads on a CPU-only box too.
"""
from __future__ import annotations

import concurrent.futures as cf
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
OBJDIR = os.path.join(HERE, "build")
LIB = os.path.join(LIBDIR, "liblowbit_cnn.so")
LIB_TRACE = os.path.join(LIBDIR, "liblowbit_cnn_trace.so")     # same sources, igemm_tc.cu with -DLBC_TRACE=1 (tools/trace_layer.py)
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
HOSTCXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"

SOURCES = ["api.cu", "depthwise.cu", "direct_conv.cu", "igemm_tc.cu", "layout.cu", "pool_add.cu", "probes.cu"]
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-ccbin", HOSTCXX, "-Xcompiler", "-fPIC,-Wall,-Wno-unused-function", "--expt-relaxed-constexpr",
]


def _stamp() -> str:
    h = hashlib.sha256()
    for root, _, files in sorted(os.walk(CSRC)):
        for f in sorted(files):
            with open(os.path.join(root, f), "rb") as fh:
                h.update(f.encode())
                h.update(fh.read())
    with open(os.path.join(HERE, "..", "include", "lowbit_cnn.h"), "rb") as fh:
        h.update(fh.read())
    h.update(" ".join(FLAGS).encode())
    return h.hexdigest()


def _compile(src: str, verbose: bool, defines: tuple = (), suffix: str = "") -> str:
    obj = os.path.join(OBJDIR, src.replace(".cu", suffix + ".o"))
    cmd = [NVCC, *FLAGS, *defines, "-c", os.path.join(CSRC, src), "-o", obj]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
    if verbose:
        sys.stderr.write(r.stderr)
    return obj


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(LIBDIR, exist_ok=True)
    os.makedirs(OBJDIR, exist_ok=True)
    stamp_file = os.path.join(OBJDIR, "stamp")
    stamp = _stamp()
    if not force and os.path.exists(LIB) and os.path.exists(LIB_TRACE) and os.path.exists(stamp_file) and open(stamp_file).read() == stamp:
        return LIB
    with cf.ThreadPoolExecutor(max_workers=len(SOURCES) + 1) as ex:
        trace_obj = ex.submit(_compile, "igemm_tc.cu", False, ("-DLBC_TRACE=1",), "_trace")
        objs = list(ex.map(lambda s: _compile(s, verbose), SOURCES))
        trace_obj = trace_obj.result()
    for out, these in ((LIB, objs), (LIB_TRACE, [trace_obj if o.endswith("igemm_tc.o") else o for o in objs])):
        cmd = [NVCC, "-shared", "-ccbin", HOSTCXX, "-gencode", "arch=compute_100a,code=sm_100a", "-o", out, *these]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    with open(stamp_file, "w") as fh:
        fh.write(stamp)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
