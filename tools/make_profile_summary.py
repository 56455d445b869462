"""make_profile_summary.py — turn the artefacts a gpurun call brings back into the tracked files under profiles/.

    python tools/make_profile_summary.py --tag r01 --launches gpurun_out/r01_bench_launches.csv \
        --layers gpurun_out/layers_r50_v21.json --bench gpurun_out/bench_v21.json \
        [--ncu gpurun_out/r01_igemm_final.ncu-rep --ncu-names conv1,l1.1.conv2,...]

Writes profiles/<tag>_bench_launches.csv.gz (raw ncu launch list), profiles/<tag>_bench_launch_table.md,
profiles/<tag>_layers_<network>.json, profiles/<tag>_bench.json, profiles/<tag>_traffic.json (DRAM bytes per launch of the
dominant kernel, read by bench.py for roofline.traffic) and, with --ncu, profiles/<tag>_ncu_<name>.md.
"""
import argparse
import collections
import csv
import gzip
import json
import os
import shutil
import subprocess

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
PROF = os.path.join(ROOT, "profiles")

NCU_METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_subpipe_imma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor_subpipe_imma.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
    "sm__cycles_elapsed.max", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    # shared-memory wavefronts of the tensor pipe (operand reads of tcgen05.mma) and of LSU (epilogue staging), TMA traffic
    "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum", "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum.per_second",
    "l1tex__m_l1tex2xbar_write_bytes_mem_global_op_tma_st.sum.per_second", "sm__memory_throughput.avg.pct_of_peak_sustained_elapsed",
]


def launches(path, layer_names, per_step):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10 and r[0].isdigit()]
    L = collections.OrderedDict()
    for r in rows:
        d = L.setdefault(int(r[0]), {"name": r[4]})
        d[r[12]] = float(r[14])
    ids = sorted(L)
    # the last COMPLETE step of the capture: a step starts with the stem's space-to-depth launch
    starts = [k for k, i in enumerate(ids) if "stem_xform" in L[i]["name"] and k + per_step <= len(ids)]
    first = starts[-1] if starts else len(ids) - per_step
    step = [L[i] for i in ids[first:first + per_step]]
    tot = sum(s["gpu__time_duration.sum"] for s in step)
    lines = ["| # | layer | kernel | ncu duration us (cold cache, serialised) | share of step | DRAM read MB | DRAM write MB |",
             "|---|---|---|---|---|---|---|"]
    dom = []
    for i, (nm, s) in enumerate(zip(layer_names, step)):
        kern = s["name"].split("(")[0].split("::")[-1]
        t, rd, wr = s["gpu__time_duration.sum"], s["dram__bytes_read.sum"], s["dram__bytes_write.sum"]
        lines.append(f"| {i} | {nm} | {kern} | {t / 1e3:.1f} | {100 * t / tot:.1f}% | {rd / 1e6:.1f} | {wr / 1e6:.1f} |")
        if "igemm" in kern:
            dom.append((t, rd + wr))
    rd_t = sum(s["dram__bytes_read.sum"] for s in step)
    wr_t = sum(s["dram__bytes_write.sum"] for s in step)
    lines.append("")
    lines.append(f"sum of launch durations: {tot / 1e6:.3f} ms per step; DRAM read {rd_t / 1e9:.2f} GB, write {wr_t / 1e9:.2f} GB per step; "
                 f"igemm_i8_kernel share of the step: {100 * sum(t for t, _ in dom) / tot:.1f}% over {len(dom)} launches")
    traffic = {"kernel": "igemm_i8_kernel", "launches_per_step": len(dom), "dram_bytes_per_launch": sum(b for _, b in dom) / max(1, len(dom)),
               "dram_bytes_per_step_all_kernels": rd_t + wr_t, "share_of_step_ncu": sum(t for t, _ in dom) / tot,
               "source": os.path.basename(path)}
    return "\n".join(lines), traffic


def ncu_tables(rep, names, tag):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {m: hdr.index(m) for m in NCU_METRICS if m in hdr}
    kcol = hdr.index("Kernel Name")
    def short(v):
        try:
            f = float(v)
            return f"{f:.0f}" if abs(f) >= 1000 else f"{f:.3g}"
        except ValueError:
            return v
    cols = [names[i] if i < len(names) else f"launch{i}" for i in range(len(data))]
    out = [f"# {tag}: ncu --set full --clock-control none, one launch per layer (source: {os.path.basename(rep)})", "",
           "Clocks are not locked (the tool's default lock is off as the profiling recipe asks): under ncu the SMs ran at ~1.77 GHz.", "",
           "| metric | unit | " + " | ".join(cols) + " |", "|---|---|" + "---|" * len(cols),
           "| kernel |  | " + " | ".join(r[kcol].split("(")[0].split("::")[-1][:40] for r in data) + " |"]
    for m, j in idx.items():
        out.append(f"| {m} | {units[j]} | " + " | ".join(short(r[j]) for r in data) + " |")
    return "\n".join(out)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tag", default="r01")
    ap.add_argument("--launches")
    ap.add_argument("--layers", nargs="*", default=[])
    ap.add_argument("--bench")
    ap.add_argument("--ncu")
    ap.add_argument("--ncu-names", default="")
    ap.add_argument("--ncu-out", default="ncu_igemm")
    a = ap.parse_args()
    os.makedirs(PROF, exist_ok=True)
    layer_names = None
    for path in a.layers:
        rep = json.load(open(path))
        shutil.copy(path, os.path.join(PROF, f"{a.tag}_layers_{rep['network']}.json"))
        if rep["network"] == "resnet50":
            layer_names = []
            for l in rep["layers"]:
                if l["kernel"] == "stem_tc":
                    layer_names.append(l["layer"] + " (space-to-depth)")
                layer_names.append(l["layer"])
    if a.bench:
        shutil.copy(a.bench, os.path.join(PROF, f"{a.tag}_bench.json"))
    if a.launches and layer_names:
        table, traffic = launches(a.launches, layer_names, len(layer_names))
        with open(os.path.join(PROF, f"{a.tag}_bench_launch_table.md"), "w") as fh:
            fh.write(f"# {a.tag}: every launch of one bench.py step (ncu --metrics gpu__time_duration.sum,dram__bytes_*.sum --clock-control none)\n\n")
            fh.write("Per-launch times are cold-cache and serialised by the profiler: compare SHARES with bench.py's event times, not absolutes.\n\n")
            fh.write(table + "\n")
        with open(os.path.join(PROF, f"{a.tag}_traffic.json"), "w") as fh:
            json.dump(traffic, fh, indent=1)
        with open(a.launches, "rb") as src, gzip.open(os.path.join(PROF, f"{a.tag}_bench_launches.csv.gz"), "wb") as dst:
            shutil.copyfileobj(src, dst)
    if a.ncu:
        with open(os.path.join(PROF, f"{a.tag}_{a.ncu_out}.md"), "w") as fh:
            fh.write(ncu_tables(a.ncu, [n for n in a.ncu_names.split(",") if n], a.tag) + "\n")


if __name__ == "__main__":
    main()
