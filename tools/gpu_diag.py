"""
gpu_diag.py — crash-isolated bring-up diagnostics for the GPU box.

Each case runs in its own subprocess with a timeout so that a device fault or a watchdog trip in one
kernel configuration cannot mask the others.  Results (mismatch structure included) go to
gpurun_out/diag.json.  This is a development tool: it uses the oracle as checker, like tests/.

    python tools/gpu_diag.py [--only igemm] [--out gpurun_out/diag.json]
"""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)

DIRECT, IGEMM, DW = 1, 2, 3


def D(**kw):
    base = dict(stride_h=1, stride_w=1, pad_h=0, pad_w=0, dil_h=1, dil_w=1, groups=1, relu=0, out_mode=1)
    base.update(kw)
    return base


CASES = [
    ("direct_3x3", D(n=2, h=9, w=7, c=8, k=12, r=3, s=3, pad_h=1, pad_w=1, relu=1, out_mode=0), DIRECT),
    ("direct_stem", D(n=2, h=20, w=20, c=3, k=16, r=7, s=7, stride_h=2, stride_w=2, pad_h=3, pad_w=3, out_mode=0), DIRECT),
    ("dw_24", D(n=2, h=9, w=9, c=24, k=24, r=3, s=3, pad_h=1, pad_w=1, groups=24, relu=1, out_mode=0), DW),
    # --- igemm, pure GEMM via tiled TMA
    ("gemm_c128_k128_i32", D(n=1, h=16, w=16, c=128, k=128, r=1, s=1), IGEMM),
    ("gemm_c128_k128_i8", D(n=1, h=16, w=16, c=128, k=128, r=1, s=1, out_mode=0, relu=1), IGEMM),
    ("gemm_c64_k256", D(n=2, h=14, w=14, c=64, k=256, r=1, s=1), IGEMM),
    ("gemm_c32_k64", D(n=1, h=9, w=11, c=32, k=64, r=1, s=1), IGEMM),
    ("gemm_c16_k16", D(n=3, h=5, w=5, c=16, k=16, r=1, s=1), IGEMM),
    ("gemm_c512_k2048", D(n=1, h=7, w=7, c=512, k=2048, r=1, s=1), IGEMM),
    ("gemm_c144_k48", D(n=1, h=12, w=12, c=144, k=48, r=1, s=1), IGEMM),
    ("gemm_c256_k320", D(n=1, h=8, w=8, c=256, k=320, r=1, s=1), IGEMM),
    ("gemm_big", D(n=8, h=56, w=56, c=256, k=64, r=1, s=1, out_mode=0, relu=1), IGEMM),
    # --- igemm, im2col TMA
    ("i2c_1x1_forced", D(n=1, h=16, w=16, c=128, k=128, r=1, s=1), IGEMM, {"force_im2col": 1}),
    ("i2c_3x3_valid", D(n=1, h=10, w=10, c=64, k=64, r=3, s=3), IGEMM),
    ("i2c_3x3_p1_c64", D(n=1, h=56, w=56, c=64, k=64, r=3, s=3, pad_h=1, pad_w=1), IGEMM),
    ("i2c_3x3_p1_c256", D(n=2, h=14, w=14, c=256, k=256, r=3, s=3, pad_h=1, pad_w=1), IGEMM),
    ("i2c_3x3_s2", D(n=2, h=28, w=28, c=128, k=128, r=3, s=3, stride_h=2, stride_w=2, pad_h=1, pad_w=1), IGEMM),
    ("i2c_1x1_s2", D(n=2, h=28, w=28, c=256, k=512, r=1, s=1, stride_h=2, stride_w=2), IGEMM),
    ("i2c_7x7_s2", D(n=1, h=20, w=20, c=16, k=32, r=7, s=7, stride_h=2, stride_w=2, pad_h=3, pad_w=3), IGEMM),
    ("i2c_dil2", D(n=1, h=15, w=15, c=64, k=64, r=3, s=3, pad_h=2, pad_w=2, dil_h=2, dil_w=2), IGEMM),
    ("i2c_rect", D(n=1, h=17, w=13, c=32, k=32, r=3, s=3, pad_h=1, pad_w=1), IGEMM),
    ("i2c_small_tensor", D(n=1, h=6, w=6, c=64, k=64, r=3, s=3, pad_h=1, pad_w=1), IGEMM),
    ("i2c_resnet_l3", D(n=8, h=14, w=14, c=256, k=256, r=3, s=3, pad_h=1, pad_w=1, out_mode=0, relu=1), IGEMM),
    # --- same stride-1 shapes with the window path disabled (pure im2col coverage)
    ("nowin_3x3_p1_c64", D(n=1, h=56, w=56, c=64, k=64, r=3, s=3, pad_h=1, pad_w=1), IGEMM, {"window": 0}),
    ("nowin_3x3_c256_i8", D(n=2, h=14, w=14, c=256, k=256, r=3, s=3, pad_h=1, pad_w=1, out_mode=0, relu=1), IGEMM, {"window": 0}),
    # --- window path specifics
    ("win_56_c64_i8", D(n=3, h=56, w=56, c=64, k=64, r=3, s=3, pad_h=1, pad_w=1, out_mode=0, relu=1), IGEMM),
    ("win_28_c128_i8", D(n=3, h=28, w=28, c=128, k=128, r=3, s=3, pad_h=1, pad_w=1, out_mode=0, relu=1), IGEMM),
    ("win_14_c256_i32", D(n=3, h=14, w=14, c=256, k=256, r=3, s=3, pad_h=1, pad_w=1), IGEMM),
    ("win_7_c512_i8", D(n=5, h=7, w=7, c=512, k=512, r=3, s=3, pad_h=1, pad_w=1, out_mode=0), IGEMM),
    ("win_224_c64_coltiles", D(n=1, h=20, w=224, c=64, k=64, r=3, s=3, pad_h=1, pad_w=1, out_mode=0, relu=1), IGEMM),
    ("win_150_c32_ragged", D(n=2, h=9, w=150, c=32, k=48, r=3, s=3, pad_h=1, pad_w=1, out_mode=0), IGEMM),
    ("win_5x5_p2", D(n=2, h=19, w=23, c=32, k=32, r=5, s=5, pad_h=2, pad_w=2, out_mode=0, relu=1), IGEMM),
    ("win_k512_ntiles", D(n=2, h=14, w=14, c=128, k=512, r=3, s=3, pad_h=1, pad_w=1, out_mode=0, relu=1), IGEMM),
    ("win_c16_3x3", D(n=2, h=30, w=40, c=16, k=32, r=3, s=3, pad_h=1, pad_w=1, out_mode=0, relu=1), IGEMM),
    ("win_c16_4x4", D(n=2, h=33, w=115, c=16, k=64, r=4, s=4, out_mode=0, relu=1), IGEMM),
    ("win_c16_2x2_i32", D(n=2, h=20, w=57, c=16, k=32, r=2, s=2), IGEMM),
    ("win_1xS", D(n=1, h=8, w=40, c=64, k=64, r=1, s=3, pad_h=0, pad_w=1, out_mode=0), IGEMM),
]


def run_case(label, desc, force, options=None):
    import numpy as np
    from oracle.oracle import ConvDesc as OD
    from tests.parity_util import check_case
    import lowbitdnn_project_b200 as lbc
    od = OD(**desc)
    t0 = time.time()
    try:
        plan = lbc.ConvPlan(lbc.ConvDesc(**desc), force=force, options=options)
        descr = plan.describe()
        plan.close()
    except Exception as e:  # noqa: BLE001
        return {"label": label, "status": "plan_error", "error": str(e)}
    try:
        nbad, total, name, detail = check_case(od, layer=1, force=force, options=options)
    except Exception as e:  # noqa: BLE001
        return {"label": label, "status": "run_error", "error": str(e)[:600], "plan": descr}
    return {"label": label, "status": "ok" if nbad == 0 else "mismatch", "bad": nbad, "total": total, "kernel": name,
            "plan": descr, "detail": detail, "sec": round(time.time() - t0, 2)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--case", default=None)
    ap.add_argument("--only", default=None)
    ap.add_argument("--probes", action="store_true")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "diag.json"))
    a = ap.parse_args()

    if a.case is not None:
        c = CASES[int(a.case)]
        print("DIAG_RESULT " + json.dumps(run_case(c[0], c[1], c[2], c[3] if len(c) > 3 else None)), flush=True)
        return
    if a.probes:
        import lowbitdnn_project_b200 as lbc
        out = {}
        try:
            out["hbm_copy_gbs"] = lbc.probe_hbm_copy(1 << 30, 10)
        except Exception as e:  # noqa: BLE001
            out["hbm_copy_error"] = str(e)
        try:
            out["int8_mma_peak_tops"] = [lbc.probe_int8_mma_peak(it) for it in (2048, 16384, 65536)]
        except Exception as e:  # noqa: BLE001
            out["int8_mma_peak_error"] = str(e)
        print("DIAG_RESULT " + json.dumps(out), flush=True)
        return

    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    results = []

    def sub(args, timeout=180):
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__)] + args, capture_output=True, text=True,
                               timeout=timeout, cwd=ROOT)
        except subprocess.TimeoutExpired:
            return {"status": "timeout"}
        for line in r.stdout.splitlines():
            if line.startswith("DIAG_RESULT "):
                return json.loads(line[len("DIAG_RESULT "):])
        return {"status": "crash", "rc": r.returncode, "stderr": r.stderr[-1500:], "stdout": r.stdout[-500:]}

    results.append({"label": "probes", **sub(["--probes"], timeout=300)})
    print(json.dumps(results[-1]), flush=True)
    for i, c in enumerate(CASES):
        if a.only and a.only not in c[0]:
            continue
        res = sub(["--case", str(i)])
        res.setdefault("label", c[0])
        results.append(res)
        print(json.dumps(res)[:1500], flush=True)
        with open(a.out, "w") as fh:
            json.dump(results, fh, indent=1)
    ok = sum(1 for r in results if r.get("status") == "ok")
    print(f"DIAG SUMMARY: {ok}/{len(results) - 1} cases ok", flush=True)


if __name__ == "__main__":
    main()
