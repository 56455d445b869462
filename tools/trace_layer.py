"""trace_layer.py — dump CTA 0's pipeline time line (clock64 stamps) for one layer of a benchmark network."""
import argparse
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
EV = ["P_ISSUE", "W_ISSUE", "M_START", "M_WIN", "M_FULL", "M_DONE", "E_START", "E_DRAIN", "E_STORE", "P_DONE"]


def handover(lbc, torch, d, name, options):
    """Two plans of the same layer, launched A, B, A, B: where launch B's CTAs are (entered, set up, released by
    griddepcontrol.wait, finished) on the wall clock of launch A's last CTA."""
    dev = torch.device("cuda:0")
    cg = d.c // d.groups
    runs = []
    for _ in range(2):
        plan = lbc.ConvPlan(d, options=options)
        w = torch.randint(-127, 128, (d.k * d.r * d.s * cg,), dtype=torch.int8, device=dev)
        wp = plan.prepack(w)
        x = torch.randint(-128, 128, (d.n, d.h, d.w, d.c), dtype=torch.int8, device=dev)
        bias = torch.randint(-1000, 1000, (d.k,), dtype=torch.int32, device=dev)
        scale = torch.full((d.k,), 2.0**-7 / (d.r * d.s * cg) ** 0.5, dtype=torch.float32, device=dev)
        y = plan.empty_output(dev)
        buf = torch.zeros(16 + 6 * 148, dtype=torch.int64, device=dev)
        runs.append((plan, x, wp, bias, scale, y, buf))
    for plan, x, wp, bias, scale, y, buf in runs:
        plan.run(x, wp, bias, scale, out=y)
        plan.set_trace(buf, 1)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(2):
        for plan, x, wp, bias, scale, y, buf in runs:
            plan.run(x, wp, bias, scale, out=y)
    e1.record()
    torch.cuda.synchronize()
    g = []
    for plan, x, wp, bias, scale, y, buf in runs:
        raw = buf.cpu().numpy()[16 + 2 * 148:].reshape(148, 4)
        g.append(raw[raw[:, 0] > 0])
        plan.set_trace(None, 0)
    a_, b_ = g                                   # the stamps left behind are those of the LAST launch of each plan: A then B
    t_ref = a_[:, 3].max()                       # the last CTA of launch A issues its last store

    def q(v):
        v = (v - t_ref) / 1e3
        return f"min {v.min():7.2f}  median {np.median(v):7.2f}  max {v.max():7.2f}"
    print(f"== {name} x4 back to back: {e0.elapsed_time(e1) * 1e3 / 4:.1f} us per launch | {runs[0][0].describe()[:110]}")
    print(f"   launch A, last store issued   (us): {q(a_[:, 3])}")
    print(f"   launch B, kernel entry        (us): {q(b_[:, 0])}")
    print(f"   launch B, set-up done         (us): {q(b_[:, 1])}    set-up itself: median {np.median(b_[:, 1] - b_[:, 0]) / 1e3:.2f}")
    print(f"   launch B, past the dep. wait  (us): {q(b_[:, 2])}")
    print(f"   launch B, last store issued   (us): {q(b_[:, 3])}    working time: median {np.median(b_[:, 3] - b_[:, 2]) / 1e3:.2f}")
    for r in runs:
        r[0].close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--network", default="resnet50")
    ap.add_argument("--layers", default="l1.1.conv2")
    ap.add_argument("--tiles", type=int, default=24)
    ap.add_argument("--skip", type=int, default=6)
    ap.add_argument("--opt", action="append", default=[], help="planner option key=value (lbc_plan_options field), repeatable")
    ap.add_argument("--handover", action="store_true",
                    help="run each layer as two plans back to back (A, B, A, B on one stream) and print, from the %%globaltimer "
                         "stamps of every CTA, how launch B starts relative to the end of launch A")
    a = ap.parse_args()
    import torch
    import lowbitdnn_project_b200 as lbc
    from lowbitdnn_project_b200 import _capi
    # the build with the pipeline stamps compiled in (lowbitdnn-project_b200/build.py makes both)
    _capi._lib = _capi.load_library(os.path.join(os.path.dirname(_capi.LIB_PATH), "liblowbit_cnn_trace.so"))
    nets = lbc.networks
    layers = nets.NETWORKS[a.network](nets.DEFAULT_BATCH[a.network])
    dev = torch.device("cuda:0")
    for i, (name, d, _) in enumerate(layers):
        if name not in a.layers.split(","):
            continue
        if a.handover:
            handover(lbc, torch, d, name, {k: int(v) for k, v in (o.split("=") for o in a.opt)} or None)
            continue
        plan = lbc.ConvPlan(d, options={k: int(v) for k, v in (o.split("=") for o in a.opt)} or None)
        cg = d.c // d.groups
        w = torch.randint(-127, 128, (d.k * d.r * d.s * cg,), dtype=torch.int8, device=dev)
        wp = plan.prepack(w)
        x = torch.randint(-128, 128, (d.n, d.h, d.w, d.c), dtype=torch.int8, device=dev)
        bias = torch.randint(-1000, 1000, (d.k,), dtype=torch.int32, device=dev)
        scale = torch.full((d.k,), 2.0**-7 / (d.r * d.s * cg) ** 0.5, dtype=torch.float32, device=dev)
        y = plan.empty_output(dev)
        for _ in range(2):
            plan.run(x, wp, bias, scale, out=y)
        ntile = a.tiles + a.skip
        buf = torch.zeros(ntile * 16 + 6 * 148, dtype=torch.int64, device=dev)
        plan.set_trace(buf, ntile)
        _, ms = plan.run(x, wp, bias, scale, out=y, timed=True)
        plan.set_trace(None, 0)
        raw = buf.cpu().numpy()
        t = raw[:ntile * 16].reshape(ntile, 16)
        cta = raw[ntile * 16:ntile * 16 + 2 * 148].reshape(148, 2)
        cta = cta[cta[:, 0] > 0]
        if len(cta):
            dur = (cta[:, 1] - cta[:, 0]) / 1965.0          # SM cycles -> us at the 1965 MHz boost clock
            order = np.argsort(dur)
            print(f"   per-CTA duration (us @1965 MHz): min {dur.min():.1f} median {np.median(dur):.1f} max {dur.max():.1f}; "
                  f"slowest {[(int(i), round(float(dur[i]), 1)) for i in order[-4:]]} fastest {[(int(i), round(float(dur[i]), 1)) for i in order[:3]]}")
        have = [k for k in range(ntile) if (t[k] > 0).any()]       # CTA 0 may own fewer tiles than asked for
        if not have:
            print(f"== {name}: no tiles traced")
            plan.close()
            continue
        skip = min(a.skip, max(0, len(have) - 4))
        ntile = have[-1] + 1
        t0 = t[skip][t[skip] > 0].min()
        print(f"== {name}: {ms * 1e3:.1f} us | {plan.describe()}")
        print("tile " + " ".join(f"{e:>8s}" for e in EV) + "   | dM(start->done) dE(start->drain) tile-to-tile(M_DONE)")
        prev = None
        for k in range(skip, ntile):
            row = t[k]
            cells = " ".join(f"{(int(row[e]) - int(t0)) if row[e] else -1:8d}" for e in range(len(EV)))
            dm = int(row[5] - row[2]) if row[5] and row[2] else -1
            de = int(row[7] - row[6]) if row[7] and row[6] else -1
            dd = int(row[5] - prev) if prev else -1
            prev = row[5]
            print(f"{k:4d} {cells}   | {dm:6d} {de:6d} {dd:6d}")
        plan.close()


if __name__ == "__main__":
    main()
