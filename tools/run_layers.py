"""run_layers.py — run selected layers of a benchmark network a few times each (profiling driver for ncu).

    python tools/run_layers.py --network resnet50 --batch 512 --layers l1.1.conv2,l1.1.conv3 --iters 3
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--network", default="resnet50")
    ap.add_argument("--batch", type=int, default=None)
    ap.add_argument("--layers", default="")
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--opt", action="append", default=[], help="planner option key=value (lbc_plan_options field), repeatable")
    a = ap.parse_args()
    import torch
    import lowbitdnn_project_b200 as lbc
    nets = lbc.networks
    batch = a.batch or nets.DEFAULT_BATCH[a.network]
    layers = nets.NETWORKS[a.network](batch)
    want = [s for s in a.layers.split(",") if s]
    dev = torch.device("cuda:0")
    for i, (name, d, _) in enumerate(layers):
        if want and name not in want:
            continue
        plan = lbc.ConvPlan(d, options={k: int(v) for k, v in (o.split("=") for o in a.opt)} or None)
        rng = np.random.default_rng(i)
        cg = d.c // d.groups
        w = torch.from_numpy(rng.integers(-127, 128, size=(d.k * d.r * d.s * cg,), dtype=np.int8)).to(dev)
        wp = plan.prepack(w)
        x = torch.randint(-128, 128, (d.n, d.h, d.w, d.c), dtype=torch.int8, device=dev)
        bias = torch.randint(-1000, 1000, (d.k,), dtype=torch.int32, device=dev)
        scale = torch.full((d.k,), 2.0**-7 / (d.r * d.s * cg) ** 0.5, dtype=torch.float32, device=dev)
        y = plan.empty_output(dev)
        ms = []
        for _ in range(a.iters):
            lbc.flush_l2()
            _, t = plan.run(x, wp, bias, scale, out=y, timed=True)
            ms.append(t)
        ops, byts = d.work
        best = min(ms)
        print(f"{name:16s} {plan.kernel:9s} best {best * 1e3:8.1f} us  {ops / best / 1e9:7.1f} TOPS  {byts / best / 1e6:7.0f} GB/s | {plan.describe()}",
              flush=True)
        plan.close()


if __name__ == "__main__":
    main()
