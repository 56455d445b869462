"""two_streams.py — experiment: the benchmark batch as M micro-batches whose layer chains run on M streams.

A layer's launch leaves SMs idle while it fills and drains (7-8 us per launch); an independent chain
on a second stream can use exactly those SMs.  Prints ms per step of the whole batch for M = 1, 2, 3, 4.
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--network", default="resnet50")
    ap.add_argument("--batch", type=int, default=512)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--splits", default="1,2,4")
    ap.add_argument("--opt", action="append", default=[], help="planner option key=value, repeatable")
    a = ap.parse_args()
    import torch
    import lowbitdnn_project_b200 as lbc
    main_s = torch.cuda.current_stream()
    for m in [int(v) for v in a.splits.split(",")]:
        nb = a.batch // m
        nets, streams = [], []
        for j in range(m):
            layers = lbc.networks.NETWORKS[a.network](nb)
            net = lbc.Net(layers, options={k: int(v) for k, v in (o.split("=") for o in a.opt)} or None)
            for i, (_, d, src) in enumerate(layers):
                cg = d.c // d.groups
                rw = np.random.default_rng(4321 + i)
                w = rw.integers(-127, 128, size=(d.k, d.r, d.s, cg), dtype=np.int8)
                bias = rw.integers(-2**15, 2**15, size=(d.k,), dtype=np.int32)
                scale = (rw.uniform(0.5, 2.0, size=(d.k,)) * 2.0**-7 / np.sqrt(d.r * d.s * cg)).astype(np.float32)
                net.set_params(i, w, bias, scale)
                if src is None:
                    net.set_input(i, np.random.default_rng(1234 + i + 100 * j).integers(-128, 128, size=(d.n, d.h, d.w, d.c), dtype=np.int8))
            nets.append(net)
            streams.append(torch.cuda.Stream())
        fork, joins = torch.cuda.Event(), [torch.cuda.Event() for _ in range(m)]

        def step():
            fork.record(main_s)
            for net, s, j in zip(nets, streams, joins):
                s.wait_event(fork)
                net.run(stream=s)
                j.record(s)
            for j in joins:
                main_s.wait_event(j)

        for _ in range(3):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(main_s)
        for _ in range(a.steps):
            step()
        e1.record(main_s)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.steps
        for net in nets:
            net.check_status()
            net.close()
        print(f"{a.network} batch {a.batch} as {m} x {nb}: {ms:.3f} ms per step, {a.batch / ms * 1e3:.0f} images/s", flush=True)


if __name__ == "__main__":
    main()
