#!/usr/bin/env python
"""
bench.py — headline benchmark of liblowbit-cnn on B200: ResNet-50 int8 convolution stack, images/s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference|reference-gpu] [--network resnet50]
                    [--batch B] [--scaling weak|strong]

A "step" is one pass of the hot path (all 53 convolutions of ResNet-50, batch 512 per GPU, fused
bias/requant/ReLU epilogues) over one batch of synthetic int8 input.  Rank 0 prints ONE JSON line.

  value      images/s over all ranks with inputs resident in HBM (CUDA events, max over ranks)
  e2e        same metric through lbc_net_run_host: pinned HOST input -> H2D -> 53 layers -> D2H -> HOST output
  roofline   igemm_i8_kernel (the dominant kernel): algorithmic ops/bytes per launch (SURVEY 8d formulas) over
             its event-timed launch durations inside this run, against MEASURED_PEAKS.json / the int8 MMA probe
  cpu_baseline  the oracle port (oracle/cpu_ref.c, OpenMP) timed on this host on a bounded sample

--impl reference times the reference's own CPU implementation (cpp/int8conv/refConv2DForward.hpp compiled
unmodified into oracle/_ref) on a bounded sample, on rank 0 only.

Multi-GPU: one process per GPU (torchrun), the batch dimension is sharded and the weights are replicated; there is
no collective on the convolution path.  NCCL is used only to gather per-rank timings and output checksums.
  --scaling weak   (default) every rank owns `batch` images (512 per GPU): per-GPU work fixed as N grows
  --scaling strong the configuration's batch (512) is the GLOBAL batch, split over the ranks with
                   shard.image_range (64 images per GPU at 8 GPUs)

After the timed region every rank re-computes the first images of EVERY layer's output with the CPU oracle chained
over the same graph and compares them with the resident GPU outputs; a mismatch makes the run fail (exit 3).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "resnet50_int8_conv_images_per_sec"
UNIT = "images/s"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ------------------------------------------------------------------------------------------------------
# clocks sampler (NVML; the recipe's "clocks line")
# ------------------------------------------------------------------------------------------------------
class ClockSampler:
    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception as e:  # noqa: BLE001
            self.nv = None
            log("clock sampler unavailable:", e)

    def _loop(self):
        nv = self.nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4),
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if mask & bit:
                        self.reasons.add(k)
            except Exception:  # noqa: BLE001
                pass
            self._stop.wait(0.002)

    def start(self):
        if self.nv:
            self._thr = threading.Thread(target=self._loop, daemon=True)
            self._thr.start()

    def stop(self):
        self._stop.set()
        if self._thr:
            self._thr.join()
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# ------------------------------------------------------------------------------------------------------
def synth_params(d, layer: int):
    """SURVEY.md 8d synthetic parameters (same generator as oracle.synth, restated to keep the product/bench
    path free of oracle imports on the GPU arm)."""
    cg = d.c // d.groups
    rw = np.random.default_rng(4321 + layer)
    w = rw.integers(-127, 128, size=(d.k, d.r, d.s, cg), dtype=np.int8)
    bias = rw.integers(-2**15, 2**15, size=(d.k,), dtype=np.int32)
    scale = (rw.uniform(0.5, 2.0, size=(d.k,)) * 2.0**-7 / np.sqrt(d.r * d.s * cg)).astype(np.float32)
    return w, bias, scale


def synth_input(d, layer: int):
    return np.random.default_rng(1234 + layer).integers(-128, 128, size=(d.n, d.h, d.w, d.c), dtype=np.int8)


def layer_work(d):
    if not hasattr(d, "groups"):        # graph nodes between convolutions: pure byte movers
        p, q = d.out_hw
        if hasattr(d, "kh"):            # max-pool: input once, output once
            return 0.0, float(d.n * d.h * d.w * d.c + d.n * p * q * d.c)
        return 0.0, float(3 * d.n * d.h * d.w * d.c)      # residual add: two operands in, one out
    p = (d.h + 2 * d.pad_h - (d.dil_h * (d.r - 1) + 1)) // d.stride_h + 1
    q = (d.w + 2 * d.pad_w - (d.dil_w * (d.s - 1) + 1)) // d.stride_w + 1
    cg = d.c // d.groups
    ops = 2.0 * d.n * p * q * d.k * cg * d.r * d.s
    byts = d.n * d.h * d.w * d.c + d.k * cg * d.r * d.s + d.n * p * q * d.k + 8 * d.k
    return ops, float(byts)


def read_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            p = json.load(fh)
        return {"hbm_gbs": float(p["hbm_gbs"]), "bf16_tflops": float(p.get("bf16_tflops", 0)),
                "bf16_tflops_sustained": float(p.get("bf16_tflops_sustained", 0)), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


# ------------------------------------------------------------------------------------------------------
# CPU legs (the only places that touch oracle/)
# ------------------------------------------------------------------------------------------------------
def cpu_baseline_port(layers, budget_s: float = 15.0):
    """Oracle port (oracle/cpu_ref.c, all host threads) on the same network at batch 1, repeated until
    ~budget_s of CPU work; returns images/s."""
    from oracle import oracle
    from oracle.oracle import ConvDesc as OD
    threads = oracle.max_threads()
    prepared = []
    for i, (_, d, _) in enumerate(layers):
        if not hasattr(d, "groups"):
            continue                    # (pool / add nodes: negligible CPU work, convolutions only)
        od = OD(**{**d.__dict__, "n": 1})
        x, w, b, s = oracle.synth(od, layer=i)
        prepared.append((od, x, w, b, s))
    t0 = time.perf_counter()
    images = 0
    while True:
        for od, x, w, b, s in prepared:
            oracle.conv_nhwc(od, x, w, b, s)
        images += 1
        el = time.perf_counter() - t0
        if el > budget_s or images >= 4096:
            break
    return {"value": images / el, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"{images} image(s) x all {len(layers)} layers at batch 1, oracle/cpu_ref.c with {threads} OpenMP threads, {el:.1f} s"}


def parity_check(net, layers, rank, images=2):
    """The checker leg: first `images` images of every layer's resident output against the oracle chained over the same
    graph (same synthetic inputs and parameters as the timed run).  Returns (layers checked, list of mismatching layers)."""
    from oracle import oracle
    from oracle.oracle import ConvDesc as OD
    outs, bad = {}, []
    for i, (name, d, src) in enumerate(layers):
        n = min(images, d.n)
        if not hasattr(d, "groups"):    # max-pool / residual add nodes of a graph network
            if hasattr(d, "kh"):
                outs[name] = oracle.max_pool_nhwc(outs[src], d.kh, d.kw, d.stride_h, d.stride_w, d.pad_h, d.pad_w)
            else:
                outs[name] = oracle.add_relu(outs[src[0]], outs[src[1]], bool(d.relu))
            if not np.array_equal(net.read_output(i, images=n), outs[name]):
                bad.append(name)
            continue
        x = synth_input(d, i + 1000 * rank)[:n] if src is None else outs[src]
        w, b, s = synth_params(d, i)
        outs[name] = oracle.conv_nhwc(OD(**{**d.__dict__, "n": n}), np.ascontiguousarray(x), w, b, s)
        if net.fused_into(i) >= 0:
            continue              # absorbed by its consumer's fused launch (checked through the consumer's output)
        got = net.read_output(i, images=n)
        if not np.array_equal(got, outs[name]):
            bad.append(name)
    return len(layers), bad


def host_link_probe(nbytes_h2d, nbytes_d2h, barrier, iters=5):
    """Pinned-memory copy bandwidth of this rank's host link with every rank copying at the same time (both
    directions at once, the way the pipelined e2e path uses the link).  Returns (h2d GB/s, d2h GB/s)."""
    import torch
    hx = torch.empty(nbytes_h2d, dtype=torch.uint8).pin_memory()
    hy = torch.empty(nbytes_d2h, dtype=torch.uint8).pin_memory()
    dx = torch.empty(nbytes_h2d, dtype=torch.uint8, device="cuda")
    dy = torch.empty(nbytes_d2h, dtype=torch.uint8, device="cuda")
    s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    for timed in (False, True):
        barrier()
        with torch.cuda.stream(s_up):
            ev[0].record()
            for _ in range(iters):
                dx.copy_(hx, non_blocking=True)
            ev[1].record()
        with torch.cuda.stream(s_dn):
            ev[2].record()
            for _ in range(iters):
                hy.copy_(dy, non_blocking=True)
            ev[3].record()
        torch.cuda.synchronize()
    up = nbytes_h2d * iters / (ev[0].elapsed_time(ev[1]) * 1e-3) / 1e9
    dn = nbytes_d2h * iters / (ev[2].elapsed_time(ev[3]) * 1e-3) / 1e9
    return up, dn


REF_SAMPLE = (1, 64, 10, 10, 64, 8, 8, 3, 3)   # config 1 (56x56x64->64 3x3) cropped to 8x8 outputs, pre-padded


def reference_arm(args, layers):
    """The reference's own CPU path (refConv2DForward.hpp, unmodified, oracle/_ref) on a bounded sample."""
    from oracle import oracle
    if not oracle.have_ref():
        emit_json({"impl": "reference", "unavailable": "oracle/_ref/libref_conv.so missing (built only where /root/reference exists)"})
        return
    # all the host threads: torchrun exports OMP_NUM_THREADS=1 to every rank, and under it only rank 0 runs this arm
    try:
        import ctypes
        ctypes.CDLL("libgomp.so.1").omp_set_num_threads(len(os.sched_getaffinity(0)))
    except Exception as e:  # noqa: BLE001
        log("could not raise the OpenMP thread count:", e)
    b, ic, ih, iw, oc, oh, ow, kh, kw = REF_SAMPLE
    rng = np.random.default_rng(99)
    x = rng.integers(-128, 128, size=(b, ic, ih, iw), dtype=np.int8)
    w = rng.integers(-128, 128, size=(oc, ic, kh, kw), dtype=np.int8)
    sample_macs = b * oc * oh * ow * ic * kh * kw
    net_macs_per_image = sum(layer_work(d)[0] for _, d, _ in layers) / 2.0 / layers[0][1].n
    for _ in range(args.warmup):
        oracle.ref_conv2d_forward(x, w)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        y = oracle.ref_conv2d_forward(x, w)
    el = time.perf_counter() - t0
    assert np.array_equal(y, oracle.ref_style_nchw_valid(x, w))
    ms = el / args.steps * 1e3
    value = (sample_macs / (ms * 1e-3)) / net_macs_per_image
    cores = oracle.ref_max_threads()
    sample = (f"refConv2DForward<1,64,10,10,64,8,8,3,3> per step ({sample_macs / 1e6:.2f} MMAC: BASELINE config 1 cropped to 8x8 "
              f"outputs), {cores} OpenMP threads; images/s = sample MAC rate / {net_macs_per_image / 1e9:.3f} GMAC per ResNet-50 image "
              "(extrapolated from the reference CPU path)")
    out = {
        "impl": "reference", "metric": METRIC if args.network == "resnet50" else f"{args.network}_int8_conv_images_per_sec", "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "int8", "data": "synthetic",
        "config": {"workload": f"{args.network}_conv_stack_b{args.batch}", "network": args.network,
                   "batch_per_gpu": args.batch, "layers": len(layers)},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "reference", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    if args.ref_full:
        # BASELINE config 1 in full (N1, 56x56x64 -> 64, 3x3 s1 p1, pre-padded to 58x58): the one configuration the
        # reference CPU path can run as configured; minutes, so opt-in
        fb, fic, fih, fiw, foc, foh, fow, fkh, fkw = 1, 64, 58, 58, 64, 56, 56, 3, 3
        xf = np.zeros((fb, fic, fih, fiw), dtype=np.int8)
        xf[:, :, 1:-1, 1:-1] = rng.integers(-128, 128, size=(fb, fic, 56, 56), dtype=np.int8)
        wf = rng.integers(-128, 128, size=(foc, fic, fkh, fkw), dtype=np.int8)
        t0 = time.perf_counter()
        yf = oracle.ref_conv2d_forward(xf, wf)
        sec = time.perf_counter() - t0
        assert np.array_equal(yf, oracle.ref_style_nchw_valid(xf, wf))
        macs = fb * foc * foh * fow * fic * fkh * fkw
        out["config1_full"] = {"shape": "refConv2DForward<1,64,58,58,64,56,56,3,3>", "seconds": sec, "mmac_per_s": macs / sec / 1e6,
                               "threads": cores, "images_per_s_extrapolated": (macs / sec) / net_macs_per_image,
                               "checked": "equal to the oracle restatement"}
    emit_json(out)


def reference_gpu_arm(args):
    """The reference's own tensor-core kernel (CUDAConv2DForward3x3TensorCoures, wmma m32n8k16, compiled unmodified for
    sm_100a from /root/reference into oracle/_ref/libref_wmma.so with a torch-free host) on ITS OWN benchmark shape
    (check.cu:31-41: 16x128x130x130 (*) 128x128x3x3, VALID, int32 out), next to liblowbit-cnn on the same shape and bytes.
    Checker / yardstick only: nothing here is on the product path."""
    import ctypes
    import torch
    import lowbitdnn_project_b200 as lbc
    so = os.path.join(ROOT, "oracle", "_ref", "libref_wmma.so")
    if not os.path.exists(so):
        emit_json({"impl": "reference-gpu", "unavailable": "oracle/_ref/libref_wmma.so missing (built only where /root/reference exists: make -C oracle ref_gpu)"})
        return
    assert torch.cuda.is_available()
    ref = ctypes.CDLL(so)
    ref.ref_wmma_conv3x3.restype = ctypes.c_float
    ref.ref_wmma_conv3x3.argtypes = [ctypes.c_void_p] * 3 + [ctypes.c_int]
    shp = (ctypes.c_int32 * 7)()
    ref.ref_wmma_shape(shp)
    n, c, h, w, k, p, q = list(shp)
    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(7)
    x = torch.from_numpy(rng.integers(-128, 128, size=(n, h, w, c), dtype=np.int8)).to(dev)
    wk = torch.from_numpy(rng.integers(-127, 128, size=(k, 3, 3, c), dtype=np.int8)).to(dev)
    x_v = lbc.nhwc_to_vect_c(x, 16)                      # [N][C/16][H][W][16]  (utils.cuh:20-26)
    w_v = lbc.nhwc_to_vect_c(wk, 16)                     # [K][C/16][3][3][16]
    y_v = torch.empty((n, k // 16, p, q, 16), dtype=torch.int32, device=dev)
    ops = 2.0 * n * p * q * k * c * 9

    def timed(fn):
        for _ in range(max(args.warmup, 2)):
            fn()
        ms = []
        for _ in range(max(args.steps, 5)):
            lbc.flush_l2()
            torch.cuda.synchronize()
            ms.append(fn())
        return float(np.median(ms))

    ms_ref = timed(lambda: ref.ref_wmma_conv3x3(x_v.data_ptr(), w_v.data_ptr(), y_v.data_ptr(), 1))
    assert ms_ref > 0, "the reference kernel failed to launch"
    out = {"impl": "reference-gpu", "metric": "int8_conv3x3_tops", "unit": "TOPS", "higher_is_better": True, "data": "synthetic",
           "config": {"workload": f"reference check.cu shape: N{n} {h}x{w}x{c} -> {k}, 3x3 VALID, int32 out", "l2": "flushed before every launch"},
           "reference": {"kernel": "CUDAConv2DForward3x3TensorCoures (wmma m32n8k16, unmodified, sm_100a)", "ms": ms_ref,
                         "tops": ops / ms_ref / 1e9}}
    for mode, name in ((lbc.OUT_INT32, "ours_int32"), (lbc.OUT_INT8, "ours_int8_fused_epilogue")):
        plan = lbc.ConvPlan(lbc.ConvDesc(n=n, h=h, w=w, c=c, k=k, r=3, s=3, relu=1, out_mode=mode))
        wp = plan.prepack(wk.reshape(-1), lbc.W_KRSC)
        bias = torch.zeros(k, dtype=torch.int32, device=dev)
        scale = torch.full((k,), 2.0**-11, dtype=torch.float32, device=dev)
        y = plan.empty_output(dev)
        ms = timed(lambda: plan.run(x, wp, bias, scale, out=y, timed=True)[1])
        out[name] = {"ms": ms, "tops": ops / ms / 1e9, "plan": plan.describe()}
        if mode == lbc.OUT_INT32:
            same = bool(torch.equal(lbc.nhwc_to_vect_c(y, 16), y_v))
            out["parity_int32_vs_reference_kernel"] = "bit-exact" if same else "MISMATCH"
        plan.close()
    out["value"] = out["ours_int32"]["tops"]
    out["speedup_int32_vs_reference_kernel"] = ms_ref / out["ours_int32"]["ms"]
    out["speedup_int8_vs_reference_kernel"] = ms_ref / out["ours_int8_fused_epilogue"]["ms"]
    emit_json(out)
    if out["parity_int32_vs_reference_kernel"] != "bit-exact":
        sys.exit(3)


_JSON_FD = None


def emit_json(obj):
    """The one JSON line of the contract, on the process's original stdout."""
    line = (json.dumps(obj) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(line.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, line)


# ------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "reference-gpu"])
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: `batch` images per GPU; strong: `batch` is the global batch, split over the ranks")
    ap.add_argument("--network", default="resnet50")
    ap.add_argument("--batch", type=int, default=None, help="images per GPU (default: the config's batch)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--opt", action="append", default=[],
                    help="planner option key=value (lbc_plan_options field) applied to every layer; A/B experiments only")
    ap.add_argument("--ref-full", action="store_true",
                    help="--impl reference: also run BASELINE config 1 in full through the reference CPU path (minutes)")
    ap.add_argument("--layer-report", default=None, help="write the per-layer table (JSON) here")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else max(args.warmup, 1)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries exactly one JSON line: anything a library prints there (NCCL's version banner) goes to stderr
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)

    import lowbitdnn_project_b200 as lbc   # raises if liblowbit_cnn.so is missing — no fallback
    nets = lbc.networks
    if args.batch is None:
        args.batch = nets.DEFAULT_BATCH[args.network]
    config_batch = args.batch
    if args.scaling == "strong" and args.impl == "ours":
        first, last = lbc.shard.image_range(config_batch, world, rank)     # this rank's slice of the global batch
        args.batch = last - first
        assert args.batch > 0, f"rank {rank} of {world} has no image of a global batch of {config_batch}"
    layers = nets.NETWORKS[args.network](args.batch)

    if args.impl == "reference":
        if rank == 0:
            reference_arm(args, layers)
        return
    if args.impl == "reference-gpu":
        if rank == 0:
            reference_gpu_arm(args)
        return

    import torch
    import torch.distributed as dist
    assert torch.cuda.is_available(), "bench.py (ours) needs a B200; there is no CPU fallback"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    # ---- build the network, load synthetic parameters / resident inputs --------------------------------
    t_setup = time.time()
    net = lbc.Net(layers, options={k: int(v) for k, v in (o.split("=") for o in args.opt)} or None)
    for i, (_, d, src) in enumerate(layers):
        if not hasattr(d, "groups"):
            continue                    # pool / add nodes have no parameters
        w, b, s = synth_params(d, i)
        net.set_params(i, w, b, s)
        if src is None:
            net.set_input(i, synth_input(d, i + 1000 * rank))
    d0, dl = layers[0][1], layers[-1][1]
    pl, ql = dl.out_hw
    x_host = torch.from_numpy(synth_input(d0, 1000 * rank)).pin_memory()
    y_host = torch.empty((dl.n, pl, ql, dl.k if hasattr(dl, "k") else dl.c), dtype=torch.int8).pin_memory()
    stream = torch.cuda.current_stream()
    log(f"[rank {rank}] setup {time.time() - t_setup:.1f}s; kernels: "
        + ", ".join(f"{k}x{v}" for k, v in sorted(_count(net).items())))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up ---------------------------------------------------------------------------------------
    for _ in range(args.warmup):
        net.run(stream=stream)
    barrier()

    # ---- timed region: K steps back to back, inputs resident in HBM ---------------------------------------
    sampler = ClockSampler(local_rank)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler.start()
    barrier()
    ev0.record(stream)
    for _ in range(args.steps):
        net.run(stream=stream)
    ev1.record(stream)
    barrier()
    clocks = sampler.stop()
    ms_total = ev0.elapsed_time(ev1)

    # ---- instrumented pass: per-launch CUDA-event durations of every layer (same stream, same inputs) -----
    per_layer = np.zeros(len(layers))
    for _ in range(args.steps):
        per, _tot = net.run(stream=stream, timed=True)
        per_layer += np.array(per)
    per_layer /= args.steps

    # ---- e2e: host buffers through the C ABI, every step: pinned HOST input -> H2D -> 53 layers -> D2H -> HOST ----
    # Steps are submitted back to back through lbc_net_submit_host (the serving steady state): the library uploads
    # step i+1's input and downloads step i's output on its own copy streams while the layers of the neighbouring
    # steps run.  Each step has its own pinned input and output buffer; the clock is the device time from the first
    # upload to the last download (and the host wall clock around submit..sync, whichever is larger).
    n_e2e = args.steps
    xs = [x_host] + [torch.from_numpy(synth_input(d0, 1000 * rank + 7 * (i + 1))).pin_memory() for i in range(min(n_e2e, 4) - 1)]
    ys = [torch.empty_like(y_host).pin_memory() for _ in range(len(xs))]
    for _ in range(2):                                   # warm-up (also the serial, blocking call)
        net.run_host(x_host, y_host, stream=stream)
    for i in range(2):
        net.submit_host(xs[i % len(xs)], ys[i % len(ys)], stream=stream)
    net.sync_host()
    barrier()
    t0 = time.perf_counter()
    for i in range(n_e2e):
        net.submit_host(xs[i % len(xs)], ys[i % len(ys)], stream=stream)
    e2e_dev_ms = net.sync_host()
    e2e_wall_ms = (time.perf_counter() - t0) * 1e3
    e2e_ms = max(e2e_dev_ms, e2e_wall_ms) / n_e2e
    serial_ms = net.run_host(x_host, y_host, stream=stream)        # one blocking step, for the record
    checksum = zlib.crc32(y_host.numpy().tobytes())
    # the pipelined path must give the same bytes as the blocking one for the same input
    assert zlib.crc32(ys[0].numpy().tobytes()) == checksum, "pipelined e2e output differs from the blocking path"

    # ---- host link ceiling of the e2e path: all ranks copy at once, both directions ----------------------------
    link_up, link_dn = host_link_probe(int(x_host.numel()), int(y_host.numel()), barrier)

    # ---- parity of what was just timed: every layer's resident output (after the last run) against the oracle ----
    n_checked, bad_layers = parity_check(net, layers, rank, images=2)
    if bad_layers:
        log(f"[rank {rank}] PARITY FAILURE in layers: {bad_layers}")

    # ---- gather (NCCL): max over ranks; checksums ------------------------------------------------------------
    job = lbc.shard.gather(lbc.shard.RankStats(ms_total, e2e_ms, checksum, args.batch, len(bad_layers), link_up, link_dn),
                           world, device=dev)
    ms_total, e2e_ms = job.ms_total, job.e2e_ms
    ms_per_step = ms_total / args.steps
    images_per_step = job.images

    if rank == 0:
        peaks = read_peaks()
        kinds = [net.layer_kernel(i) for i in range(len(layers))]
        works = [layer_work(d) for _, d, _ in layers]
        # dominant kernel = the one with the largest share of the step
        # the small-C stem runs the same igemm_i8_kernel (after its space-to-depth pass), so it counts with it
        group = {"stem_tc": "igemm_tc", "maxpool": "pool_add", "add_relu": "pool_add"}
        share = {}
        for k, ms in zip(kinds, per_layer):
            share[group.get(k, k)] = share.get(group.get(k, k), 0.0) + ms
        dom = max(share, key=share.get)
        sel = [i for i, k in enumerate(kinds) if group.get(k, k) == dom]
        # The per-layer pass records an event after every launch, which breaks the programmatic-dependent-launch overlap
        # between layers: its sum exceeds the un-instrumented step.  Only the SHARES come from it; every time below is
        # that share of the timed region's ms_per_step, so kernel_ms_per_step <= ms_per_step by construction.
        instrumented_ms = float(per_layer.sum())
        per_layer = per_layer * (ms_per_step / instrumented_ms)
        dom_ms = sum(per_layer[i] for i in sel)
        dom_ops = sum(works[i][0] for i in sel)
        dom_bytes = sum(works[i][1] for i in sel)
        int8_peak = int8_sustained = None
        try:
            int8_peak = lbc.probe_int8_mma_peak(16384)            # a few ms: burst clocks
            t_end = time.perf_counter() + 0.6                      # ~0.6 s of back-to-back MMA: settles under the power cap
            while time.perf_counter() < t_end:
                int8_sustained = lbc.probe_int8_mma_peak(262144)
        except Exception as e:  # noqa: BLE001
            log("int8 peak probe failed:", e)
        tops = dom_ops / (dom_ms * 1e-3) / 1e12
        gbs = dom_bytes / (dom_ms * 1e-3) / 1e9
        # roofline time of the dominant kernel's launches: each launch is bound by its slower side
        tc_peak = (int8_peak or 2 * peaks["bf16_tflops"]) * 1e12
        roof_ms = sum(max(works[i][0] / tc_peak, works[i][1] / (peaks["hbm_gbs"] * 1e9)) for i in sel) * 1e3
        hbm_bound_ms = sum(per_layer[i] for i in sel if works[i][1] / (peaks["hbm_gbs"] * 1e9) >= works[i][0] / tc_peak)
        bound = "hbm" if hbm_bound_ms >= dom_ms / 2 else "tensor"
        if bound == "hbm":
            roofline = {"bound": "hbm", "achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                        "frac": gbs / peaks["hbm_gbs"], "traffic": None}
        else:
            roofline = {"bound": "tensor", "achieved": tops, "peak": tc_peak / 1e12, "unit": "TFLOP/s",
                        "frac": tops / (tc_peak / 1e12), "traffic": None}
        # DRAM traffic per launch of the dominant kernel: from the committed ncu launch list of this same command
        # (profiles/*_traffic.json, written by tools/make_profile_summary.py); only valid for the default workload
        try:
            tfiles = sorted(f for f in os.listdir(os.path.join(ROOT, "profiles")) if f.endswith("_traffic.json"))
            if tfiles and args.network == "resnet50" and args.batch == nets.DEFAULT_BATCH["resnet50"] and dom == "igemm_tc":
                with open(os.path.join(ROOT, "profiles", tfiles[-1])) as fh:
                    tr = json.load(fh)
                roofline["traffic"] = tr["dram_bytes_per_launch"] / 1e9 if roofline["unit"] == "GB/s" else tr["dram_bytes_per_launch"]
                roofline["traffic_unit"] = "GB per launch (dram__bytes_read.sum + dram__bytes_write.sum, mean over the kernel's launches of one step)"
                roofline["traffic_source"] = "profiles/" + tfiles[-1]
                roofline["algorithmic_gb_per_launch"] = dom_bytes / len(sel) / 1e9
                # the same launch list's share of the step (cold-cache, serialised launches); the live share below counts the
                # stem's space-to-depth pre-pass (1.7% of the step, its own small kernel) with the stem's igemm launch
                roofline["kernel_share_of_step_ncu"] = tr.get("share_of_step_ncu")
        except Exception as e:  # noqa: BLE001
            log("traffic file unreadable:", e)
        roofline.update({
            "kernel": {"igemm_tc": "igemm_i8_kernel", "direct": "direct_conv_kernel", "depthwise": "depthwise_kernel",
                       "pool_add": "maxpool / add_relu kernels"}[dom],
            "launches_per_step": len(sel), "kernel_ms_per_step": dom_ms, "kernel_share_of_step": dom_ms / ms_per_step,
            "instrumented_pass_ms": instrumented_ms,
            "timing_note": "kernel_ms_per_step = the kernel's share of the per-launch event pass x ms_per_step of the timed region "
                           "(per-LAYER events: the stem layer's time includes its space-to-depth pre-pass)",
            "peak_source": f"{peaks['source']} (MEASURED_PEAKS.json hbm_gbs); tensor fractions are against the BURST on-box tcgen05 "
                           "kind::i8 MMA-only probe (int8_mma_peak_tops), nominal dense int8 is 4500 TOPS",
            "achieved_tops": tops, "achieved_gbs": gbs, "int8_mma_peak_tops": int8_peak,
            "int8_mma_sustained_tops": int8_sustained,
            "roofline_ms": roof_ms, "frac_of_mixed_roofline": roof_ms / dom_ms,
        })
        total_ops = sum(w[0] for w in works)
        out = {
            "metric": METRIC if args.network == "resnet50" else f"{args.network}_int8_conv_images_per_sec", "value": images_per_step / (ms_per_step * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "int8", "data": "synthetic",
            "config": {"workload": f"{args.network}_conv_stack_b{config_batch}", "network": args.network,
                       "batch_per_gpu": args.batch, "global_batch": images_per_step, "layers": len(layers),
                       "parallelism": f"batch-sharded x{world} ({args.scaling} scaling), no collective on the conv path",
                       **({"planner_options": args.opt} if args.opt else {}),
                       "l2": f"no flush: a step touches {sum(w[1] for w in works) / 1e9:.2f} GB (algorithmic) against 126 MB of L2; "
                             f"the smallest layer input is {min(d.n * d.h * d.w * d.c for _, d, _ in layers) / 1e6:.1f} MB"},
            "parity": "bit-exact" if job.bad_layers == 0 else f"MISMATCH in {job.bad_layers} layer outputs (summed over ranks)",
            "parity_checked_images": 2, "parity_checked_layers": n_checked,
            "tops_per_gpu": total_ops / (ms_per_step * 1e-3) / 1e12,
            "clocks": clocks,
            "e2e": {"value": images_per_step / (e2e_ms * 1e-3), "unit": UNIT,
                    "h2d_bytes_per_step": int(x_host.numel()), "d2h_bytes_per_step": int(y_host.numel()),
                    "ms_per_step": e2e_ms,
                    "serial_ms_per_step": serial_ms,
                    "host_link_gbs": {"h2d_per_rank_min": job.link_up, "d2h_per_rank_min": job.link_dn,
                                      "how": f"pinned copies of the step's input / output sizes on {world} rank(s) at once, both directions at once"},
                    "host_link_ceiling": args.batch / max(int(x_host.numel()) / (job.link_up * 1e9), int(y_host.numel()) / (job.link_dn * 1e9),
                                                         ms_per_step * 1e-3) * world,
                    "note": "lbc_net_submit_host x steps + lbc_net_sync_host: every step copies its own pinned host input H2D, runs all "
                            "layers and copies the last layer's output D2H; copies of neighbouring steps overlap compute. "
                            "serial_ms_per_step = one blocking lbc_net_run_host call"},
            "gpu_launches": int(args.steps * net.launches),
            "roofline": roofline,
            "output_crc32": job.checksums,
        }
        if world == 1 and not args.no_cpu_baseline:
            try:
                out["cpu_baseline"] = cpu_baseline_port(layers)
            except Exception as e:  # noqa: BLE001
                out["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": f"failed: {e}"}
        if args.layer_report:
            rep = []
            for i, (name, d, _) in enumerate(layers):
                ops, byts = works[i]
                ms = float(per_layer[i])
                into = net.fused_into(i)
                if into >= 0:      # no launch of its own: its work is in the consumer's (fused) line
                    rep.append({"layer": name, "kernel": kinds[i], "plan": f"fused into {layers[into][0]}", "ms": 0.0, "tops": 0.0,
                                "gbs": 0.0, "ai": ops / byts, "frac_tc": 0.0, "frac_hbm": 0.0, "fused_into": layers[into][0]})
                    continue
                src = [j for j in range(len(layers)) if net.fused_into(j) == i]
                if src:            # a fused pair: both convolutions' algorithmic work over the one launch
                    ops, byts = ops + works[src[0]][0], byts + works[src[0]][1]
                rep.append({"layer": name + (f" (+{layers[src[0]][0]})" if src else ""), "kernel": kinds[i], "plan": net.layer_describe(i), "ms": ms,
                            "tops": ops / ms / 1e9, "gbs": byts / ms / 1e6, "ai": ops / byts,
                            "frac_tc": ops / ms / 1e9 / (tc_peak / 1e12), "frac_hbm": byts / ms / 1e6 / peaks["hbm_gbs"]})
            with open(args.layer_report, "w") as fh:
                json.dump({"network": args.network, "batch": args.batch, "int8_peak_tops": int8_peak,
                           "hbm_gbs": peaks["hbm_gbs"], "layers": rep}, fh, indent=1)
        emit_json(out)
    net.close()
    if world > 1:
        dist.destroy_process_group()
    if job.bad_layers:
        sys.exit(3)


def _count(net):
    c = {}
    for i in range(len(net)):
        k = net.layer_kernel(i)
        c[k] = c.get(k, 0) + 1
    return c


if __name__ == "__main__":
    main()
