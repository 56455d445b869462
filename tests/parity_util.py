"""Shared helpers for the GPU parity tests: run a convolution through the C ABI and through the oracle."""
from __future__ import annotations

import numpy as np

from oracle import oracle
from oracle.oracle import ConvDesc as ODesc


def lbc_desc(od: ODesc):
    import lowbitdnn_project_b200 as lbc
    return lbc.ConvDesc(**od.__dict__)


def run_gpu(od: ODesc, x, w, bias, scale, force=0, w_layout="krsc", options=None):
    """Returns (y numpy, kernel name, ms).  x NHWC int8, w KRSC int8 numpy arrays.
    options: lbc_plan_options fields (dict) forcing planner decisions."""
    import torch
    import lowbitdnn_project_b200 as lbc
    dev = torch.device("cuda:0")
    plan = lbc.ConvPlan(lbc_desc(od), force=force, options=options)
    if w_layout == "oihw":
        wt = torch.from_numpy(np.ascontiguousarray(w.transpose(0, 3, 1, 2))).to(dev)
        wp = plan.prepack(wt.reshape(-1), lbc.W_OIHW)
    else:
        wp = plan.prepack(torch.from_numpy(w).to(dev).reshape(-1), lbc.W_KRSC)
    xt = torch.from_numpy(x).to(dev)
    bt = None if bias is None else torch.from_numpy(bias).to(dev)
    st = None if scale is None else torch.from_numpy(scale).to(dev)
    y, ms = plan.run(xt, wp, bt, st, timed=True)
    torch.cuda.synchronize()
    out = y.cpu().numpy()
    name = plan.kernel
    plan.close()
    return out, name, ms


def check_case(od: ODesc, layer=0, style="full", force=0, use_bias=True, w_layout="krsc", bias_range=None, options=None):
    """Runs GPU and oracle on identical seeded inputs; returns (mismatches, total, kernel, detail).
    bias_range: replace the synthetic biases by uniform int32 values in [-bias_range, bias_range]."""
    x, w, bias, scale = oracle.synth(od, layer=layer, style=style)
    if bias_range is not None:
        bias = np.random.default_rng(99 + layer).integers(-bias_range, bias_range + 1, size=bias.shape).astype(np.int32)
    if not use_bias:
        bias = None
    want = oracle.conv_nhwc(od, x, w, bias, scale)
    got, name, ms = run_gpu(od, x, w, bias, scale if od.out_mode == 0 else None, force=force, w_layout=w_layout, options=options)
    assert got.shape == want.shape and got.dtype == want.dtype, (got.shape, want.shape, got.dtype, want.dtype)
    bad = got != want
    nbad = int(bad.sum())
    detail = ""
    if nbad:
        detail = mismatch_report(got, want)
    return nbad, int(want.size), name, detail


def mismatch_report(got, want) -> str:
    """Structure of a mismatch (which output rows / channels are wrong) to localise descriptor bugs."""
    bad = got != want
    k = got.shape[-1]
    b2 = bad.reshape(-1, k)
    rows = np.nonzero(b2.any(axis=1))[0]
    cols = np.nonzero(b2.any(axis=0))[0]
    g2, w2 = got.reshape(-1, k), want.reshape(-1, k)
    lines = [f"mismatched {int(bad.sum())}/{bad.size}; rows {len(rows)}/{b2.shape[0]} cols {len(cols)}/{k}",
             f"first bad rows {rows[:12].tolist()} ; first bad cols {cols[:12].tolist()}"]
    for mod in (8, 32, 128):
        lines.append(f"bad rows mod {mod}: {np.bincount(rows % mod, minlength=mod).tolist() if len(rows) else []}")
    lines.append(f"bad cols mod 16: {np.bincount(cols % 16, minlength=16).tolist() if len(cols) else []}")
    if len(rows):
        r = rows[0]
        lines.append(f"row {r}: got {g2[r, :8].tolist()} want {w2[r, :8].tolist()}")
        lines.append(f"all-zero rows in got: {int((g2 == 0).all(axis=1).sum())}")
    return "\n".join(lines)
