"""CPU tests: the oracle against the reference's outputs (golden + live), and against itself."""
import os
import zlib

import numpy as np
import pytest

from oracle import oracle
from oracle.oracle import ConvDesc

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "ref_conv_golden.npz")


def _golden_cases():
    z = np.load(GOLDEN)
    n = len([k for k in z.files if k.endswith("_y")])
    return z, n


def _golden_inputs(shape, style, seed):
    # must match tests/golden/make_golden.py::inputs
    b, ic, ih, iw, oc, oh, ow, kh, kw = shape
    rng = np.random.default_rng(seed)
    if style == "ref":
        x = rng.integers(0, 2, size=(b, ic, ih, iw), dtype=np.int8)
        w = rng.integers(0, 2, size=(oc, ic, kh, kw), dtype=np.int8)
    elif style == "extreme":
        x = rng.choice(np.array([-128, 127], dtype=np.int8), size=(b, ic, ih, iw))
        w = rng.choice(np.array([-128, 127], dtype=np.int8), size=(oc, ic, kh, kw))
    else:
        x = rng.integers(-128, 128, size=(b, ic, ih, iw), dtype=np.int8)
        w = rng.integers(-128, 128, size=(oc, ic, kh, kw), dtype=np.int8)
    return x, w


def iter_golden():
    z, n = _golden_cases()
    for i in range(n):
        key = f"case{i:02d}"
        shape = tuple(int(v) for v in z[key + "_shape"])
        style = str(z[key + "_style"])
        seed = int(z[key + "_seed"])
        x, w = _golden_inputs(shape, style, seed)
        assert zlib.crc32(x.tobytes() + w.tobytes()) == int(z[key + "_crc"]), "golden input regeneration drifted"
        yield key, shape, x, w, z[key + "_y"]


def test_golden_file_has_cases():
    _, n = _golden_cases()
    assert n >= 20


def test_c_restatement_matches_reference_golden():
    """oracle_ref_conv_nchw_valid == refConv2DForward outputs, bit for bit, every golden case."""
    for key, shape, x, w, y in iter_golden():
        got = oracle.ref_style_nchw_valid(x, w)
        assert got.dtype == np.int32 and got.shape == y.shape
        assert np.array_equal(got, y), key


def test_general_oracle_matches_reference_golden():
    """oracle_conv_nhwc (NHWC/KRSC, int32 out) == the reference after layout transposition."""
    for key, shape, x, w, y in iter_golden():
        b, ic, ih, iw, oc, oh, ow, kh, kw = shape
        d = ConvDesc(n=b, h=ih, w=iw, c=ic, k=oc, r=kh, s=kw, out_mode=1)
        got = oracle.conv_nhwc(d, x.transpose(0, 2, 3, 1), w.transpose(0, 2, 3, 1), None, None)
        assert np.array_equal(got.transpose(0, 3, 1, 2), y), key


def test_numpy_restatement_matches_reference_golden():
    for key, shape, x, w, y in list(iter_golden())[::3]:
        b, ic, ih, iw, oc, oh, ow, kh, kw = shape
        d = ConvDesc(n=b, h=ih, w=iw, c=ic, k=oc, r=kh, s=kw, out_mode=1)
        got = oracle.np_conv_nhwc(d, x.transpose(0, 2, 3, 1), w.transpose(0, 2, 3, 1), None, None)
        assert np.array_equal(got.transpose(0, 3, 1, 2), y), key


@pytest.mark.skipif(not oracle.have_ref(), reason="oracle/_ref not built (needs /root/reference)")
def test_live_reference_agrees_with_restatement():
    """Run the compiled reference itself on fresh draws of its two cheapest shapes."""
    for shape in [(1, 8, 11, 11, 4, 5, 5, 7, 7), (1, 64, 6, 6, 32, 6, 6, 1, 1)]:
        for seed in (1, 2):
            x, w = _golden_inputs(shape, "full", seed)
            assert np.array_equal(oracle.ref_conv2d_forward(x, w), oracle.ref_style_nchw_valid(x, w))


GENERAL = [
    ConvDesc(n=2, h=9, w=7, c=8, k=12, r=3, s=3, pad_h=1, pad_w=1, relu=1),
    ConvDesc(n=1, h=12, w=12, c=16, k=8, r=3, s=3, stride_h=2, stride_w=2, pad_h=1, pad_w=1),
    ConvDesc(n=2, h=8, w=8, c=3, k=16, r=7, s=7, stride_h=2, stride_w=2, pad_h=3, pad_w=3, relu=1),
    ConvDesc(n=1, h=6, w=5, c=32, k=24, r=1, s=1, relu=1),
    ConvDesc(n=1, h=7, w=7, c=16, k=16, r=1, s=1, stride_h=2, stride_w=2),
    ConvDesc(n=2, h=9, w=9, c=24, k=24, r=3, s=3, pad_h=1, pad_w=1, groups=24, relu=1),      # depthwise
    ConvDesc(n=1, h=9, w=9, c=24, k=24, r=3, s=3, stride_h=2, stride_w=2, pad_h=1, pad_w=1, groups=24),
    ConvDesc(n=1, h=10, w=10, c=8, k=8, r=3, s=3, pad_h=2, pad_w=2, dil_h=2, dil_w=2, groups=2, relu=1),
    ConvDesc(n=1, h=5, w=9, c=4, k=6, r=3, s=2, pad_h=0, pad_w=1, stride_h=1, stride_w=2),
]


@pytest.mark.parametrize("d", GENERAL, ids=lambda d: f"c{d.c}k{d.k}r{d.r}s{d.stride_h}g{d.groups}")
@pytest.mark.parametrize("out_mode", [0, 1])
def test_c_vs_numpy_general(d, out_mode):
    d = ConvDesc(**{**d.__dict__, "out_mode": out_mode})
    x, w, bias, scale = oracle.synth(d, layer=3)
    a = oracle.conv_nhwc(d, x, w, bias, scale)
    b = oracle.np_conv_nhwc(d, x, w, bias, scale)
    assert a.dtype == b.dtype
    assert np.array_equal(a, b)
    if out_mode == 0:
        assert len(np.unique(a)) > 8, "synthetic scale should spread outputs over the int8 range"


def test_requant_rounding_rule():
    """Round-half-to-even then saturate (WinogradFused.cuh:39-46 / quantization.py:27-49)."""
    cases = [
        (1, 0, 0.5, False, 0),      # 0.5 -> 0 (even)
        (3, 0, 0.5, False, 2),      # 1.5 -> 2
        (5, 0, 0.5, False, 2),      # 2.5 -> 2
        (-1, 0, 0.5, False, 0),     # -0.5 -> -0
        (-3, 0, 0.5, False, -2),    # -1.5 -> -2
        (-5, 0, 0.5, False, -2),    # -2.5 -> -2
        (1000, 0, 1.0, False, 127),
        (-1000, 0, 1.0, False, -128),
        (-1000, 0, 1.0, True, 0),
        (-3, 0, 0.5, True, 0),
        (253, 0, 0.5, False, 126),  # 126.5 -> 126
        (255, 0, 0.5, False, 127),  # 127.5 -> clamp 127
        (-255, 0, 0.5, False, -128),  # -127.5 -> -128 (even) / clamp
        (10, -7, 1.0, False, 3),
        (2**31 - 1, 1, 1.0, False, -128),  # int32 wraparound of acc+bias
        (7, 0, float("nan"), False, 0),     # __float2int_rn(NaN) == 0, then the clamp (WinogradFused.cuh:39-46)
        (7, 0, float("nan"), True, 0),
        (7, 0, float("inf"), False, 127),
        (7, 0, -float("inf"), False, -128),
        (16777217, 0, 2.0**-17, False, 127),  # (float)t rounds 2^24+1 -> 2^24 ; 128.0 -> clamp
    ]
    for acc, bias, scale, relu, want in cases:
        assert oracle.requant(acc, bias, scale, relu) == want, (acc, bias, scale, relu)
        t = np.array([(acc + bias + 2**31) % 2**32 - 2**31], dtype=np.int32)
        assert int(oracle.np_requant(t, np.array([scale], dtype=np.float32), relu)[0]) == want


def test_vect_c_roundtrip_matches_reference_definition():
    """utils.cuh:11-26: reshape [N,C/V,V,H,W] + permute(0,1,3,4,2)."""
    rng = np.random.default_rng(0)
    for v, dt in ((16, np.int8), (4, np.int8), (32, np.int8), (16, np.int32)):
        a = rng.integers(-100, 100, size=(2, 2 * v, 3, 5)).astype(dt)
        want = a.reshape(2, 2, v, 3, 5).transpose(0, 1, 3, 4, 2)
        got = oracle.to_vect_c(a, v)
        assert np.array_equal(got, want)
        assert np.array_equal(oracle.from_vect_c(got), a)


def test_out_dim_formula():
    assert oracle.out_dim(224, 3, 1, 7, 2) == 112
    assert oracle.out_dim(56, 1, 1, 3, 1) == 56
    assert oracle.out_dim(56, 0, 1, 1, 2) == 28
    assert oracle.out_dim(130, 0, 1, 3, 1) == 128  # check.cu:31-41


def test_pool_and_add_restatements_against_torch_cpu():
    """The numpy max-pool / residual-add restatements against an independent implementation (torch on the CPU, fp32 is
    exact for int8 values): window, stride and padding combinations incl. the reference's own (python/tmp.py:43-56)."""
    import torch
    rng = np.random.default_rng(21)
    for (h, w, c, k, s, p) in [(12, 12, 8, 2, 2, 0), (9, 11, 5, 3, 1, 0), (14, 14, 16, 3, 2, 1), (7, 7, 4, (3, 2), (2, 1), (1, 0))]:
        pair = lambda v: (v, v) if isinstance(v, int) else v
        (kh, kw), (sh, sw), (ph, pw) = pair(k), pair(s), pair(p)
        x = rng.integers(-128, 128, size=(2, h, w, c), dtype=np.int8)
        want = torch.nn.functional.max_pool2d(torch.from_numpy(x).permute(0, 3, 1, 2).float(), (kh, kw), (sh, sw), (ph, pw))
        got = oracle.max_pool_nhwc(x, kh, kw, sh, sw, ph, pw)
        assert np.array_equal(got, want.permute(0, 2, 3, 1).numpy().astype(np.int8))
    a = rng.integers(-128, 128, size=(1000,), dtype=np.int8)
    b = rng.integers(-128, 128, size=(1000,), dtype=np.int8)
    for relu in (False, True):
        want = [min(127, max(0 if relu else -128, int(u) + int(v))) for u, v in zip(a, b)]
        assert oracle.add_relu(a, b, relu).tolist() == want
    x = rng.integers(-128, 128, size=(2, 3, 5, 7), dtype=np.int8)
    want = np.rint(np.clip(x.astype(np.int64).sum(axis=(1, 2)).astype(np.float32) * np.float32(0.07), -128, 127)).astype(np.int8)
    assert np.array_equal(oracle.global_avg_pool(x, 0.07), want)
