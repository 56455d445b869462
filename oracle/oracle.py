"""
oracle.py — Python face of the CPU oracle.  TEST INFRASTRUCTURE ONLY.

Importable only from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs (see oracle/cpu_ref.c for the full statement).  Three checkers live here:

  * ``ref_conv2d_forward``   the REFERENCE's own refConv2DForward (cpp/int8conv/refConv2DForward.hpp:56-80),
                             compiled unmodified into oracle/_ref/libref_conv.so by oracle/Makefile.
  * ``conv_nhwc`` & co.      our C restatement (oracle/cpu_ref.c), pinned against the above.
  * ``np_conv_nhwc``         a numpy restatement of the same arithmetic (independent code path used to
                             cross-check the C restatement, small shapes only).
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_CPU_SO = os.path.join(_HERE, "_build", "libcpu_ref.so")
_REF_SO = os.path.join(_HERE, "_ref", "libref_conv.so")


class _Desc(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in (
        "n", "h", "w", "c", "k", "r", "s", "stride_h", "stride_w", "pad_h", "pad_w",
        "dil_h", "dil_w", "groups", "relu", "out_mode")]


@dataclass(frozen=True)
class ConvDesc:
    """Mirror of lbc_conv_desc (include/lowbit_cnn.h)."""
    n: int
    h: int
    w: int
    c: int
    k: int
    r: int
    s: int
    stride_h: int = 1
    stride_w: int = 1
    pad_h: int = 0
    pad_w: int = 0
    dil_h: int = 1
    dil_w: int = 1
    groups: int = 1
    relu: int = 0
    out_mode: int = 0  # 0 int8, 1 int32

    @property
    def p(self) -> int:
        return out_dim(self.h, self.pad_h, self.dil_h, self.r, self.stride_h)

    @property
    def q(self) -> int:
        return out_dim(self.w, self.pad_w, self.dil_w, self.s, self.stride_w)

    @property
    def macs(self) -> int:
        return self.n * self.p * self.q * self.k * (self.c // self.groups) * self.r * self.s

    @property
    def bytes(self) -> int:
        """Algorithmic bytes, SURVEY.md 8d."""
        out_elt = 1 if self.out_mode == 0 else 4
        return (self.n * self.h * self.w * self.c + self.k * (self.c // self.groups) * self.r * self.s
                + self.n * self.p * self.q * self.k * out_elt + 8 * self.k)

    def as_struct(self) -> _Desc:
        return _Desc(self.n, self.h, self.w, self.c, self.k, self.r, self.s, self.stride_h, self.stride_w,
                     self.pad_h, self.pad_w, self.dil_h, self.dil_w, self.groups, self.relu, self.out_mode)


def out_dim(i: int, pad: int, dil: int, k: int, stride: int) -> int:
    """cpp/int8conv/cudnn2DConvolution.cuh:33-36."""
    return (i + 2 * pad - (dil * (k - 1) + 1)) // stride + 1


# ---------------------------------------------------------------------------------------------------
# C restatement
# ---------------------------------------------------------------------------------------------------
_cpu = None


def build_cpu_ref(force: bool = False) -> str:
    """Compile oracle/cpu_ref.c (gcc only). Returns the .so path."""
    if force or not os.path.exists(_CPU_SO) or os.path.getmtime(_CPU_SO) < os.path.getmtime(
            os.path.join(_HERE, "cpu_ref.c")):
        subprocess.check_call(["make", "-s", "-C", _HERE, "oracle"])
    return _CPU_SO


def _cpu_lib():
    global _cpu
    if _cpu is None:
        build_cpu_ref()
        lib = ctypes.CDLL(_CPU_SO)
        lib.oracle_conv_nhwc.restype = ctypes.c_int
        lib.oracle_conv_nhwc.argtypes = [ctypes.POINTER(_Desc)] + [ctypes.c_void_p] * 5 + [ctypes.c_int]
        lib.oracle_ref_conv_nchw_valid.restype = None
        lib.oracle_ref_conv_nchw_valid.argtypes = [ctypes.c_int32] * 9 + [ctypes.c_void_p] * 3
        lib.oracle_requant.restype = ctypes.c_int8
        lib.oracle_requant.argtypes = [ctypes.c_int32, ctypes.c_int32, ctypes.c_float, ctypes.c_int]
        lib.oracle_max_threads.restype = ctypes.c_int
        for f in (lib.oracle_to_vect_c, lib.oracle_from_vect_c):
            f.restype = None
            f.argtypes = [ctypes.c_void_p, ctypes.c_void_p] + [ctypes.c_int32] * 6
        _cpu = lib
    return _cpu


def max_threads() -> int:
    return int(_cpu_lib().oracle_max_threads())


def _ptr(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def conv_nhwc(d: ConvDesc, x: np.ndarray, w_krsc: np.ndarray, bias, scale, threads: int = 0) -> np.ndarray:
    """General oracle: x int8 [N,H,W,C], w int8 [K,R,S,C/g], bias int32[K]|None, scale f32[K]|None."""
    x = np.ascontiguousarray(x, dtype=np.int8)
    w = np.ascontiguousarray(w_krsc, dtype=np.int8)
    assert x.shape == (d.n, d.h, d.w, d.c), (x.shape, d)
    assert w.shape == (d.k, d.r, d.s, d.c // d.groups), (w.shape, d)
    b = None if bias is None else np.ascontiguousarray(bias, dtype=np.int32)
    sc = None if scale is None else np.ascontiguousarray(scale, dtype=np.float32)
    y = np.empty((d.n, d.p, d.q, d.k), dtype=np.int8 if d.out_mode == 0 else np.int32)
    st = d.as_struct()
    rc = _cpu_lib().oracle_conv_nhwc(ctypes.byref(st), _ptr(x), _ptr(w), _ptr(b), _ptr(sc), _ptr(y), threads)
    if rc != 0:
        raise ValueError(f"oracle_conv_nhwc rejected {d}")
    return y


def ref_style_nchw_valid(x_nchw: np.ndarray, w_oihw: np.ndarray) -> np.ndarray:
    """C restatement of refConv2DForwardImpl (pre-padded NCHW in, OIHW kernel, int32 NCHW out)."""
    x = np.ascontiguousarray(x_nchw, dtype=np.int8)
    w = np.ascontiguousarray(w_oihw, dtype=np.int8)
    b, ic, ih, iw = x.shape
    oc, ic2, kh, kw = w.shape
    assert ic == ic2
    oh, ow = ih - kh + 1, iw - kw + 1
    y = np.empty((b, oc, oh, ow), dtype=np.int32)
    _cpu_lib().oracle_ref_conv_nchw_valid(b, ic, ih, iw, oc, oh, ow, kh, kw, _ptr(x), _ptr(w), _ptr(y))
    return y


def requant(acc: int, bias: int, scale: float, relu: bool) -> int:
    return int(_cpu_lib().oracle_requant(int(acc), int(bias), float(scale), int(relu)))


def to_vect_c(a: np.ndarray, v: int = 16) -> np.ndarray:
    a = np.ascontiguousarray(a)
    n, c, h, w = a.shape
    out = np.empty((n, c // v, h, w, v), dtype=a.dtype)
    _cpu_lib().oracle_to_vect_c(_ptr(a), _ptr(out), n, c, h, w, v, a.dtype.itemsize)
    return out


def from_vect_c(a: np.ndarray) -> np.ndarray:
    a = np.ascontiguousarray(a)
    n, cg, h, w, v = a.shape
    out = np.empty((n, cg * v, h, w), dtype=a.dtype)
    _cpu_lib().oracle_from_vect_c(_ptr(a), _ptr(out), n, cg * v, h, w, v, a.dtype.itemsize)
    return out


# ---------------------------------------------------------------------------------------------------
# The reference itself (oracle/_ref)
# ---------------------------------------------------------------------------------------------------
_ref = None


def have_ref() -> bool:
    return os.path.exists(_REF_SO)


def build_ref() -> str:
    """Compile the reference header from /root/reference (authoring container only)."""
    if not os.path.exists("/root/reference/cpp/int8conv/refConv2DForward.hpp"):
        raise FileNotFoundError("/root/reference is not present; oracle/_ref cannot be rebuilt here")
    subprocess.check_call(["make", "-s", "-C", _HERE, "ref"])
    return _REF_SO


def _ref_lib():
    global _ref
    if _ref is None:
        import torch  # noqa: F401  (libtorch must be resident before the reference TU is loaded)
        lib = ctypes.CDLL(_REF_SO)
        lib.ref_conv2d_forward.restype = ctypes.c_int
        lib.ref_conv2d_forward.argtypes = [ctypes.c_int] * 9 + [ctypes.c_void_p] * 3
        lib.ref_num_shapes.restype = ctypes.c_int
        lib.ref_shape.restype = ctypes.c_int
        lib.ref_shape.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_int * 9)]
        lib.ref_max_threads.restype = ctypes.c_int
        _ref = lib
    return _ref


def ref_shapes() -> list[tuple[int, ...]]:
    lib = _ref_lib()
    out = []
    for i in range(lib.ref_num_shapes()):
        buf = (ctypes.c_int * 9)()
        assert lib.ref_shape(i, ctypes.byref(buf)) == 0
        out.append(tuple(buf))
    return out


def ref_max_threads() -> int:
    return int(_ref_lib().ref_max_threads())


def ref_conv2d_forward(x_nchw: np.ndarray, w_oihw: np.ndarray) -> np.ndarray:
    """The reference's refConv2DForward on NCHW int8 / OIHW int8 -> int32 NCHW (VALID, stride 1)."""
    x = np.ascontiguousarray(x_nchw, dtype=np.int8)
    w = np.ascontiguousarray(w_oihw, dtype=np.int8)
    b, ic, ih, iw = x.shape
    oc, _, kh, kw = w.shape
    oh, ow = ih - kh + 1, iw - kw + 1
    y = np.empty((b, oc, oh, ow), dtype=np.int32)
    rc = _ref_lib().ref_conv2d_forward(b, ic, ih, iw, oc, oh, ow, kh, kw, _ptr(x), _ptr(w), _ptr(y))
    if rc == 1:
        raise KeyError(f"shape {(b, ic, ih, iw, oc, oh, ow, kh, kw)} is not instantiated in ref_wrap.cpp")
    if rc != 0:
        raise RuntimeError("reference raised")
    return y


# ---------------------------------------------------------------------------------------------------
# numpy restatement (independent of the C code)
# ---------------------------------------------------------------------------------------------------
def np_requant(t: np.ndarray, scale: np.ndarray, relu: bool) -> np.ndarray:
    """t int32 [..., K] (acc+bias), scale f32[K] -> int8.  The reference's quantize(): round half-to-even FIRST
    (__float2int_rn: NaN -> 0), then clamp to [relu ? 0 : -128, 127] (conv2DForward3x3WinogradFused.cuh:39-46;
    quantization.py:27-49 clamps then rounds, identical for integer bounds and finite values)."""
    with np.errstate(invalid="ignore", over="ignore"):
        f = t.astype(np.float32) * scale.astype(np.float32)  # one fp32 multiply
    f = np.where(np.isnan(f), np.float32(0.0), f)            # __float2int_rn(NaN) == 0
    q = np.rint(np.clip(f, np.float32(-1e9), np.float32(1e9)))   # np.rint = half-to-even; +-inf end up at the clamps
    return np.clip(q, 0.0 if relu else -128.0, 127.0).astype(np.int8)


def np_conv_nhwc(d: ConvDesc, x: np.ndarray, w_krsc: np.ndarray, bias, scale) -> np.ndarray:
    """Small shapes only. int64 accumulate then wrap to int32 (same result as int32 wraparound)."""
    p, q = d.p, d.q
    cg, kg = d.c // d.groups, d.k // d.groups
    xp = np.zeros((d.n, d.h + 2 * d.pad_h, d.w + 2 * d.pad_w, d.c), dtype=np.int64)
    xp[:, d.pad_h:d.pad_h + d.h, d.pad_w:d.pad_w + d.w, :] = x
    acc = np.zeros((d.n, p, q, d.k), dtype=np.int64)
    w64 = w_krsc.astype(np.int64)
    for r in range(d.r):
        for s in range(d.s):
            h0, w0 = r * d.dil_h, s * d.dil_w
            patch = xp[:, h0:h0 + (p - 1) * d.stride_h + 1:d.stride_h,
                       w0:w0 + (q - 1) * d.stride_w + 1:d.stride_w, :]
            for g in range(d.groups):
                acc[..., g * kg:(g + 1) * kg] += np.einsum(
                    "npqc,kc->npqk", patch[..., g * cg:(g + 1) * cg], w64[g * kg:(g + 1) * kg, r, s, :])
    if bias is not None:
        acc = acc + bias.astype(np.int64)
    t = ((acc + 2**31) % 2**32 - 2**31).astype(np.int32)  # int32 wraparound
    if d.out_mode == 1:
        return t
    return np_requant(t, scale, bool(d.relu))


# ---------------------------------------------------------------------------------------------------
# Synthetic inputs: identical bytes for oracle and GPU.
# ---------------------------------------------------------------------------------------------------
def synth(d: ConvDesc, layer: int = 0, style: str = "full"):
    """Returns (x NHWC int8, w KRSC int8, bias int32[K], scale f32[K]).
    style 'full': activations U[-128,127], weights U[-127,127]; 'ref': values in {0,1} (check.cu:43-44,69-75)."""
    cg = d.c // d.groups
    rx = np.random.default_rng(1234 + layer)
    rw = np.random.default_rng(4321 + layer)
    if style == "ref":
        x = rx.integers(0, 2, size=(d.n, d.h, d.w, d.c), dtype=np.int8)
        w = rw.integers(0, 2, size=(d.k, d.r, d.s, cg), dtype=np.int8)
    else:
        x = rx.integers(-128, 128, size=(d.n, d.h, d.w, d.c), dtype=np.int8)
        w = rw.integers(-127, 128, size=(d.k, d.r, d.s, cg), dtype=np.int8)
    bias = rw.integers(-2**15, 2**15, size=(d.k,), dtype=np.int32)
    scale = (rw.uniform(0.5, 2.0, size=(d.k,)) * 2.0**-7 / np.sqrt(d.r * d.s * cg)).astype(np.float32)
    return x, w, bias, scale


# ---------------------------------------------------------------------------------------------------
# int8 ops between convolutions (restated in numpy; integer arithmetic, so there is one right answer)
# ---------------------------------------------------------------------------------------------------
def pool_out_dim(i: int, pad: int, k: int, stride: int) -> int:
    """cudnnGetPooling2dForwardOutputDim as the reference uses it (python/qtorch/cpp/pool2d.cuh:76-78)."""
    return 1 + (i + 2 * pad - k) // stride


def max_pool_nhwc(x: np.ndarray, kh: int, kw: int, sh: int, sw: int, ph: int, pw: int) -> np.ndarray:
    """int8 NHWC max-pool, padding never wins (cuDNN CUDNN_POOLING_MAX_DETERMINISTIC, pool2d.cuh:40-43)."""
    n, h, w, c = x.shape
    p, q = pool_out_dim(h, ph, kh, sh), pool_out_dim(w, pw, kw, sw)
    xp = np.full((n, h + 2 * ph + sh, w + 2 * pw + sw, c), -128, dtype=np.int8)
    xp[:, ph:ph + h, pw:pw + w, :] = x
    out = np.full((n, p, q, c), -128, dtype=np.int8)
    for a in range(kh):
        for b in range(kw):
            out = np.maximum(out, xp[:, a:a + (p - 1) * sh + 1:sh, b:b + (q - 1) * sw + 1:sw, :])
    return out


def add_relu(a: np.ndarray, b: np.ndarray, relu: bool) -> np.ndarray:
    """Residual join: exact int sum, saturated to int8 (the quantizer's saturation: WinogradFused.cuh:39-46), ReLU."""
    s = a.astype(np.int32) + b.astype(np.int32)
    return np.clip(s, 0 if relu else -128, 127).astype(np.int8)


def global_avg_pool(x: np.ndarray, scale: float) -> np.ndarray:
    """[N,H,W,C] int8 -> [N,C] int8 through the convolutions' requantisation rule with one scale."""
    t = x.astype(np.int32).sum(axis=(1, 2), dtype=np.int32)
    return np_requant(t, np.full((x.shape[3],), scale, dtype=np.float32), relu=False)
