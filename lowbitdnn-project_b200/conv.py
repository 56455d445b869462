"""Host-side wrappers over the C ABI (plans, weight pre-pack, run, layout converters).

PyTorch is used here only as the device allocator / stream provider: every tensor argument is a CUDA
torch tensor whose ``data_ptr()`` is handed to liblowbit-cnn.  All arithmetic happens inside the library.

The call surface mirrors the reference operators (names are the reference's, argument meaning too):
  conv2DForward3x3(input_vect_c, kernel_vect_c) -> (out_int32_vect_c, ms)
        cpp/int8conv/conv2DForward3x3TensorCores.cuh:695-751
  to_vect_c / from_vect_c                         cpp/int8conv/utils.cuh:11-26
"""
from __future__ import annotations

import ctypes
from dataclasses import asdict, dataclass

from . import _capi
from ._capi import CConvDesc, check, load_library


@dataclass(frozen=True)
class ConvDesc:
    """lbc_conv_desc (include/lowbit_cnn.h)."""
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
    out_mode: int = _capi.OUT_INT8

    def c_struct(self) -> CConvDesc:
        return CConvDesc(**asdict(self))

    def replace(self, **kw) -> "ConvDesc":
        d = asdict(self)
        d.update(kw)
        return ConvDesc(**d)

    @property
    def out_hw(self) -> tuple[int, int]:
        p, q = ctypes.c_int32(), ctypes.c_int32()
        st = self.c_struct()
        check(load_library().lbc_conv_out_shape(ctypes.byref(st), ctypes.byref(p), ctypes.byref(q)))
        return p.value, q.value

    @property
    def work(self) -> tuple[float, float]:
        """(ops, algorithmic bytes) — SURVEY.md 8d formulas, computed by the library."""
        ops, byts = ctypes.c_double(), ctypes.c_double()
        st = self.c_struct()
        check(load_library().lbc_conv_work(ctypes.byref(st), ctypes.byref(ops), ctypes.byref(byts)))
        return ops.value, byts.value


def _stream_ptr(stream) -> ctypes.c_void_p:
    import torch
    s = stream if stream is not None else torch.cuda.current_stream()
    return ctypes.c_void_p(s.cuda_stream)


def _ptr(t) -> ctypes.c_void_p:
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


class ConvPlan:
    """An lbc_plan plus convenience methods on torch CUDA tensors."""

    def __init__(self, desc: ConvDesc, force: int = _capi.KERNEL_AUTO, options: dict | None = None):
        """options: lbc_plan_options fields by name (tests / tuning), e.g. {"cta_pairs": 1, "max_grid": 3}."""
        self.desc = desc
        self._lib = load_library()
        self._h = ctypes.c_void_p()
        st = desc.c_struct()
        if options:
            opt = _capi.plan_options(**options)
            check(self._lib.lbc_conv_plan_create_ex(ctypes.byref(st), force, ctypes.byref(opt), ctypes.byref(self._h)))
        else:
            check(self._lib.lbc_conv_plan_create(ctypes.byref(st), force, ctypes.byref(self._h)))
        self.p, self.q = desc.out_hw

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._lib.lbc_conv_plan_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def kernel(self) -> str:
        k = ctypes.c_int32()
        check(self._lib.lbc_conv_plan_kernel(self._h, ctypes.byref(k)))
        return _capi.KERNEL_NAMES[k.value]

    def describe(self) -> str:
        buf = ctypes.create_string_buffer(512)
        check(self._lib.lbc_conv_plan_describe(self._h, buf, 512))
        return buf.value.decode()

    @property
    def launches(self) -> int:
        n = ctypes.c_int32()
        check(self._lib.lbc_conv_plan_launches(self._h, ctypes.byref(n)))
        return n.value

    def set_trace(self, buf, tiles: int) -> None:
        """Development aid: int64 CUDA tensor receiving CTA 0's pipeline time stamps (None switches it off)."""
        check(self._lib.lbc_conv_plan_set_trace(self._h, _ptr(buf), tiles))

    def check_status(self) -> None:
        """Raises LbcError(LBC_ERR_KERNEL_TIMEOUT) if a launch of this plan tripped its device watchdog (sync first)."""
        check(self._lib.lbc_conv_plan_check(self._h))

    @property
    def packed_weight_bytes(self) -> int:
        n = ctypes.c_size_t()
        check(self._lib.lbc_conv_packed_weight_bytes(self._h, ctypes.byref(n)))
        return n.value

    def prepack(self, w, layout: int = _capi.W_KRSC, stream=None):
        """w: int8 CUDA tensor in KRSC or OIHW order -> packed int8 CUDA tensor for this plan's kernel."""
        import torch
        assert w.is_cuda and w.dtype == torch.int8 and w.is_contiguous()
        d = self.desc
        assert w.numel() == d.k * d.r * d.s * (d.c // d.groups), "weight element count does not match the descriptor"
        out = torch.empty(self.packed_weight_bytes, dtype=torch.int8, device=w.device)
        check(self._lib.lbc_conv_prepack_weights(self._h, _ptr(w), layout, _ptr(out), _stream_ptr(stream)))
        return out

    def empty_output(self, device):
        import torch
        d = self.desc
        dt = torch.int8 if d.out_mode == _capi.OUT_INT8 else torch.int32
        return torch.empty((d.n, self.p, self.q, d.k), dtype=dt, device=device)

    def run(self, x, w_packed, bias=None, scale=None, out=None, stream=None, timed: bool = False):
        """x: int8 NHWC CUDA tensor.  Returns y, or (y, ms) when timed (the reference's (Tensor, float))."""
        import torch
        d = self.desc
        assert x.is_cuda and x.dtype == torch.int8 and x.is_contiguous() and tuple(x.shape) == (d.n, d.h, d.w, d.c)
        assert bias is None or (bias.dtype == torch.int32 and bias.numel() == d.k and bias.is_cuda)
        assert scale is None or (scale.dtype == torch.float32 and scale.numel() == d.k and scale.is_cuda)
        y = out if out is not None else self.empty_output(x.device)
        ms = ctypes.c_float()
        check(self._lib.lbc_conv_run(self._h, _ptr(x), _ptr(w_packed), _ptr(bias), _ptr(scale), _ptr(y),
                                     _stream_ptr(stream), ctypes.byref(ms) if timed else None))
        return (y, ms.value) if timed else y

    def run_host(self, x_host, w_packed, bias, scale, y_host, stream=None) -> float:
        """x_host / y_host: CPU (ideally pinned) tensors; H2D + kernel + D2H inside the call. Returns ms."""
        ms = ctypes.c_float()
        check(self._lib.lbc_conv_run_host(self._h, _ptr(x_host), _ptr(w_packed), _ptr(bias), _ptr(scale),
                                          _ptr(y_host), _stream_ptr(stream), ctypes.byref(ms)))
        return ms.value


# ---- layout converters (reference tensor formats) --------------------------------------------------
def _convert(fn_name: str, src, n, c, h, w, v, out_shape, stream=None):
    import torch
    assert src.is_cuda and src.is_contiguous() and src.dtype in (torch.int8, torch.int32)
    out = torch.empty(out_shape, dtype=src.dtype, device=src.device)
    fn = getattr(load_library(), fn_name)
    args = [_ptr(src), _ptr(out), n, c, h, w] + ([v] if v is not None else []) + [src.element_size(), _stream_ptr(stream)]
    check(fn(*args))
    return out


def to_vect_c(t, v: int = 16, stream=None):
    """[N,C,H,W] -> [N,C/V,H,W,V], materialised (utils.cuh:20-26)."""
    n, c, h, w = t.shape
    return _convert("lbc_to_vect_c", t, n, c, h, w, v, (n, c // v, h, w, v), stream)


def from_vect_c(t, stream=None):
    """[N,C/V,H,W,V] -> [N,C,H,W] (utils.cuh:11-17)."""
    n, cg, h, w, v = t.shape
    return _convert("lbc_from_vect_c", t, n, cg * v, h, w, v, (n, cg * v, h, w), stream)


def nhwc_to_vect_c(t, v: int = 16, stream=None):
    n, h, w, c = t.shape
    return _convert("lbc_nhwc_to_vect_c", t, n, c, h, w, v, (n, c // v, h, w, v), stream)


def vect_c_to_nhwc(t, stream=None):
    n, cg, h, w, v = t.shape
    return _convert("lbc_vect_c_to_nhwc", t, n, cg * v, h, w, v, (n, h, w, cg * v), stream)


def nchw_to_nhwc(t, stream=None):
    n, c, h, w = t.shape
    return _convert("lbc_nchw_to_nhwc", t, n, c, h, w, None, (n, h, w, c), stream)


def nhwc_to_nchw(t, stream=None):
    n, h, w, c = t.shape
    return _convert("lbc_nhwc_to_nchw", t, n, c, h, w, None, (n, c, h, w), stream)


def conv2DForward3x3(rinput, rkernel):
    """Reference-signature operator (conv2DForward3x3TensorCores.cuh:695-751): VECT_C=16 int8 input
    [N][C/16][H][W][16] and kernel [K][C/16][3][3][16]; 3x3, stride 1, no padding; returns
    (int32 [N][K/16][P][Q][16], elapsed_ms).  Unlike the reference there is no P,Q % 32 restriction."""
    import torch
    n, cg, h, w, v = rinput.shape
    k, cg2, r, s, v2 = rkernel.shape
    assert v == 16 and v2 == 16 and cg == cg2 and r == 3 and s == 3, "reference operator is 3x3 on VECT_C=16 tensors"
    c = cg * v
    x = vect_c_to_nhwc(rinput.contiguous())
    wk = vect_c_to_nhwc(rkernel.contiguous())            # [K][3][3][C] == KRSC
    plan = ConvPlan(ConvDesc(n=n, h=h, w=w, c=c, k=k, r=3, s=3, out_mode=_capi.OUT_INT32))
    wp = plan.prepack(wk.reshape(-1), _capi.W_KRSC)
    y, ms = plan.run(x, wp, timed=True)
    out = nhwc_to_vect_c(y, 16)
    torch.cuda.current_stream().synchronize()
    plan.close()
    return out, ms


# ---- backward passes as int8 convolutions (qconv2d.py:90-114) ----------------------------------------------
def _desc_from_c(cd: CConvDesc) -> ConvDesc:
    return ConvDesc(**{n: getattr(cd, n) for n, _ in CConvDesc._fields_})


def conv_backward_data(desc: ConvDesc, dy, w_krsc, stream=None):
    """dx (int32 NHWC [N,H,W,C]) of `desc` for dy int8 NHWC [N,P,Q,K] and the forward filter w int8 [K,R,S,C]: the forward
    engine on the 180-degree-rotated, transposed filter with padding R-1-pad (conv2DBackwardData3x3.cuh:61-64,126-127)."""
    import torch
    lib = load_library()
    fwd, dg = desc.c_struct(), CConvDesc()
    check(lib.lbc_conv_dgrad_desc(ctypes.byref(fwd), ctypes.byref(dg)))
    assert dy.is_cuda and dy.dtype == torch.int8 and dy.is_contiguous() and tuple(dy.shape) == (dg.n, dg.h, dg.w, dg.c)
    assert w_krsc.is_cuda and w_krsc.dtype == torch.int8 and w_krsc.is_contiguous()
    w2 = torch.empty(desc.c * desc.r * desc.s * desc.k, dtype=torch.int8, device=dy.device)
    check(lib.lbc_conv_dgrad_weights(ctypes.byref(fwd), _ptr(w_krsc), _ptr(w2), _stream_ptr(stream)))
    plan = ConvPlan(_desc_from_c(dg))
    dx = plan.run(dy, plan.prepack(w2, _capi.W_KRSC, stream=stream), stream=stream)
    plan.close()
    return dx


def conv_backward_weights(desc: ConvDesc, x, dy, stream=None):
    """dw (int32 [K,R,S,C]) of `desc` for x int8 NHWC and dy int8 NHWC [N,P,Q,K]: a convolution over the batch - x^T
    [C,H,W,N] as the input, dy^T [K,P,Q,N] as a P x Q filter (qconv2d.py:96-103; conv2DBackwardWeights3x3.cuh:15-100)."""
    import torch
    lib = load_library()
    fwd, wg = desc.c_struct(), CConvDesc()
    check(lib.lbc_conv_wgrad_desc(ctypes.byref(fwd), ctypes.byref(wg)))
    p, q = desc.out_hw
    assert x.is_cuda and x.dtype == torch.int8 and x.is_contiguous() and tuple(x.shape) == (desc.n, desc.h, desc.w, desc.c)
    assert dy.is_cuda and dy.dtype == torch.int8 and dy.is_contiguous() and tuple(dy.shape) == (desc.n, p, q, desc.k)
    xt = torch.empty((desc.c, desc.h, desc.w, desc.n), dtype=torch.int8, device=x.device)
    dyt = torch.empty((desc.k, p, q, desc.n), dtype=torch.int8, device=x.device)
    check(lib.lbc_nhwc_to_chwn(_ptr(x), _ptr(xt), desc.n, desc.h, desc.w, desc.c, 1, _stream_ptr(stream)))
    check(lib.lbc_nhwc_to_chwn(_ptr(dy), _ptr(dyt), desc.n, p, q, desc.k, 1, _stream_ptr(stream)))
    plan = ConvPlan(_desc_from_c(wg))
    dw_crsk = plan.run(xt, plan.prepack(dyt.reshape(-1), _capi.W_KRSC, stream=stream), stream=stream)     # [C][R][S][K]
    plan.close()
    dw = torch.empty((desc.k, desc.r, desc.s, desc.c), dtype=torch.int32, device=x.device)
    check(lib.lbc_nhwc_to_chwn(_ptr(dw_crsk), _ptr(dw), desc.c, desc.r, desc.s, desc.k, 4, _stream_ptr(stream)))
    return dw


class FusedTailPlan:
    """lbc_fused_plan: conv_a (R x S, stride 1, -> 64 channels) and conv_b (1x1, 64 -> 256) as ONE launch."""

    def __init__(self, conv_a: ConvDesc, conv_b: ConvDesc):
        self._lib = load_library()
        self.a, self.b = conv_a, conv_b
        self._h = ctypes.c_void_p()
        ca, cb = conv_a.c_struct(), conv_b.c_struct()
        check(self._lib.lbc_fused_tail_plan_create(ctypes.byref(ca), ctypes.byref(cb), ctypes.byref(self._h)))
        pa, pb = ctypes.c_void_p(), ctypes.c_void_p()
        check(self._lib.lbc_fused_tail_plan_parts(self._h, ctypes.byref(pa), ctypes.byref(pb)))
        self._pa, self._pb = pa, pb

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._lib.lbc_fused_tail_plan_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _prepack(self, part, w, n_elts):
        import torch
        assert w.is_cuda and w.dtype == torch.int8 and w.is_contiguous() and w.numel() == n_elts
        nb = ctypes.c_size_t()
        check(self._lib.lbc_conv_packed_weight_bytes(part, ctypes.byref(nb)))
        out = torch.empty(nb.value, dtype=torch.int8, device=w.device)
        check(self._lib.lbc_conv_prepack_weights(part, _ptr(w), _capi.W_KRSC, _ptr(out), _stream_ptr(None)))
        return out

    def prepack(self, wa_krsc, wb_krsc):
        """int8 CUDA tensors [64][R][S][C] and [256][1][1][64] -> the two packed filter matrices."""
        a, b = self.a, self.b
        return self._prepack(self._pa, wa_krsc, a.k * a.r * a.s * a.c), self._prepack(self._pb, wb_krsc, b.k * b.c)

    def run(self, x, wa, bias_a, scale_a, wb, bias_b, scale_b, out=None, stream=None, timed: bool = False):
        import torch
        b = self.b
        p, q = b.out_hw
        y = out if out is not None else torch.empty((b.n, p, q, b.k), dtype=torch.int8, device=x.device)
        ms = ctypes.c_float()
        check(self._lib.lbc_fused_tail_run(self._h, _ptr(x), _ptr(wa), _ptr(bias_a), _ptr(scale_a), _ptr(wb), _ptr(bias_b), _ptr(scale_b),
                                           _ptr(y), _stream_ptr(stream), ctypes.byref(ms) if timed else None))
        return (y, ms.value) if timed else y
