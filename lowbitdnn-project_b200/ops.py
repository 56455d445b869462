"""int8 ops between convolutions (include/lowbit_cnn.h: lbc_maxpool2d_run / lbc_add_relu_run / lbc_global_avgpool_run).

Names and argument order follow the reference's Python surface where it has one:
  max_pool2d(input, kernel, stride, padding)        python/qtorch/cpp/module.cu:9 -> pool2d.cuh:54-92 (cuDNN int8 max-pool)
The tensors here are int8 NHWC CUDA tensors (the library's activation layout); PyTorch only provides the memory."""
from __future__ import annotations

import ctypes
from dataclasses import asdict, dataclass

from . import _capi
from ._capi import CPoolDesc, check, load_library
from .conv import _ptr, _stream_ptr


@dataclass(frozen=True)
class PoolDesc:
    """lbc_pool_desc."""
    n: int
    h: int
    w: int
    c: int
    kh: int
    kw: int
    stride_h: int = 1
    stride_w: int = 1
    pad_h: int = 0
    pad_w: int = 0

    def c_struct(self) -> CPoolDesc:
        return CPoolDesc(**asdict(self))

    @property
    def out_hw(self) -> tuple[int, int]:
        p, q = ctypes.c_int32(), ctypes.c_int32()
        st = self.c_struct()
        check(load_library().lbc_pool_out_shape(ctypes.byref(st), ctypes.byref(p), ctypes.byref(q)))
        return p.value, q.value


@dataclass(frozen=True)
class AddDesc:
    """A residual join: y = sat_int8(a + b), ReLU when `relu`; n/h/w/c is the shape of both operands."""
    n: int
    h: int
    w: int
    c: int
    relu: int = 1

    @property
    def out_hw(self) -> tuple[int, int]:
        return self.h, self.w


def _pair(v):
    return (v, v) if isinstance(v, int) else tuple(v)


def max_pool2d(x, kernel, stride=None, padding=0, stream=None):
    """int8 NHWC max-pool; kernel / stride / padding are ints or (h, w) pairs, stride defaults to the kernel."""
    import torch
    assert x.is_cuda and x.dtype == torch.int8 and x.is_contiguous() and x.dim() == 4
    n, h, w, c = x.shape
    kh, kw = _pair(kernel)
    sh, sw = _pair(kernel if stride is None else stride)
    ph, pw = _pair(padding)
    d = PoolDesc(n, h, w, c, kh, kw, sh, sw, ph, pw)
    p, q = d.out_hw
    y = torch.empty((n, p, q, c), dtype=torch.int8, device=x.device)
    st = d.c_struct()
    check(load_library().lbc_maxpool2d_run(ctypes.byref(st), _ptr(x), _ptr(y), _stream_ptr(stream)))
    return y


def add_relu(a, b, relu: bool = True, out=None, stream=None):
    """y = clamp(a + b, relu ? 0 : -128, 127) on int8 CUDA tensors of equal shape (the bottleneck's residual join)."""
    import torch
    assert a.is_cuda and b.is_cuda and a.dtype == torch.int8 and b.dtype == torch.int8
    assert a.shape == b.shape and a.is_contiguous() and b.is_contiguous()
    y = out if out is not None else torch.empty_like(a)
    check(load_library().lbc_add_relu_run(_ptr(a), _ptr(b), _ptr(y), a.numel(), int(relu), _stream_ptr(stream)))
    return y


def global_avg_pool(x, scale: float, stream=None):
    """[N,H,W,C] int8 -> [N,C] int8: sat_int8(rint(sum_{h,w} x * scale)), the convolutions' requantisation rule."""
    import torch
    assert x.is_cuda and x.dtype == torch.int8 and x.is_contiguous() and x.dim() == 4
    n, h, w, c = x.shape
    y = torch.empty((n, c), dtype=torch.int8, device=x.device)
    check(load_library().lbc_global_avgpool_run(_ptr(x), n, h * w, c, ctypes.c_float(scale), _ptr(y), _stream_ptr(stream)))
    return y
