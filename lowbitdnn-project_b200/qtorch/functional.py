"""Functional half of the qtorch-compatible surface (see package docstring)."""
from __future__ import annotations

import numpy as np
import torch
from torch import Tensor
from torch.autograd import Function

from .. import _capi
from ..conv import ConvDesc, ConvPlan, conv_backward_data, conv_backward_weights, nchw_to_nhwc, nhwc_to_nchw
from ..ops import max_pool2d as _max_pool2d_nhwc

__all__ = ["quantize", "dequantize", "QUANTIZATION_PARAMETERS", "to_vect_c", "from_vect_c", "qconv2d", "qmax_pool2d"]

NUM_BITS = 8
DTYPE = torch.int8
# quantized tensor -> its scale (the reference's "work around to store quantization parameters", quantization.py:22-23)
QUANTIZATION_PARAMETERS: dict = {}


# ---- layout views (python/qtorch/nn/functional/utils.py) ----------------------------------------------------
def to_vect_c(tensor: Tensor, contiguous: bool = False) -> Tensor:
    n, c, h, w = tensor.shape
    if c % 4 != 0:
        raise NotImplementedError("VECT_C works only with tensors which channels are either multiple of 4")
    tensor = tensor.reshape(n, c // 4, 4, h, w).permute([0, 1, 3, 4, 2])
    return tensor.contiguous() if contiguous else tensor


def from_vect_c(tensor: Tensor, contiguous: bool = True) -> Tensor:
    n, c, h, w, v = tensor.shape
    tensor = tensor.permute([0, 1, 4, 2, 3]).reshape(n, c * v, h, w)
    return tensor.contiguous() if contiguous else tensor


# ---- quantize / dequantize (quantization.py:27-152) ----------------------------------------------------------
def _quantization_params(num_bits: int, min_value: float, max_value: float, signed: bool):
    if signed:
        qmin, qmax = -2**num_bits // 2, 2**num_bits // 2 - 1
    else:
        qmin, qmax = 0, 2**num_bits - 1
    scale = (max_value - min_value) / (qmax - qmin)          # qvalue * scale = float_value
    inv_scale = (qmax - qmin) / (max_value - min_value)      # float_value * inv_scale = qvalue
    qzero = int(np.clip(int(round(qmin - min_value * inv_scale)), qmin, qmax))
    return qmin, qmax, qzero, scale, inv_scale


def quantize(tensor: Tensor, to_vect_c: bool = False, num_bits=None, min_value=None, max_value=None, stochastic: bool = False) -> Tensor:
    """Symmetric per-tensor quantisation to int8: mul by inv_scale, [+U(-.5,.5)], clamp to [-128,127], round half-to-even.
    Already-quantized tensors are returned as they are (quantization.py:116-117)."""
    if tensor in QUANTIZATION_PARAMETERS:
        return tensor
    num_bits = NUM_BITS if not num_bits else num_bits
    assert num_bits <= NUM_BITS, "num bits > 8 are not supported"
    with torch.no_grad():
        lo = float(tensor.min()) if not min_value else float(min_value)
        hi = float(tensor.max()) if not max_value else float(max_value)
        max_abs = max(abs(lo), hi)
        if max_abs == 0.0:
            max_abs = 1.0                                     # an all-zero tensor quantises to zeros (the reference divides by 0)
        qmin, qmax, _qzero, scale, inv_scale = _quantization_params(num_bits, -max_abs, max_abs, signed=True)
        t = globals()["to_vect_c"](tensor) if to_vect_c else tensor
        t = t.detach().float().mul(inv_scale)
        if stochastic:
            t = t + torch.empty_like(t).uniform_(-0.5, 0.5)
        result = t.clamp_(qmin, qmax).round_().to(DTYPE)
    QUANTIZATION_PARAMETERS[result] = scale
    return result


def dequantize(tensor: Tensor, from_vect_c: bool = False, scale: float = None) -> Tensor:
    if not scale:
        assert tensor in QUANTIZATION_PARAMETERS, "Tried to dequantize not quantized tensor"
        scale = QUANTIZATION_PARAMETERS[tensor]
    assert tensor.dtype == torch.int8
    if from_vect_c:
        tensor = globals()["from_vect_c"](tensor)
    return tensor.float() * scale


# ---- the int8 convolution behind cpp.conv2d(..., "external", scale, ...) (python/qtorch/cpp/conv2d.cuh:95-157) ----
def _pair(v):
    return (int(v), int(v)) if isinstance(v, int) else (int(v[0]), int(v[1]))


def _conv_desc(x_shape, w_shape, stride, padding, dilation, groups) -> ConvDesc:
    n, c, h, w = x_shape
    k, cg, r, s = w_shape
    (sh, sw), (ph, pw), (dh, dw) = _pair(stride), _pair(padding), _pair(dilation)
    assert cg * groups == c, "weight channels x groups must equal the input channels"
    return ConvDesc(n=n, h=h, w=w, c=c, k=k, r=r, s=s, stride_h=sh, stride_w=sw, pad_h=ph, pad_w=pw, dil_h=dh, dil_w=dw,
                    groups=groups, out_mode=_capi.OUT_INT32)


def _int8_conv_int32(xq_nchw: Tensor, wq_oihw: Tensor, desc: ConvDesc) -> Tensor:
    """int8 NCHW x int8 OIHW -> int32 NCHW, exact accumulators, on the library's kernels."""
    x = nchw_to_nhwc(xq_nchw.contiguous())
    plan = ConvPlan(desc)
    wp = plan.prepack(wq_oihw.contiguous().reshape(-1), _capi.W_OIHW)
    y = plan.run(x, wp)
    plan.close()
    return nhwc_to_nchw(y)


class _QConv2d(Function):
    """quantize -> int8 convolution -> float (forward); both gradients as int8 convolutions of the quantized output
    gradient (backward) - the scheme of python/qtorch/nn/functional/qconv2d.py:50-116."""

    @staticmethod
    def forward(ctx, input: Tensor, weight: Tensor, strides, padding, dilation, groups: int) -> Tensor:
        xq = quantize(input, to_vect_c=False, stochastic=False)
        wq = quantize(weight, to_vect_c=False, stochastic=False)
        iscale, wscale = QUANTIZATION_PARAMETERS[xq], QUANTIZATION_PARAMETERS[wq]
        desc = _conv_desc(xq.shape, wq.shape, strides, padding, dilation, groups)
        ctx.desc, ctx.iscale, ctx.wscale = desc, iscale, wscale
        ctx.save_for_backward(xq, wq)
        acc = _int8_conv_int32(xq, wq, desc)
        # tensors quantized here (not handed in already quantized) leave the registry again: it would otherwise keep every
        # activation of a training run alive (the reference's registry only ever grows)
        if xq is not input:
            QUANTIZATION_PARAMETERS.pop(xq, None)
        if wq is not weight:
            QUANTIZATION_PARAMETERS.pop(wq, None)
        return acc.float() * (wscale * iscale)                 # cuDNN "external": float(alpha * acc), alpha = iscale * wscale

    @staticmethod
    def backward(ctx, grad_output: Tensor):
        xq, wq = ctx.saved_tensors
        d: ConvDesc = ctx.desc
        assert (d.dil_h, d.dil_w) == (1, 1), "Only (1, 1) dilation is supported"
        assert d.groups == 1, "Only 1 groups is supported"
        assert (d.stride_h, d.stride_w) == (1, 1), "No way to compute for strides != 1. Only by inserting dummy rows and columns"
        gq = quantize(grad_output.contiguous(), to_vect_c=False)
        gscale = QUANTIZATION_PARAMETERS[gq]
        fwd = d.replace(out_mode=_capi.OUT_INT8)
        x_nhwc, g_nhwc = nchw_to_nhwc(xq.contiguous()), nchw_to_nhwc(gq)
        w_krsc = nchw_to_nhwc(wq.contiguous())                  # OIHW -> [K][R][S][C]
        grad_weight = conv_backward_weights(fwd, x_nhwc, g_nhwc)                       # int32 [K][R][S][C]
        grad_weight = nhwc_to_nchw(grad_weight).float() * (ctx.iscale * gscale)        # -> OIHW, qconv_scale = iscale * gscale
        grad_input = conv_backward_data(fwd, g_nhwc, w_krsc.reshape(-1))               # int32 [N][H][W][C]
        grad_input = nhwc_to_nchw(grad_input).float() * (gscale * ctx.wscale)
        QUANTIZATION_PARAMETERS.pop(gq, None)
        return grad_input, grad_weight, None, None, None, None


def qconv2d(input, weight, stride, padding, dilation, groups):
    return _QConv2d.apply(input, weight, stride, padding, dilation, groups)


def qmax_pool2d(input: Tensor, kernel, stride, padding) -> Tensor:
    """int8 max-pool on the reference's [N, C/V, H, W, V] tensors (cpp.max_pool2d; python/tmp.py:44,48,52)."""
    assert input.dim() == 5 and input.dtype == torch.int8, "input should have exactly 5 dimensions (N x C/V x H x W x V), int8"
    n, cg, h, w, v = input.shape
    x = input.permute(0, 2, 3, 1, 4).reshape(n, h, w, cg * v).contiguous()           # NHWC
    y = _max_pool2d_nhwc(x, kernel, stride, padding)
    p, q = y.shape[1], y.shape[2]
    return y.reshape(n, p, q, cg, v).permute(0, 3, 1, 2, 4).contiguous()
