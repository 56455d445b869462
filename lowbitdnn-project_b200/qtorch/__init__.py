"""qtorch-compatible surface on top of liblowbit-cnn.

The reference's ``python/qtorch`` package wraps cuDNN int8 convolutions for quantisation-aware training.  This package
keeps its names, argument order and arithmetic so that the reference's callers run against the new kernels:

    quantize / dequantize / QUANTIZATION_PARAMETERS      python/qtorch/nn/functional/quantization.py:113-152
    to_vect_c / from_vect_c (V = 4)                       python/qtorch/nn/functional/utils.py:5-30
    qconv2d(input, weight, stride, padding, dilation, groups)   python/qtorch/nn/functional/qconv2d.py:119-123
    QConv2D(nn.Conv2d)                                    python/qtorch/nn/QConv2d.py:6-22
    qmax_pool2d(input, kernel, stride, padding)           cpp.max_pool2d, python/qtorch/cpp/module.cu:9 (used by tmp.py:44)

Every int8 contraction (forward, data gradient, weight gradient) runs on the library's kernels with int32 accumulators;
PyTorch does the float <-> int8 conversions around them, as it does in the reference.
"""
from .functional import (QUANTIZATION_PARAMETERS, dequantize, from_vect_c, qconv2d, qmax_pool2d, quantize, to_vect_c)
from .nn import QConv2D

__all__ = ["quantize", "dequantize", "QUANTIZATION_PARAMETERS", "to_vect_c", "from_vect_c", "qconv2d", "qmax_pool2d", "QConv2D"]
