"""lowbitdnn-project_b200 — B200-native (sm_100a) int8 convolution library.

The product is liblowbit-cnn (lib/liblowbit_cnn.so, C ABI in include/lowbit_cnn.h): hand-written
tcgen05/TMA implicit-GEMM and CUDA-core kernels.  This package is the thin host-side binding used by the
tests and the benchmark harness; importing it never falls back to PyTorch or to the CPU oracle.
"""
from . import _capi
from ._capi import (KERNEL_AUTO, KERNEL_DEPTHWISE, KERNEL_DIRECT, KERNEL_IGEMM_TC, OUT_INT8, OUT_INT32, W_KRSC,
                    W_OIHW, LbcError, load_library)
from .conv import (ConvDesc, ConvPlan, FusedTailPlan, conv2DForward3x3, conv_backward_data, conv_backward_weights, from_vect_c, nchw_to_nhwc,
                   nhwc_to_nchw, nhwc_to_vect_c, to_vect_c, vect_c_to_nhwc)
from .net import Net
from .ops import AddDesc, PoolDesc, add_relu, global_avg_pool, max_pool2d
from . import networks
from . import shard

__all__ = [
    "ConvDesc", "ConvPlan", "FusedTailPlan", "Net", "networks", "shard", "LbcError", "load_library",
    "PoolDesc", "AddDesc", "max_pool2d", "add_relu", "global_avg_pool",
    "conv_backward_data", "conv_backward_weights", "conv2DForward3x3", "to_vect_c", "from_vect_c", "nhwc_to_vect_c", "vect_c_to_nhwc", "nchw_to_nhwc", "nhwc_to_nchw",
    "OUT_INT8", "OUT_INT32", "W_KRSC", "W_OIHW", "KERNEL_AUTO", "KERNEL_DIRECT", "KERNEL_IGEMM_TC", "KERNEL_DEPTHWISE",
]


def probe_int8_mma_peak(iters: int = 4096) -> float:
    import ctypes
    import torch
    v = ctypes.c_double()
    _capi.check(load_library().lbc_probe_int8_mma_peak(iters, ctypes.byref(v),
                                                      ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
    return v.value


def probe_hbm_copy(nbytes: int = 1 << 30, iters: int = 10) -> float:
    import ctypes
    import torch
    v = ctypes.c_double()
    _capi.check(load_library().lbc_probe_hbm_copy(nbytes, iters, ctypes.byref(v),
                                                 ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
    return v.value


def flush_l2(stream=None) -> None:
    import ctypes
    import torch
    s = stream if stream is not None else torch.cuda.current_stream()
    _capi.check(load_library().lbc_flush_l2(ctypes.c_void_p(s.cuda_stream)))
