"""ctypes declarations for include/lowbit_cnn.h.  The library is REQUIRED: nothing here falls back to
PyTorch or to the CPU oracle — a missing or unloadable liblowbit_cnn.so raises immediately."""
from __future__ import annotations

import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "liblowbit_cnn.so")

LBC_OK = 0
STATUS_NAMES = {0: "LBC_OK", 1: "LBC_ERR_INVALID_ARG", 2: "LBC_ERR_UNSUPPORTED", 3: "LBC_ERR_NO_DEVICE",
                4: "LBC_ERR_CUDA", 5: "LBC_ERR_ALLOC", 6: "LBC_ERR_KERNEL_TIMEOUT"}
OUT_INT8, OUT_INT32 = 0, 1
W_KRSC, W_OIHW = 0, 1
KERNEL_AUTO, KERNEL_DIRECT, KERNEL_IGEMM_TC, KERNEL_DEPTHWISE, KERNEL_STEM_TC = 0, 1, 2, 3, 4
KERNEL_NAMES = {1: "direct", 2: "igemm_tc", 3: "depthwise", 4: "stem_tc"}


class LbcError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"{STATUS_NAMES.get(status, status)}: {message}")
        self.status = status


class CConvDesc(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in (
        "n", "h", "w", "c", "k", "r", "s", "stride_h", "stride_w", "pad_h", "pad_w",
        "dil_h", "dil_w", "groups", "relu", "out_mode")]


class CPoolDesc(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in ("n", "h", "w", "c", "kh", "kw", "stride_h", "stride_w", "pad_h", "pad_w")]


class CNode(ctypes.Structure):
    """lbc_node: one node of a network graph."""
    _fields_ = [("kind", ctypes.c_int32), ("input_of", ctypes.c_int32), ("input2_of", ctypes.c_int32), ("relu", ctypes.c_int32),
                ("conv", CConvDesc), ("pool", CPoolDesc)]


NODE_CONV, NODE_MAXPOOL, NODE_ADD = 0, 1, 2


class CPlanOptions(ctypes.Structure):
    """lbc_plan_options (include/lowbit_cnn.h); build one with plan_options(**kw)."""
    _fields_ = [(n, ctypes.c_int32) for n in (
        "struct_size", "cta_pairs", "warp_store", "fold_bias", "paired_tiles", "resident_filter", "window", "keep_window",
        "force_im2col", "pixel_groups", "dw_tiled", "reverse", "pdl", "two_mma_warps", "tiles_per_iter2", "small_teams",
        "four_acc", "n_stationary", "epi_pipeline", "max_grid", "max_bn", "max_stages", "max_win_stages", "stage_bufs",
        "tps_kb", "resident_kb", "epi_split", "fuse", "early_weights", "tail_split")] + [("reserved", ctypes.c_int32 * 4)]


def plan_options(**kw) -> CPlanOptions:
    """Planner options with the library's defaults, then the given fields (e.g. cta_pairs=1, max_grid=3)."""
    o = CPlanOptions()
    load_library().lbc_plan_options_init(ctypes.byref(o))
    for k, v in kw.items():
        if k not in dict(CPlanOptions._fields_) or k in ("struct_size", "reserved"):
            raise KeyError(f"unknown planner option {k!r}")
        setattr(o, k, int(v))
    return o


_vp = ctypes.c_void_p
_i32 = ctypes.c_int32
_PROTOS = {
    "lbc_version": (ctypes.c_int, []),
    "lbc_last_error_string": (ctypes.c_char_p, []),
    "lbc_device_info": (ctypes.c_int, [ctypes.c_int] + [ctypes.POINTER(ctypes.c_int)] * 3 + [ctypes.POINTER(ctypes.c_size_t)]),
    "lbc_conv_out_shape": (ctypes.c_int, [ctypes.POINTER(CConvDesc), ctypes.POINTER(_i32), ctypes.POINTER(_i32)]),
    "lbc_conv_work": (ctypes.c_int, [ctypes.POINTER(CConvDesc), ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double)]),
    "lbc_conv_plan_create": (ctypes.c_int, [ctypes.POINTER(CConvDesc), _i32, ctypes.POINTER(_vp)]),
    "lbc_plan_options_init": (None, [ctypes.POINTER(CPlanOptions)]),
    "lbc_conv_plan_create_ex": (ctypes.c_int, [ctypes.POINTER(CConvDesc), _i32, ctypes.POINTER(CPlanOptions), ctypes.POINTER(_vp)]),
    "lbc_conv_plan_dry_ex": (ctypes.c_int, [ctypes.POINTER(CConvDesc), _i32, ctypes.POINTER(CPlanOptions), _i32, ctypes.POINTER(_i32),
                                            ctypes.c_char_p, ctypes.c_size_t]),
    "lbc_conv_plan_set_trace": (ctypes.c_int, [_vp, _vp, _i32]),
    "lbc_conv_plan_check": (ctypes.c_int, [_vp]),
    "lbc_conv_plan_destroy": (ctypes.c_int, [_vp]),
    "lbc_conv_plan_kernel": (ctypes.c_int, [_vp, ctypes.POINTER(_i32)]),
    "lbc_conv_plan_describe": (ctypes.c_int, [_vp, ctypes.c_char_p, ctypes.c_size_t]),
    "lbc_conv_plan_launches": (ctypes.c_int, [_vp, ctypes.POINTER(_i32)]),
    "lbc_conv_packed_weight_bytes": (ctypes.c_int, [_vp, ctypes.POINTER(ctypes.c_size_t)]),
    "lbc_conv_prepack_weights": (ctypes.c_int, [_vp, _vp, _i32, _vp, _vp]),
    "lbc_conv_run": (ctypes.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, ctypes.POINTER(ctypes.c_float)]),
    "lbc_conv_run_host": (ctypes.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, ctypes.POINTER(ctypes.c_float)]),
    "lbc_fused_tail_plan_create": (ctypes.c_int, [ctypes.POINTER(CConvDesc), ctypes.POINTER(CConvDesc), ctypes.POINTER(_vp)]),
    "lbc_fused_tail_plan_destroy": (ctypes.c_int, [_vp]),
    "lbc_fused_tail_plan_parts": (ctypes.c_int, [_vp, ctypes.POINTER(_vp), ctypes.POINTER(_vp)]),
    "lbc_fused_tail_run": (ctypes.c_int, [_vp] * 10 + [ctypes.POINTER(ctypes.c_float)]),
    "lbc_net_layer_fused_into": (ctypes.c_int, [_vp, _i32, ctypes.POINTER(_i32)]),
    "lbc_pool_out_shape": (ctypes.c_int, [ctypes.POINTER(CPoolDesc), ctypes.POINTER(_i32), ctypes.POINTER(_i32)]),
    "lbc_maxpool2d_run": (ctypes.c_int, [ctypes.POINTER(CPoolDesc), _vp, _vp, _vp]),
    "lbc_add_relu_run": (ctypes.c_int, [_vp, _vp, _vp, ctypes.c_size_t, _i32, _vp]),
    "lbc_global_avgpool_run": (ctypes.c_int, [_vp, _i32, _i32, _i32, ctypes.c_float, _vp, _vp]),
    "lbc_to_vect_c": (ctypes.c_int, [_vp, _vp] + [_i32] * 6 + [_vp]),
    "lbc_from_vect_c": (ctypes.c_int, [_vp, _vp] + [_i32] * 6 + [_vp]),
    "lbc_nhwc_to_vect_c": (ctypes.c_int, [_vp, _vp] + [_i32] * 6 + [_vp]),
    "lbc_vect_c_to_nhwc": (ctypes.c_int, [_vp, _vp] + [_i32] * 6 + [_vp]),
    "lbc_nchw_to_nhwc": (ctypes.c_int, [_vp, _vp] + [_i32] * 5 + [_vp]),
    "lbc_nhwc_to_nchw": (ctypes.c_int, [_vp, _vp] + [_i32] * 5 + [_vp]),
    "lbc_conv_dgrad_desc": (ctypes.c_int, [ctypes.POINTER(CConvDesc), ctypes.POINTER(CConvDesc)]),
    "lbc_conv_wgrad_desc": (ctypes.c_int, [ctypes.POINTER(CConvDesc), ctypes.POINTER(CConvDesc)]),
    "lbc_conv_dgrad_weights": (ctypes.c_int, [ctypes.POINTER(CConvDesc), _vp, _vp, _vp]),
    "lbc_nhwc_to_chwn": (ctypes.c_int, [_vp, _vp] + [_i32] * 5 + [_vp]),
    "lbc_conv_plan_dry": (ctypes.c_int, [ctypes.POINTER(CConvDesc), _i32, _i32, ctypes.POINTER(_i32), ctypes.c_char_p, ctypes.c_size_t]),
    "lbc_net_create": (ctypes.c_int, [ctypes.POINTER(CConvDesc), ctypes.POINTER(_i32), _i32, ctypes.POINTER(_vp)]),
    "lbc_net_create_ex": (ctypes.c_int, [ctypes.POINTER(CConvDesc), ctypes.POINTER(_i32), _i32, ctypes.POINTER(CPlanOptions),
                                         ctypes.POINTER(_vp)]),
    "lbc_net_check": (ctypes.c_int, [_vp]),
    "lbc_net_create_graph": (ctypes.c_int, [ctypes.POINTER(CNode), _i32, ctypes.POINTER(CPlanOptions), ctypes.POINTER(_vp)]),
    "lbc_net_destroy": (ctypes.c_int, [_vp]),
    "lbc_net_layer_plan": (ctypes.c_int, [_vp, _i32, ctypes.POINTER(_vp)]),
    "lbc_net_set_params_host": (ctypes.c_int, [_vp, _i32, _vp, _i32, _vp, _vp]),
    "lbc_net_set_input_host": (ctypes.c_int, [_vp, _i32, _vp]),
    "lbc_net_read_output_host": (ctypes.c_int, [_vp, _i32, _vp, ctypes.c_size_t]),
    "lbc_net_layer_io": (ctypes.c_int, [_vp, _i32, ctypes.POINTER(_vp), ctypes.POINTER(_vp)]),
    "lbc_net_run": (ctypes.c_int, [_vp, _vp, _vp, ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_float)]),
    "lbc_net_run_host": (ctypes.c_int, [_vp, _vp, _vp, _vp, ctypes.POINTER(ctypes.c_float)]),
    "lbc_net_submit_host": (ctypes.c_int, [_vp, _vp, _vp, _vp]),
    "lbc_net_sync_host": (ctypes.c_int, [_vp, ctypes.POINTER(ctypes.c_float)]),
    "lbc_net_launches": (ctypes.c_int, [_vp, ctypes.POINTER(_i32)]),
    "lbc_probe_int8_mma_peak": (ctypes.c_int, [_i32, ctypes.POINTER(ctypes.c_double), _vp]),
    "lbc_probe_hbm_copy": (ctypes.c_int, [ctypes.c_size_t, _i32, ctypes.POINTER(ctypes.c_double), _vp]),
    "lbc_flush_l2": (ctypes.c_int, [_vp]),
}
EXPORTED_SYMBOLS = tuple(_PROTOS)

_lib = None


def load_library(path: str | None = None):
    """Loads liblowbit_cnn.so.  Raises FileNotFoundError / OSError if it is missing — by design."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise FileNotFoundError(
            f"{p} is missing: build it with `python lowbitdnn-project_b200/build.py` "
            "(liblowbit-cnn has no Python/CPU fallback)")
    lib = ctypes.CDLL(p)
    for name, (res, args) in _PROTOS.items():
        fn = getattr(lib, name)  # AttributeError if the header and the library disagree
        fn.restype = res
        fn.argtypes = args
    if path is None:
        _lib = lib
    return lib


def check(status: int) -> None:
    if status != LBC_OK:
        raise LbcError(status, load_library().lbc_last_error_string().decode())
