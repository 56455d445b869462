"""lbc_net wrapper: a list of planned convolutions run back to back (cpp/apps/benchmark.cpp:54-81 role)."""
from __future__ import annotations

import ctypes

from . import _capi
from ._capi import CConvDesc, check, load_library
from .conv import ConvDesc, _ptr, _stream_ptr
from .ops import AddDesc, PoolDesc


class Net:
    def __init__(self, layers, options: dict | None = None):
        """layers: [(name, ConvDesc, input_name_or_None)]; options: lbc_plan_options fields applied to every layer"""
        self._lib = load_library()
        self.names = [l[0] for l in layers]
        self.descs = [l[1] for l in layers]           # ConvDesc, or PoolDesc / AddDesc in a graph network
        idx = {n: i for i, n in enumerate(self.names)}
        n = len(layers)
        self._h = ctypes.c_void_p()
        opt = _capi.plan_options(**options) if options else None
        if all(isinstance(d, ConvDesc) for d in self.descs):
            self.input_of = [(-1 if l[2] is None else idx[l[2]]) for l in layers]
            arr = (CConvDesc * n)(*[d.c_struct() for d in self.descs])
            inp = (ctypes.c_int32 * n)(*self.input_of)
            if opt is not None:
                check(self._lib.lbc_net_create_ex(arr, inp, n, ctypes.byref(opt), ctypes.byref(self._h)))
            else:
                check(self._lib.lbc_net_create(arr, inp, n, ctypes.byref(self._h)))
            return
        # graph network: convolutions, max-pools and residual adds (an add names its two producers as a pair)
        nodes = (_capi.CNode * n)()
        self.input_of = []
        for i, (_, d, src) in enumerate(layers):
            nd = nodes[i]
            nd.input2_of = -1
            if isinstance(d, AddDesc):
                a, b = src
                nd.kind, nd.input_of, nd.input2_of, nd.relu = _capi.NODE_ADD, idx[a], idx[b], d.relu
            else:
                nd.input_of = -1 if src is None else idx[src]
                if isinstance(d, PoolDesc):
                    nd.kind, nd.pool = _capi.NODE_MAXPOOL, d.c_struct()
                else:
                    nd.kind, nd.conv = _capi.NODE_CONV, d.c_struct()
            self.input_of.append(nd.input_of)
        check(self._lib.lbc_net_create_graph(nodes, n, ctypes.byref(opt) if opt is not None else None, ctypes.byref(self._h)))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._lib.lbc_net_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __len__(self):
        return len(self.descs)

    def set_params(self, layer: int, w_host, bias_host, scale_host, layout: int = _capi.W_KRSC):
        """numpy arrays (int8 KRSC/OIHW weights, int32 bias, f32 scale) -> device, pre-packed."""
        def p(a):
            return None if a is None else a.ctypes.data_as(ctypes.c_void_p)
        check(self._lib.lbc_net_set_params_host(self._h, layer, p(w_host), layout, p(bias_host), p(scale_host)))

    def set_input(self, layer: int, x_host):
        """numpy int8 NHWC -> the resident input buffer of a layer fed from outside the conv chain."""
        check(self._lib.lbc_net_set_input_host(self._h, layer, x_host.ctypes.data_as(ctypes.c_void_p)))

    def layer_io(self, layer: int) -> tuple[int, int]:
        x, y = ctypes.c_void_p(), ctypes.c_void_p()
        check(self._lib.lbc_net_layer_io(self._h, layer, ctypes.byref(x), ctypes.byref(y)))
        return x.value, y.value

    def read_output(self, layer: int, images: int | None = None):
        """One layer's resident output (its first `images` images, default all) as a numpy array (NHWC; int8 or int32
        per the layer's out_mode)."""
        import numpy as np
        d = self.descs[layer]
        p, q = d.out_hw
        n = d.n if images is None else min(images, d.n)
        k = d.k if isinstance(d, ConvDesc) else d.c
        out = np.empty((n, p, q, k), dtype=np.int32 if getattr(d, "out_mode", _capi.OUT_INT8) == _capi.OUT_INT32 else np.int8)
        check(self._lib.lbc_net_read_output_host(self._h, layer, out.ctypes.data_as(ctypes.c_void_p), out.nbytes))
        return out

    def fused_into(self, layer: int) -> int:
        """Index of the layer whose fused launch absorbs `layer` (its own output is then never materialised), or -1."""
        v = ctypes.c_int32()
        check(self._lib.lbc_net_layer_fused_into(self._h, layer, ctypes.byref(v)))
        return v.value

    def check_status(self) -> None:
        check(self._lib.lbc_net_check(self._h))

    def layer_kernel(self, layer: int) -> str:
        if isinstance(self.descs[layer], PoolDesc):
            return "maxpool"
        if isinstance(self.descs[layer], AddDesc):
            return "add_relu"
        plan = ctypes.c_void_p()
        check(self._lib.lbc_net_layer_plan(self._h, layer, ctypes.byref(plan)))
        k = ctypes.c_int32()
        check(self._lib.lbc_conv_plan_kernel(plan, ctypes.byref(k)))
        return _capi.KERNEL_NAMES[k.value]

    def layer_describe(self, layer: int) -> str:
        if not isinstance(self.descs[layer], ConvDesc):
            return f"{self.layer_kernel(layer)} {self.descs[layer]}"
        plan = ctypes.c_void_p()
        check(self._lib.lbc_net_layer_plan(self._h, layer, ctypes.byref(plan)))
        buf = ctypes.create_string_buffer(512)
        check(self._lib.lbc_conv_plan_describe(plan, buf, 512))
        return buf.value.decode()

    @property
    def launches(self) -> int:
        n = ctypes.c_int32()
        check(self._lib.lbc_net_launches(self._h, ctypes.byref(n)))
        return n.value

    def run(self, x_dev=None, stream=None, timed: bool = False):
        """Enqueue every layer. timed=True synchronises and returns (per_layer_ms list, total_ms)."""
        n = len(self)
        if not timed:
            check(self._lib.lbc_net_run(self._h, _ptr(x_dev), _stream_ptr(stream), None, None))
            return None
        per = (ctypes.c_float * n)()
        tot = ctypes.c_float()
        check(self._lib.lbc_net_run(self._h, _ptr(x_dev), _stream_ptr(stream), per, ctypes.byref(tot)))
        return list(per), tot.value

    def submit_host(self, x_host, y_host, stream=None) -> None:
        """Pipelined end to end: enqueue H2D -> all layers -> D2H and return; see sync_host()."""
        check(self._lib.lbc_net_submit_host(self._h, _ptr(x_host), _ptr(y_host), _stream_ptr(stream)))

    def sync_host(self) -> float:
        """Wait for every submit_host(); device ms from the first upload to the last download."""
        tot = ctypes.c_float()
        check(self._lib.lbc_net_sync_host(self._h, ctypes.byref(tot)))
        return tot.value

    def run_host(self, x_host, y_host, stream=None) -> float:
        tot = ctypes.c_float()
        check(self._lib.lbc_net_run_host(self._h, _ptr(x_host), _ptr(y_host), _stream_ptr(stream), ctypes.byref(tot)))
        return tot.value
