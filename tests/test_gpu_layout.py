"""GPU tests for the reference tensor-format converters and the reference-signature operator."""
import numpy as np
import pytest

from oracle import oracle

pytestmark = pytest.mark.gpu


def test_vect_c_converters_match_reference_definition():
    import torch
    import lowbitdnn_project_b200 as lbc
    rng = np.random.default_rng(0)
    for v, dt in ((16, np.int8), (4, np.int8), (32, np.int8), (16, np.int32)):
        a = rng.integers(-100, 100, size=(2, 2 * v, 5, 7)).astype(dt)
        t = torch.from_numpy(a).cuda()
        got = lbc.to_vect_c(t, v)
        assert np.array_equal(got.cpu().numpy(), oracle.to_vect_c(a, v))
        back = lbc.from_vect_c(got)
        assert np.array_equal(back.cpu().numpy(), a)
        nhwc = lbc.nchw_to_nhwc(t)
        assert np.array_equal(nhwc.cpu().numpy(), a.transpose(0, 2, 3, 1))
        assert np.array_equal(lbc.nhwc_to_nchw(nhwc).cpu().numpy(), a)
        vc = lbc.nhwc_to_vect_c(nhwc, v)
        assert np.array_equal(vc.cpu().numpy(), oracle.to_vect_c(a, v))
        assert np.array_equal(lbc.vect_c_to_nhwc(vc).cpu().numpy(), a.transpose(0, 2, 3, 1))


def test_reference_signature_operator():
    """conv2DForward3x3(to_vect_c(x), to_vect_c(w)) == to_vect_c(refConv2DForward(x, w)) — the comparison
    check.cu:107-129 makes, on a shape the reference kernel itself cannot run (P,Q not multiples of 32)."""
    import torch
    import lowbitdnn_project_b200 as lbc
    rng = np.random.default_rng(1)
    x = rng.integers(-128, 128, size=(2, 32, 12, 14), dtype=np.int8)      # NCHW, pre-padded
    w = rng.integers(-128, 128, size=(48, 32, 3, 3), dtype=np.int8)       # OIHW
    want = oracle.ref_style_nchw_valid(x, w)                               # int32 NCHW
    xv = lbc.to_vect_c(torch.from_numpy(x).cuda(), 16)
    wv = lbc.to_vect_c(torch.from_numpy(w).cuda(), 16)
    out, ms = lbc.conv2DForward3x3(xv, wv)
    assert ms > 0
    assert np.array_equal(out.cpu().numpy(), oracle.to_vect_c(want, 16))


def test_pipelined_host_path_matches_blocking_path():
    """lbc_net_submit_host (double-buffered input, copy streams) == lbc_net_run_host == oracle, step after step."""
    import numpy as np
    import torch
    import lowbitdnn_project_b200 as lbc
    from oracle import oracle
    from oracle.oracle import ConvDesc as OD
    n = 4
    layers = [("a", lbc.ConvDesc(n=n, h=20, w=20, c=3, k=32, r=3, s=3, stride_h=2, stride_w=2, pad_h=1, pad_w=1, relu=1), None),
              ("b", lbc.ConvDesc(n=n, h=10, w=10, c=32, k=64, r=3, s=3, pad_h=1, pad_w=1, relu=1), "a"),
              ("c", lbc.ConvDesc(n=n, h=10, w=10, c=64, k=32, r=1, s=1), "b")]
    net = lbc.Net(layers)
    params = []
    for i, (_, d, _) in enumerate(layers):
        od = OD(**d.__dict__)
        _, w, b, s = oracle.synth(od, layer=i)
        net.set_params(i, w, b, s)
        params.append((od, w, b, s))
    rng = np.random.default_rng(5)
    xs = [rng.integers(-128, 128, size=(n, 20, 20, 3), dtype=np.int8) for _ in range(5)]
    want = []
    for x in xs:
        t = x
        for od, w, b, s in params:
            t = oracle.conv_nhwc(od, t, w, b, s)
        want.append(t)
    xh = [torch.from_numpy(x).pin_memory() for x in xs]
    yh = [torch.empty((n, 10, 10, 32), dtype=torch.int8).pin_memory() for _ in xs]
    for x, y in zip(xh, yh):
        net.submit_host(x, y)
    assert net.sync_host() > 0
    for y, w in zip(yh, want):
        assert np.array_equal(y.numpy(), w)
    yb = torch.empty_like(yh[0])
    net.run_host(xh[2], yb)
    assert np.array_equal(yb.numpy(), want[2])
    net.close()


def test_network_runner_matches_oracle_with_branches():
    """A bottleneck-shaped graph (one producer feeding two consumers, a stride-2 branch, an odd image count) through
    lbc_net_run: the runner alternates the tile traversal direction along producer -> consumer edges, and every layer's
    resident output must still equal the oracle's, image for image."""
    import numpy as np
    import torch
    import lowbitdnn_project_b200 as lbc
    from oracle import oracle
    from oracle.oracle import ConvDesc as OD
    n = 5
    layers = [("in", lbc.ConvDesc(n=n, h=28, w=28, c=64, k=64, r=1, s=1, relu=1), None),
              ("conv1", lbc.ConvDesc(n=n, h=28, w=28, c=64, k=32, r=1, s=1, relu=1), "in"),
              ("conv2", lbc.ConvDesc(n=n, h=28, w=28, c=32, k=32, r=3, s=3, pad_h=1, pad_w=1, relu=1), "conv1"),
              ("conv3", lbc.ConvDesc(n=n, h=28, w=28, c=32, k=128, r=1, s=1, relu=1), "conv2"),
              ("down", lbc.ConvDesc(n=n, h=28, w=28, c=64, k=128, r=1, s=1, stride_h=2, stride_w=2), "in"),
              ("next", lbc.ConvDesc(n=n, h=28, w=28, c=128, k=256, r=3, s=3, stride_h=2, stride_w=2, pad_h=1, pad_w=1, relu=1), "conv3")]
    net = lbc.Net(layers)
    names = [l[0] for l in layers]
    rng = np.random.default_rng(17)
    x = rng.integers(-128, 128, size=(n, 28, 28, 64), dtype=np.int8)
    outs = {}
    for i, (name, d, src) in enumerate(layers):
        od = OD(**d.__dict__)
        _, w, b, s = oracle.synth(od, layer=i)
        net.set_params(i, w, b, s)
        outs[name] = oracle.conv_nhwc(od, x if src is None else outs[src], w, b, s)
    net.set_input(0, x)
    for _ in range(2):
        net.run()
    torch.cuda.synchronize()
    class _DevBuf:     # the layer's resident output buffer, viewed through the CUDA array interface
        def __init__(self, ptr, shape):
            self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "|i1", "data": (ptr, False), "version": 3}

    for i, name in enumerate(names):
        _, y_ptr = net.layer_io(i)
        want = outs[name]
        got = torch.as_tensor(_DevBuf(y_ptr, want.shape), device="cuda:0").cpu().numpy()
        assert np.array_equal(got, want), name
    net.close()
