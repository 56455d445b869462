"""GPU tests of the backward int8 convolutions and of the qtorch-compatible module (8f-3).

Gradients are int32-exact: dgrad / wgrad through the C ABI equal the direct definitions of the sums (numpy, int64),
and QConv2D's forward / backward equal float64 convolutions of the quantized integer tensors up to the final fp32
scale multiply.  The last test is the reference's own fixture matrix (python/qtorch/tests/conftest.py:10-67), repaired,
thinned out and run against nn.Conv2d within the quantisation error."""
import itertools

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def np_dgrad(dy, w, pad, h, wd):
    """dx[n,h,w,c] = sum_{k,r,s} dy[n,h+pad-r,w+pad-s,k] * w[k,r,s,c]   (conv2DBackwardData3x3.cuh:61-64)."""
    n, p, q, k = dy.shape
    _, r, s, c = w.shape
    dx = np.zeros((n, h, wd, c), dtype=np.int64)
    dy64, w64 = dy.astype(np.int64), w.astype(np.int64)
    for a in range(r):
        for b in range(s):
            # dy row index = hh + pad - a must lie in [0, p)
            h0, h1 = max(0, a - pad), min(h, p + a - pad)
            w0, w1 = max(0, b - pad), min(wd, q + b - pad)
            if h0 >= h1 or w0 >= w1:
                continue
            dx[:, h0:h1, w0:w1, :] += dy64[:, h0 + pad - a:h1 + pad - a, w0 + pad - b:w1 + pad - b, :] @ w64[:, a, b, :]
    return dx.astype(np.int32)


def np_wgrad(x, dy, pad, r, s):
    """dw[k,r,s,c] = sum_{n,p,q} dy[n,p,q,k] * x[n,p-pad+r,q-pad+s,c]   (conv2DBackwardWeights3x3.cuh:15-100)."""
    n, h, w, c = x.shape
    _, p, q, k = dy.shape
    xp = np.zeros((n, h + 2 * pad, w + 2 * pad, c), dtype=np.int64)
    xp[:, pad:pad + h, pad:pad + w, :] = x
    dy64 = dy.astype(np.int64).reshape(-1, k)
    dw = np.zeros((k, r, s, c), dtype=np.int64)
    for a in range(r):
        for b in range(s):
            dw[:, a, b, :] = dy64.T @ xp[:, a:a + p, b:b + q, :].reshape(-1, c)
    return dw.astype(np.int32)


BWD_CASES = [
    # (n, h, w, c, k, r, pad)
    (16, 14, 14, 32, 32, 3, 1),
    (32, 8, 8, 64, 16, 3, 1),
    (16, 10, 12, 16, 48, 1, 0),
    (4, 9, 9, 12, 20, 3, 0),       # odd channel counts: CUDA-core path
    (16, 12, 12, 16, 16, 5, 2),
    (3, 7, 7, 8, 8, 3, 2),         # padding == filter - 1
]


@pytest.mark.parametrize("case", BWD_CASES, ids=lambda c: "n%dh%dc%dk%dr%dp%d" % (c[0], c[1], c[3], c[4], c[5], c[6]))
def test_backward_convolutions_are_int32_exact(case):
    import torch
    import lowbitdnn_project_b200 as lbc
    n, h, w, c, k, r, pad = case
    d = lbc.ConvDesc(n=n, h=h, w=w, c=c, k=k, r=r, s=r, pad_h=pad, pad_w=pad)
    p, q = d.out_hw
    rng = np.random.default_rng(17)
    x = rng.integers(-128, 128, size=(n, h, w, c), dtype=np.int8)
    wt = rng.integers(-128, 128, size=(k, r, r, c), dtype=np.int8)
    dy = rng.integers(-128, 128, size=(n, p, q, k), dtype=np.int8)
    dev = torch.device("cuda:0")
    dx = lbc.conv_backward_data(d, torch.from_numpy(dy).to(dev), torch.from_numpy(wt).to(dev).reshape(-1))
    dw = lbc.conv_backward_weights(d, torch.from_numpy(x).to(dev), torch.from_numpy(dy).to(dev))
    torch.cuda.synchronize()
    assert np.array_equal(dx.cpu().numpy(), np_dgrad(dy, wt, pad, h, w))
    assert np.array_equal(dw.cpu().numpy(), np_wgrad(x, dy, pad, r, r))


def test_backward_refuses_what_the_reference_refuses():
    import lowbitdnn_project_b200 as lbc
    import torch
    d = lbc.ConvDesc(n=2, h=8, w=8, c=16, k=16, r=3, s=3, stride_h=2, stride_w=2, pad_h=1, pad_w=1)
    with pytest.raises(lbc.LbcError):
        lbc.conv_backward_data(d, torch.zeros((2, 4, 4, 16), dtype=torch.int8, device="cuda"),
                               torch.zeros(16 * 9 * 16, dtype=torch.int8, device="cuda"))


@pytest.mark.parametrize("shape", [(2, 16, 12, 12, 32, 3, 1), (3, 64, 9, 11, 16, 1, 0), (1, 12, 6, 3, 20, 3, 1)],
                         ids=["c16k32r3", "c64k16r1", "c12k20r3"])
def test_qconv2d_forward_and_backward_equal_float64_on_the_quantized_tensors(shape):
    import torch
    import torch.nn.functional as F
    import lowbitdnn_project_b200  # noqa: F401
    import lowbitdnn_project_b200.qtorch as qt
    n, c, h, w, k, r, pad = shape
    torch.manual_seed(4)
    dev = torch.device("cuda:0")
    x = torch.randn(n, c, h, w, device=dev, requires_grad=True)
    conv = qt.QConv2D(c, k, r, stride=1, padding=pad).to(dev)
    y = conv(x)
    g = torch.randn_like(y)
    y.backward(g)
    # the same arithmetic in float64 on the same integer tensors
    xq, wq, gq = qt.quantize(x.detach().clone()), qt.quantize(conv.weight.detach().clone()), qt.quantize(g.clone())
    si, sw, sg = (qt.QUANTIZATION_PARAMETERS[t] for t in (xq, wq, gq))
    xd, wd, gd = xq.double().requires_grad_(True), wq.double().requires_grad_(True), gq.double()
    yd = F.conv2d(xd, wd, padding=pad)
    assert torch.equal(y.detach(), (yd.detach().round().to(torch.int64).float() * (sw * si)))
    yd.backward(gd)
    assert torch.allclose(x.grad, xd.grad.float() * (sg * sw), rtol=1e-6, atol=0)
    assert torch.allclose(conv.weight.grad, wd.grad.float() * (si * sg), rtol=1e-6, atol=0)
    # and it approximates the fp32 layer within the quantisation error
    yf = F.conv2d(x.detach(), conv.weight.detach(), padding=pad)
    assert (y.detach() - yf).abs().max() <= 0.05 * yf.abs().max() + 1e-3


def test_reference_fixture_matrix_forward_against_nn_conv2d():
    """python/qtorch/tests/conftest.py:10-67 (batch x channels x size x kernel x stride x padding x filters), every 13th
    combination: QConv2D's output stays within the quantisation error of nn.Conv2d with the same weights."""
    import torch
    import lowbitdnn_project_b200  # noqa: F401
    import lowbitdnn_project_b200.qtorch as qt
    batches, channels, sizes = (1, 3), (12, 64, 512), ((1, 1), (1, 5), (6, 3), (64, 112), (224, 224))
    kernels, strides, paddings, filters = ((1, 1), (3, 3)), (1, 2), ((1, 2), (3, 3), 0), (4, 20, 32, 128)
    combos = list(itertools.product(batches, channels, sizes, kernels, strides, paddings, filters))
    torch.manual_seed(1)
    dev = torch.device("cuda:0")
    ran = 0
    for i, (b, c, (h, w), kern, stride, pad, k) in enumerate(combos):
        if i % 13:
            continue
        ph, pw = (pad, pad) if isinstance(pad, int) else pad
        if h + 2 * ph < kern[0] or w + 2 * pw < kern[1] or (h * w >= 224 * 224 and c >= 512):
            continue
        x = torch.randn(b, c, h, w, device=dev)
        conv = qt.QConv2D(c, k, kern, stride=stride, padding=pad).to(dev)
        with torch.no_grad():
            y = conv(x)
            ref = torch.nn.functional.conv2d(x, conv.weight, None, stride, pad)
        assert y.shape == ref.shape
        # per-tensor 8-bit quantisation of both operands: error ~ sqrt(taps) * (dx*|w| + dw*|x|); generous bound
        taps = c * kern[0] * kern[1]
        bound = 4.0 * np.sqrt(taps) * (x.abs().max() * conv.weight.abs().max()).item() / 127.0
        assert (y - ref).abs().max().item() <= bound, (b, c, h, w, kern, stride, pad, k)
        ran += 1
    assert ran >= 25


def test_qmax_pool2d_on_vect_c_tensors():
    import torch
    import lowbitdnn_project_b200  # noqa: F401
    import lowbitdnn_project_b200.qtorch as qt
    x = torch.randint(-128, 128, (2, 8, 12, 12), dtype=torch.int8, device="cuda")
    v = qt.to_vect_c(x, contiguous=True)
    y = qt.qmax_pool2d(v, (2, 2), (2, 2), (0, 0))
    want = torch.nn.functional.max_pool2d(x.float(), 2, 2).to(torch.int8)
    assert torch.equal(qt.from_vect_c(y), want)
