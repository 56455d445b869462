"""GPU parity of the int8 ops between convolutions (max-pool, residual add + ReLU, global average pool) and of whole
int8 graphs (conv -> pool -> bottlenecks with residual joins) against the numpy/C oracle, bit for bit."""
import numpy as np
import pytest

from oracle import oracle
from oracle.oracle import ConvDesc as OD

pytestmark = pytest.mark.gpu


POOL_CASES = [
    # (n, h, w, c, kernel, stride, pad)
    (2, 112, 112, 64, 3, 2, 1),       # ResNet stem pool
    (3, 28, 28, 128, 2, 2, 0),        # VGG pool; tmp.py:44 qmax_pool2d(x, (2,2), (2,2), (0,0))
    (2, 13, 17, 32, 3, 1, 0),         # tmp.py:52 (3,3),(1,1),(0,0); odd sizes
    (1, 9, 9, 16, (3, 2), (2, 1), (1, 0)),
    (2, 7, 7, 24, 3, 2, 1),           # C % 16 != 0: scalar kernel
    (1, 5, 6, 3, 2, 2, 0),
    (2, 8, 8, 48, 8, 8, 0),           # one window per image
]


@pytest.mark.parametrize("case", POOL_CASES, ids=lambda c: f"h{c[1]}w{c[2]}c{c[3]}k{c[4]}s{c[5]}p{c[6]}")
def test_max_pool_matches_oracle(case):
    import torch
    import lowbitdnn_project_b200 as lbc
    n, h, w, c, k, s, p = case
    pair = lambda v: (v, v) if isinstance(v, int) else v
    (kh, kw), (sh, sw), (ph, pw) = pair(k), pair(s), pair(p)
    x = np.random.default_rng(3).integers(-128, 128, size=(n, h, w, c), dtype=np.int8)
    want = oracle.max_pool_nhwc(x, kh, kw, sh, sw, ph, pw)
    got = lbc.max_pool2d(torch.from_numpy(x).cuda(), k, s, p).cpu().numpy()
    assert got.shape == want.shape and np.array_equal(got, want)
    # all-minimum input: padding (which never wins) must not leak a different value
    x[:] = -128
    got = lbc.max_pool2d(torch.from_numpy(x).cuda(), k, s, p).cpu().numpy()
    assert (got == -128).all()


@pytest.mark.parametrize("numel", [16, 4096 + 7, 3 * 56 * 56 * 256, 5])
@pytest.mark.parametrize("relu", [True, False])
def test_add_relu_matches_oracle(numel, relu):
    import torch
    import lowbitdnn_project_b200 as lbc
    rng = np.random.default_rng(numel)
    a = rng.integers(-128, 128, size=(numel,), dtype=np.int8)
    b = rng.integers(-128, 128, size=(numel,), dtype=np.int8)
    a[:4] = [127, -128, 127, -128]
    b[:4] = [127, -128, -128, 127]                      # both saturations and exact cancellation
    got = lbc.add_relu(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda(), relu).cpu().numpy()
    assert np.array_equal(got, oracle.add_relu(a, b, relu))
    # unaligned views take the scalar path
    if numel > 64:
        ta, tb = torch.from_numpy(a).cuda()[3:], torch.from_numpy(b).cuda()[3:]
        got = lbc.add_relu(ta.contiguous(), tb.contiguous(), relu).cpu().numpy()
        assert np.array_equal(got, oracle.add_relu(a[3:], b[3:], relu))


@pytest.mark.parametrize("shape", [(4, 7, 7, 2048), (3, 5, 9, 100), (2, 1, 1, 16)])
def test_global_avg_pool_matches_oracle(shape):
    import torch
    import lowbitdnn_project_b200 as lbc
    x = np.random.default_rng(9).integers(-128, 128, size=shape, dtype=np.int8)
    scale = 1.0 / (shape[1] * shape[2]) * 1.7
    got = lbc.global_avg_pool(torch.from_numpy(x).cuda(), scale).cpu().numpy()
    assert np.array_equal(got, oracle.global_avg_pool(x, scale))


def _graph_oracle(layers):
    from tests.test_gpu_networks import synth_input, synth_params
    import lowbitdnn_project_b200 as lbc
    out = {}
    for i, (name, d, src) in enumerate(layers):
        if isinstance(d, lbc.AddDesc):
            out[name] = oracle.add_relu(out[src[0]], out[src[1]], bool(d.relu))
        elif isinstance(d, lbc.PoolDesc):
            out[name] = oracle.max_pool_nhwc(out[src], d.kh, d.kw, d.stride_h, d.stride_w, d.pad_h, d.pad_w)
        else:
            x = synth_input(d, i) if src is None else out[src]
            out[name] = oracle.conv_nhwc(OD(**d.__dict__), x, *synth_params(d, i))
    return out


@pytest.mark.parametrize("name,batch", [("resnet50_full", 3), ("resnet18_full", 4), ("vgg16_full", 2)],
                         ids=["resnet50_full", "resnet18_full", "vgg16_full"])
def test_whole_int8_graph_matches_oracle(name, batch):
    """conv1 -> max-pool -> residual blocks as ONE device graph (lbc_net_create_graph): every node's output, twice."""
    import torch
    import lowbitdnn_project_b200 as lbc
    from tests.test_gpu_networks import synth_input, synth_params
    layers = lbc.networks.NETWORKS[name](batch)
    want = _graph_oracle(layers)
    net = lbc.Net(layers)
    for i, (_, d, src) in enumerate(layers):
        if isinstance(d, lbc.ConvDesc):
            net.set_params(i, *synth_params(d, i))
            if src is None:
                net.set_input(i, synth_input(d, i))
    kinds = {net.layer_kernel(i) for i in range(len(layers))}
    assert "maxpool" in kinds and ("add_relu" in kinds or name == "vgg16_full")
    for rep in range(2):
        net.run(stream=torch.cuda.current_stream())
    torch.cuda.synchronize()
    net.check_status()
    for i, (lname, _, _) in enumerate(layers):
        if net.fused_into(i) >= 0:
            continue
        assert np.array_equal(net.read_output(i), want[lname]), f"{name}: node {i} {lname} ({net.layer_kernel(i)})"
    # the host path: network input in, last node out
    d0 = layers[0][1]
    x = synth_input(d0, 0)
    xh = torch.from_numpy(x).pin_memory()
    last = layers[-1][1]
    p, q = last.out_hw
    yh = torch.empty((batch, p, q, last.k if isinstance(last, lbc.ConvDesc) else last.c), dtype=torch.int8).pin_memory()
    net.run_host(xh, yh, stream=torch.cuda.current_stream())
    assert np.array_equal(yh.numpy(), want[layers[-1][0]])
    net.close()
