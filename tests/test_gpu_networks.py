"""GPU parity of the benchmark configurations AS CONFIGURED (BASELINE.json configs 2-5): every network table of
lowbitdnn_project_b200.networks is built through lbc.Net (the object bench.py times) at a small batch, run twice
back to back, and EVERY layer's resident output is compared bit for bit with the CPU oracle chained over the same
graph.  This covers what the single-layer tests cannot: the planner's real per-layer choices on the real shapes, the
alternating traversal direction along producer -> consumer edges, programmatic dependent launch across 53 layers,
and buffers handed from one kernel family to the next."""
import numpy as np
import pytest

from oracle import oracle
from oracle.oracle import ConvDesc as OD

pytestmark = pytest.mark.gpu


def synth_params(d, layer):
    """Same generator as bench.py."""
    cg = d.c // d.groups
    rw = np.random.default_rng(4321 + layer)
    w = rw.integers(-127, 128, size=(d.k, d.r, d.s, cg), dtype=np.int8)
    bias = rw.integers(-2**15, 2**15, size=(d.k,), dtype=np.int32)
    scale = (rw.uniform(0.5, 2.0, size=(d.k,)) * 2.0**-7 / np.sqrt(d.r * d.s * cg)).astype(np.float32)
    return w, bias, scale


def synth_input(d, layer):
    return np.random.default_rng(1234 + layer).integers(-128, 128, size=(d.n, d.h, d.w, d.c), dtype=np.int8)


def oracle_chain(layers):
    """{layer name: oracle output} for a [(name, ConvDesc, input_name_or_None)] table (same inputs as load_net)."""
    out = {}
    for i, (name, d, src) in enumerate(layers):
        x = synth_input(d, i) if src is None else out[src]
        w, b, s = synth_params(d, i)
        out[name] = oracle.conv_nhwc(OD(**d.__dict__), x, w, b, s)
    return out


def load_net(lbc, layers, options=None):
    net = lbc.Net(layers, options=options)
    for i, (_, d, src) in enumerate(layers):
        net.set_params(i, *synth_params(d, i))
        if src is None:
            net.set_input(i, synth_input(d, i))
    return net


def compare(net, layers, want, tag):
    for i, (name, _, _) in enumerate(layers):
        if net.fused_into(i) >= 0:
            continue                 # absorbed by its consumer's fused launch: the consumer's output covers it
        got = net.read_output(i)
        if not np.array_equal(got, want[name]):
            from tests.parity_util import mismatch_report
            raise AssertionError(f"{tag}: layer {i} {name} [{net.layer_describe(i)}]\n{mismatch_report(got, want[name])}")


# batch sizes: small enough for the oracle (seconds), large enough that every layer has several tiles per CTA row
CASES = [("resnet50", 4), ("resnet18", 4), ("vgg16", 2), ("mobilenet_v2", 8), ("single_3x3", 1)]


@pytest.mark.parametrize("name,batch", CASES, ids=[c[0] for c in CASES])
def test_network_matches_oracle_chain(name, batch):
    import torch
    import lowbitdnn_project_b200 as lbc
    layers = lbc.networks.NETWORKS[name](batch)
    want = oracle_chain(layers)
    net = load_net(lbc, layers)
    assert all(net.fused_into(i) == -1 for i in range(len(layers)))      # fusion is opt-in
    stream = torch.cuda.current_stream()
    # two untimed runs back to back (launches overlap through programmatic dependent launch), then a timed one
    net.run(stream=stream)
    net.run(stream=stream)
    torch.cuda.synchronize()
    net.check_status()
    compare(net, layers, want, f"{name} N={batch} (2 runs)")
    per, tot = net.run(stream=stream, timed=True)
    assert len(per) == len(layers) and tot > 0
    compare(net, layers, want, f"{name} N={batch} (timed run)")
    net.close()


@pytest.mark.parametrize("name,batch", [("resnet50", 3), ("mobilenet_v2", 5)], ids=["resnet50", "mobilenet_v2"])
def test_network_without_alternating_traversal_and_capped_grid(name, batch):
    """The same graphs with the traversal alternation off and the persistent grids capped to 5 CTAs, so every CTA walks
    many tiles of every layer (what a batch-512 launch does on 148 SMs), odd batch sizes included."""
    import torch
    import lowbitdnn_project_b200 as lbc
    layers = lbc.networks.NETWORKS[name](batch)
    want = oracle_chain(layers)
    for options in ({"reverse": 0, "max_grid": 5}, {"max_grid": 6}, {"fuse": 1}, {"fuse": 1, "max_grid": 7, "reverse": 0}):
        net = load_net(lbc, layers, options=options)
        if name == "resnet50" and options.get("fuse"):
            # stage 1's conv2 -> conv3 pairs run as fused launches (the middle tensor is never written)
            names = [l[0] for l in layers]
            for b in range(3):
                assert net.fused_into(names.index(f"l1.{b}.conv2")) == names.index(f"l1.{b}.conv3")
            assert net.fused_into(names.index("l2.1.conv2")) == -1
            assert net.launches == len(layers) + 1 - 3           # the stem launches twice, three pairs launch once
        net.run(stream=torch.cuda.current_stream())
        net.run(stream=torch.cuda.current_stream())
        torch.cuda.synchronize()
        net.check_status()
        compare(net, layers, want, f"{name} N={batch} {options}")
        net.close()


def test_network_host_paths_match_oracle():
    """lbc_net_run_host (blocking) and lbc_net_submit_host / sync_host (pipelined, double-buffered input) return the
    oracle's bytes for the last layer, for several different inputs in flight."""
    import torch
    import lowbitdnn_project_b200 as lbc
    layers = lbc.networks.resnet18(2)
    # make the graph a single chain from the network input so that the host input determines the output
    chain = [l for l in layers if "downsample" not in l[0] and l[0] != "conv1"]     # (a max-pool sits behind conv1)
    chain = [(n, d, (chain[i - 1][0] if i else None)) for i, (n, d, _) in enumerate(chain)]
    net = load_net(lbc, chain)
    d0, dl = chain[0][1], chain[-1][1]
    p, q = dl.out_hw
    xs = [np.random.default_rng(50 + j).integers(-128, 128, size=(d0.n, d0.h, d0.w, d0.c), dtype=np.int8) for j in range(3)]
    wants = []
    for x in xs:
        cur = x
        for i, (_, d, _) in enumerate(chain):
            cur = oracle.conv_nhwc(OD(**d.__dict__), cur, *synth_params(d, i))
        wants.append(cur)
    stream = torch.cuda.current_stream()
    xh = [torch.from_numpy(x).pin_memory() for x in xs]
    yh = [torch.empty((dl.n, p, q, dl.k), dtype=torch.int8).pin_memory() for _ in xs]
    for j in range(3):
        net.run_host(xh[j], yh[j], stream=stream)
        assert np.array_equal(yh[j].numpy(), wants[j]), f"blocking host path, input {j}"
        yh[j].zero_()
    for rounds in range(2):
        for j in range(3):
            net.submit_host(xh[j], yh[j], stream=stream)
        net.sync_host()
        for j in range(3):
            assert np.array_equal(yh[j].numpy(), wants[j]), f"pipelined host path, round {rounds}, input {j}"
            yh[j].zero_()
    # the resident-input path still works after the pipelined one (layer 0 goes back to its own buffer)
    net.set_input(0, xs[1])
    net.run(stream=stream)
    assert np.array_equal(net.read_output(len(chain) - 1), wants[1])
    net.close()


def test_weights_replaced_between_runs_take_effect():
    """Networks fetch resident filter matrices before the programmatic-dependency wait (lbc_plan_options.early_weights):
    legal because nothing in the stream writes packed weights - lbc_net_set_params is device-synchronous.  Replace the
    parameters of every layer between back-to-back runs and check that the next run sees exactly the new ones, with the
    early fetch on (default) and off."""
    import torch
    import lowbitdnn_project_b200 as lbc
    layers = [l for l in lbc.networks.resnet50(2) if l[0].startswith(("l1.0", "l2.0")) and "downsample" not in l[0]]
    layers = [(n, d, (layers[i - 1][0] if i else None)) for i, (n, d, _) in enumerate(layers)]
    layers = [l for l in layers[:3]]                       # l1.0.conv1 -> conv2 -> conv3: all resident-filter layers
    # ... and two layers that stream their filter matrix through the ring (256 -> 256 3x3 in CTA pairs, 256 -> 512 1x1)
    CD = lbc.ConvDesc
    layers += [("s.conv2", CD(n=2, h=56, w=56, c=256, k=256, r=3, s=3, stride_h=2, stride_w=2, pad_h=1, pad_w=1, relu=1), "l1.0.conv3"),
               ("s.conv3", CD(n=2, h=28, w=28, c=256, k=512, r=1, s=1, relu=1), "s.conv2")]
    stream = torch.cuda.current_stream()
    for options in (None, {"early_weights": 2}, {"early_weights": 0}):
        net = load_net(lbc, layers, options=options)
        for gen in range(3):
            if gen:
                for i, (_, d, _) in enumerate(layers):
                    net.set_params(i, *synth_params(d, i + 100 * gen))
            net.run(stream=stream)
            net.run(stream=stream)
            torch.cuda.synchronize()
            net.check_status()
            out = {}
            for i, (name, d, src) in enumerate(layers):
                x = synth_input(d, i) if src is None else out[src]
                out[name] = oracle.conv_nhwc(OD(**d.__dict__), x, *synth_params(d, i + 100 * gen))
            compare(net, layers, out, f"generation {gen}, options {options}")
        net.close()
