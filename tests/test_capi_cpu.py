"""CPU tests: the C-ABI library builds, loads, exports every symbol include/lowbit_cnn.h declares,
does its host arithmetic right, and FAILS LOUDLY (no fallback) when there is no GPU."""
import os
import re

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


@pytest.fixture(scope="module")
def lbc():
    import lowbitdnn_project_b200 as m
    import importlib.util
    spec = importlib.util.spec_from_file_location("lbc_build", os.path.join(ROOT, "lowbitdnn-project_b200", "build.py"))
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    b.build()
    m.load_library()
    return m


def test_header_symbols_are_exported(lbc):
    hdr = open(os.path.join(ROOT, "include", "lowbit_cnn.h")).read()
    declared = set(re.findall(r"\b(lbc_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 30
    lib = lbc.load_library()
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/lowbit_cnn.h but not exported"
    from lowbitdnn_project_b200 import _capi
    assert declared == set(_capi.EXPORTED_SYMBOLS), declared ^ set(_capi.EXPORTED_SYMBOLS)


def test_host_arithmetic(lbc):
    d = lbc.ConvDesc(n=1, h=56, w=56, c=64, k=64, r=3, s=3, pad_h=1, pad_w=1)
    assert d.out_hw == (56, 56)
    ops, byts = d.work
    assert ops == 2 * 115605504 and byts == 438784
    d = lbc.ConvDesc(n=512, h=224, w=224, c=3, k=64, r=7, s=7, stride_h=2, stride_w=2, pad_h=3, pad_w=3)
    assert d.out_hw == (112, 112)
    with pytest.raises(lbc.LbcError):
        _ = lbc.ConvDesc(n=1, h=2, w=2, c=4, k=4, r=5, s=5).out_hw      # filter larger than input
    with pytest.raises(lbc.LbcError):
        _ = lbc.ConvDesc(n=1, h=8, w=8, c=6, k=4, r=1, s=1, groups=4).out_hw


def test_network_tables_match_survey(lbc):
    nets = lbc.networks
    want = {"resnet18": (20, 464.27, 1.206), "resnet50": (53, 2092.61, 11.173), "vgg16": (13, 1964.37, 2.911),
            "mobilenet_v2": (52, 306.68, 13.769), "single_3x3": (1, 0.115605504, 0.000438784)}
    for name, (n_layers, gmac, gb) in want.items():
        layers = nets.NETWORKS[name](nets.DEFAULT_BATCH[name])
        g, b = nets.total_work(layers)
        assert len(layers) == n_layers
        assert abs(g - gmac) / gmac < 1e-3 and abs(b - gb) / gb < 2e-3, (name, g, b)
        names = [l[0] for l in layers]
        for _, _, src in layers:
            assert src is None or src in names


def test_no_silent_fallback_without_gpu(lbc):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(lbc.LbcError) as e:
        lbc.ConvPlan(lbc.ConvDesc(n=1, h=8, w=8, c=16, k=16, r=1, s=1))
    assert e.value.status == 3  # LBC_ERR_NO_DEVICE


def test_product_path_does_not_import_oracle():
    """The product path may never import, include, link or dlopen anything under oracle/."""
    pkg = os.path.join(ROOT, "lowbitdnn-project_b200")
    pat = re.compile(r"import\s+oracle|from\s+oracle|from\s+\.+oracle|oracle/|libcpu_ref|libref_conv|#include[^\n]*oracle")
    for root, _, files in os.walk(pkg):
        if os.path.basename(root) in ("build", "lib", "__pycache__"):
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                src = open(os.path.join(root, f)).read()
                assert not pat.search(src), f"{os.path.join(root, f)} reaches into oracle/"


def test_planner_dry_run_covers_every_benchmark_layer(lbc):
    """Every layer of the five benchmark configurations gets a plan (tile, stages, shared memory all fit) and the kernel
    family the design assigns to it - checked through the host-only dry planner, no GPU needed."""
    import ctypes
    from collections import Counter
    lib = lbc.load_library()
    want = {"resnet50": {"igemm_tc": 52, "stem_tc": 1}, "resnet18": {"igemm_tc": 19, "stem_tc": 1},
            "vgg16": {"igemm_tc": 12, "stem_tc": 1}, "mobilenet_v2": {"igemm_tc": 34, "depthwise": 17, "stem_tc": 1},
            "single_3x3": {"igemm_tc": 1}}
    for net, kinds in want.items():
        got = Counter()
        for name, d, _ in lbc.networks.NETWORKS[net](lbc.networks.DEFAULT_BATCH[net]):
            kind = ctypes.c_int32()
            buf = ctypes.create_string_buffer(512)
            cd = d.c_struct()
            st = lib.lbc_conv_plan_dry(ctypes.byref(cd), 0, 148, ctypes.byref(kind), buf, 512)
            assert st == 0, (net, name, lib.lbc_last_error_string().decode())
            got[lbc._capi.KERNEL_NAMES[kind.value]] += 1
            text = buf.value.decode()
            if "smem" in text:
                assert int(text.split("smem ")[1].split()[0]) <= 227 * 1024, text
        assert dict(got) == kinds, (net, dict(got))


def test_planner_choices_on_resnet50(lbc):
    """The planner's per-layer decisions that DESIGN.md section 4.1 describes, pinned through the host-only dry planner:
    which layers run in CTA pairs, which keep their filter matrix resident (also across two N tiles), where the bias is
    folded into the MMA, where warps store their own rows, and which A mode the 3x3 layers use."""
    import ctypes
    lib = lbc.load_library()
    plans = {}
    for name, d, _ in lbc.networks.NETWORKS["resnet50"](512):
        kind = ctypes.c_int32()
        buf = ctypes.create_string_buffer(512)
        cd = d.c_struct()
        assert lib.lbc_conv_plan_dry(ctypes.byref(cd), 0, 148, ctypes.byref(kind), buf, 512) == 0
        plans[name] = buf.value.decode()
    expect = {
        "conv1": ["stem_tc", "b=resident,2mma", "a=window", "4x4-warp-teams"],
        "l1.0.conv2": ["b=resident,2mma", "a=window(2x56", "tmem 512", "4x4-warp-teams"],            # 8 stages of 64 columns
        "l1.0.conv3": ["b=resident", "a=tiled", "warp-stores,bias-in-mma"],                          # 64 -> 256
        "l1.1.conv1": ["b=resident", "a=tiled", "tile 128x64"],                                      # 256 -> 64
        "l2.0.conv1": ["b=resident", "bias-in-mma", "tile 128x128"],                                 # 256 -> 128
        "l2.0.conv3": ["b=resident", "tiles 3136x2", "warp-stores,bias-in-mma"],                     # 128 -> 512, two N tiles
        "l2.1.conv2": ["b=resident,cta-pair", "a=window(4x28", "tile 128x128"],                      # 3x3, 128-wide: windows, filter halves resident
        "l2.0.conv2": ["b=ring,cta-pair", "a=im2col"],                                               # its stride-2 sibling streams (measured)
        "l3.0.conv1": ["b=ring ", "a=tiled", "tile 128x256"],                                        # 512 -> 256
        "l3.0.conv3": ["b=resident,n-stationary", "a=tiled", "tiles 784x4", "grid 148", "warp-stores,bias-in-mma"],   # 256 -> 1024: 64 KB tiles
        "l2.0.downsample": ["b=resident,n-stationary", "a=im2col", "tiles 3136x2"],                  # 256 -> 512, stride 2
        "l3.0.downsample": ["b=ring ", "a=im2col", "tiles 784x4"],                                   # 512 -> 1024: 128 KB tiles stream
        "l3.1.conv1": ["b=ring,cta-pair", "a=tiled"],                                                # 1024 -> 256: long K loop
        "l3.1.conv2": ["b=ring,cta-pair", "a=im2col", "tile 128x256", "2x8-warp-teams,tail-split"],  # wide 3x3 in pairs: im2col; 392 steps on 74 pairs
        "l3.0.conv2": ["b=ring,cta-pair", "a=im2col"],                                               # stride 2
        "l4.1.conv2": ["b=ring,cta-pair", "a=im2col", "2x8-warp-teams"],                             # two N tiles: no split last round
        "l4.0.conv3": ["b=ring ", "tiles 196x8", "2x8-warp-teams"],                                  # 512 -> 2048
    }
    for name, needles in expect.items():
        for n in needles:
            assert n in plans[name], (name, n, plans[name])
    # every CTA-pair plan has an even grid; nothing exceeds the shared-memory or TMEM limits
    for name, text in plans.items():
        if "cta-pair" in text:
            assert int(text.split("grid ")[1].split()[0]) % 2 == 0, text
        assert int(text.split("tmem ")[1].split()[0]) <= 512, text


def test_plan_options_struct_is_checked_and_steers_the_planner(lbc):
    """lbc_plan_options (the replacement of round 1's environment variables): the ctypes mirror has the header's size,
    a wrong struct_size is refused, and the tri-states change the dry planner's answer the way the header says."""
    import ctypes
    from lowbitdnn_project_b200 import _capi
    lib = lbc.load_library()
    o = _capi.CPlanOptions()
    lib.lbc_plan_options_init(ctypes.byref(o))
    assert o.struct_size == ctypes.sizeof(_capi.CPlanOptions)
    hdr = open(os.path.join(ROOT, "include", "lowbit_cnn.h")).read()
    body = hdr[hdr.index("typedef struct lbc_plan_options"):hdr.index("} lbc_plan_options;")]
    n_fields = len(re.findall(r"^\s*int32_t\s+[a-z_0-9]+;", body, re.M)) + sum(int(m) for m in re.findall(r"int32_t\s+reserved\[(\d+)\];", body))
    assert n_fields * 4 == ctypes.sizeof(_capi.CPlanOptions), (n_fields, ctypes.sizeof(_capi.CPlanOptions))
    assert o.cta_pairs == -1 and o.fuse == -1 and o.early_weights == -1 and o.max_grid == 0       # "planner decides" / "default"

    def dry(d, **kw):
        opt = _capi.plan_options(**kw)
        kind, buf, cd = ctypes.c_int32(), ctypes.create_string_buffer(512), d.c_struct()
        st = lib.lbc_conv_plan_dry_ex(ctypes.byref(cd), 0, ctypes.byref(opt), 148, ctypes.byref(kind), buf, 512)
        return st, buf.value.decode()

    d = lbc.ConvDesc(n=512, h=28, w=28, c=128, k=128, r=3, s=3, pad_h=1, pad_w=1, relu=1)
    assert "b=resident,cta-pair" in dry(d)[1]
    assert "b=ring,cta-pair" in dry(d, resident_filter=2)[1]
    assert "cta-pair" not in dry(d, cta_pairs=0)[1]
    assert "a=im2col" in dry(d, force_im2col=1)[1]
    assert "grid 6 " in dry(d, max_grid=6)[1]
    d64 = lbc.ConvDesc(n=512, h=56, w=56, c=64, k=64, r=3, s=3, pad_h=1, pad_w=1, relu=1)
    assert "4x4-warp-teams" in dry(d64)[1] and "narrow-warp-stores" in dry(d64, warp_store=1)[1]
    bad = _capi.plan_options()
    bad.struct_size = 8
    kind, buf, cd = ctypes.c_int32(), ctypes.create_string_buffer(512), d.c_struct()
    assert lib.lbc_conv_plan_dry_ex(ctypes.byref(cd), 0, ctypes.byref(bad), 148, ctypes.byref(kind), buf, 512) != 0
    assert b"struct_size" in lib.lbc_last_error_string()
