"""Layer lists of the benchmark configurations (BASELINE.json `configs`).

Each entry is (name, ConvDesc-kwargs, input_of) where input_of names the layer whose output feeds it
(None = fed from a resident synthetic activation buffer: network input, or a point where the real
topology has a non-conv op — max-pool / residual add — between two convolutions).
torchvision topologies; ResNet-50 is v1.5 (stride on the 3x3).  All convs: bias + per-channel scale + ReLU.
"""
from __future__ import annotations

from .conv import ConvDesc
from .ops import AddDesc, PoolDesc


def _cd(n, h, c, k, r, stride=1, pad=None, groups=1, relu=1):
    pad = (r // 2) if pad is None else pad
    return ConvDesc(n=n, h=h, w=h, c=c, k=k, r=r, s=r, stride_h=stride, stride_w=stride, pad_h=pad, pad_w=pad,
                    groups=groups, relu=relu)


def single_3x3(n=1):
    """Config 1: N1, 56x56x64 -> 64, 3x3 s1 p1 (the only config the reference CPU path can run in full)."""
    return [("conv3x3_56_64", _cd(n, 56, 64, 64, 3), None)]


def resnet50(n=512):
    L = [("conv1", _cd(n, 224, 3, 64, 7, stride=2, pad=3), None)]
    h, cin = 56, 64
    prev = None  # after max-pool: synthetic buffer
    for stage, (mid, blocks) in enumerate(((64, 3), (128, 4), (256, 6), (512, 3)), start=1):
        out = mid * 4
        for b in range(blocks):
            stride = 2 if (b == 0 and stage > 1) else 1
            pre = f"l{stage}.{b}"
            L.append((pre + ".conv1", _cd(n, h, cin, mid, 1), prev))
            L.append((pre + ".conv2", _cd(n, h, mid, mid, 3, stride=stride), pre + ".conv1"))
            h2 = h // stride
            L.append((pre + ".conv3", _cd(n, h2, mid, out, 1), pre + ".conv2"))
            if b == 0:
                L.append((pre + ".downsample", _cd(n, h, cin, out, 1, stride=stride, relu=0), prev))
            prev = pre + ".conv3"   # (residual add + ReLU happen outside the conv path)
            h, cin = h2, out
    return L


def resnet18(n=256):
    L = [("conv1", _cd(n, 224, 3, 64, 7, stride=2, pad=3), None)]
    h, cin = 56, 64
    prev = None
    for stage, ch in enumerate((64, 128, 256, 512), start=1):
        for b in range(2):
            stride = 2 if (b == 0 and stage > 1) else 1
            pre = f"l{stage}.{b}"
            L.append((pre + ".conv1", _cd(n, h, cin, ch, 3, stride=stride), prev))
            h2 = h // stride
            L.append((pre + ".conv2", _cd(n, h2, ch, ch, 3), pre + ".conv1"))
            if b == 0 and stage > 1:
                L.append((pre + ".downsample", _cd(n, h, cin, ch, 1, stride=stride, relu=0), prev))
            prev = pre + ".conv2"
            h, cin = h2, ch
    return L


def vgg16(n=128):
    cfg = [(64, 2, 224), (128, 2, 112), (256, 3, 56), (512, 3, 28), (512, 3, 14)]
    L, cin, prev = [], 3, None
    for bi, (ch, reps, h) in enumerate(cfg, start=1):
        for r in range(reps):
            name = f"conv{bi}_{r + 1}"
            L.append((name, _cd(n, h, cin, ch, 3), prev))
            prev, cin = name, ch
        prev = None  # max-pool between blocks
    return L


def mobilenet_v2(n=1024):
    L = [("stem", _cd(n, 224, 3, 32, 3, stride=2), None)]
    h, cin, prev = 112, 32, "stem"
    settings = [(1, 16, 1, 1), (6, 24, 2, 2), (6, 32, 3, 2), (6, 64, 4, 2), (6, 96, 3, 1), (6, 160, 3, 2), (6, 320, 1, 1)]
    bi = 0
    for t, c, reps, s in settings:
        for i in range(reps):
            stride = s if i == 0 else 1
            hid = cin * t
            pre = f"b{bi}"
            if t != 1:
                L.append((pre + ".expand", _cd(n, h, cin, hid, 1), prev))
                prev = pre + ".expand"
            L.append((pre + ".dw", _cd(n, h, hid, hid, 3, stride=stride, groups=hid), prev))
            h = h // stride
            L.append((pre + ".project", _cd(n, h, hid, c, 1, relu=0), pre + ".dw"))
            prev, cin = pre + ".project", c
            bi += 1
    L.append(("last", _cd(n, h, cin, 1280, 1), prev))
    return L


def resnet50_full(n=32):
    """ResNet-50 as ONE int8 graph: conv1 -> 3x3/2 max-pool -> 16 bottlenecks whose residual joins are saturating int8
    adds + ReLU (conv3 and the downsample convolutions requantise without ReLU, as in the torchvision topology).
    Entries: (name, ConvDesc | PoolDesc | AddDesc, input name | (a, b) for an add)."""
    L = [("conv1", _cd(n, 224, 3, 64, 7, stride=2, pad=3), None),
         ("maxpool", PoolDesc(n, 112, 112, 64, 3, 3, 2, 2, 1, 1), "conv1")]
    h, cin, prev = 56, 64, "maxpool"
    for stage, (mid, blocks) in enumerate(((64, 3), (128, 4), (256, 6), (512, 3)), start=1):
        out = mid * 4
        for b in range(blocks):
            stride = 2 if (b == 0 and stage > 1) else 1
            pre = f"l{stage}.{b}"
            L.append((pre + ".conv1", _cd(n, h, cin, mid, 1), prev))
            L.append((pre + ".conv2", _cd(n, h, mid, mid, 3, stride=stride), pre + ".conv1"))
            h2 = h // stride
            L.append((pre + ".conv3", _cd(n, h2, mid, out, 1, relu=0), pre + ".conv2"))
            identity = prev
            if b == 0:
                L.append((pre + ".downsample", _cd(n, h, cin, out, 1, stride=stride, relu=0), prev))
                identity = pre + ".downsample"
            L.append((pre + ".add", AddDesc(n, h2, h2, out, relu=1), (pre + ".conv3", identity)))
            prev, h, cin = pre + ".add", h2, out
    return L


def resnet18_full(n=32):
    L = [("conv1", _cd(n, 224, 3, 64, 7, stride=2, pad=3), None),
         ("maxpool", PoolDesc(n, 112, 112, 64, 3, 3, 2, 2, 1, 1), "conv1")]
    h, cin, prev = 56, 64, "maxpool"
    for stage, ch in enumerate((64, 128, 256, 512), start=1):
        for b in range(2):
            stride = 2 if (b == 0 and stage > 1) else 1
            pre = f"l{stage}.{b}"
            L.append((pre + ".conv1", _cd(n, h, cin, ch, 3, stride=stride), prev))
            h2 = h // stride
            L.append((pre + ".conv2", _cd(n, h2, ch, ch, 3, relu=0), pre + ".conv1"))
            identity = prev
            if b == 0 and stage > 1:
                L.append((pre + ".downsample", _cd(n, h, cin, ch, 1, stride=stride, relu=0), prev))
                identity = pre + ".downsample"
            L.append((pre + ".add", AddDesc(n, h2, h2, ch, relu=1), (pre + ".conv2", identity)))
            prev, h, cin = pre + ".add", h2, ch
    return L


def vgg16_full(n=16):
    """VGG-16's convolutional part with its 2x2/2 max-pools between the blocks."""
    cfg = [(64, 2, 224), (128, 2, 112), (256, 3, 56), (512, 3, 28), (512, 3, 14)]
    L, cin, prev = [], 3, None
    for bi, (ch, reps, h) in enumerate(cfg, start=1):
        for r in range(reps):
            name = f"conv{bi}_{r + 1}"
            L.append((name, _cd(n, h, cin, ch, 3), prev))
            prev, cin = name, ch
        L.append((f"pool{bi}", PoolDesc(n, h, h, ch, 2, 2, 2, 2, 0, 0), prev))
        prev = f"pool{bi}"
    return L


NETWORKS = {
    "single_3x3": single_3x3,
    "resnet18": resnet18,
    "resnet50": resnet50,
    "vgg16": vgg16,
    "mobilenet_v2": mobilenet_v2,
    "resnet50_full": resnet50_full,
    "resnet18_full": resnet18_full,
    "vgg16_full": vgg16_full,
}
DEFAULT_BATCH = {"single_3x3": 1, "resnet18": 256, "resnet50": 512, "vgg16": 128, "mobilenet_v2": 1024,
                 "resnet50_full": 512, "resnet18_full": 256, "vgg16_full": 128}


def total_work(layers):
    """(GMAC, algorithmic GB) over a layer list, host arithmetic only."""
    macs = byts = 0
    for _, d, _ in layers:
        if not isinstance(d, ConvDesc):
            continue
        p = (d.h + 2 * d.pad_h - (d.dil_h * (d.r - 1) + 1)) // d.stride_h + 1
        q = (d.w + 2 * d.pad_w - (d.dil_w * (d.s - 1) + 1)) // d.stride_w + 1
        cg = d.c // d.groups
        macs += d.n * p * q * d.k * cg * d.r * d.s
        byts += d.n * d.h * d.w * d.c + d.k * cg * d.r * d.s + d.n * p * q * d.k + 8 * d.k
    return macs / 1e9, byts / 1e9
