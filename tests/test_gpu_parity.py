"""GPU parity tests (run with -m gpu on a B200): every kernel of liblowbit-cnn, called through the C ABI,
must equal the CPU oracle bit for bit — int32 accumulators and requantised int8 alike."""
import numpy as np
import pytest

from oracle import oracle
from oracle.oracle import ConvDesc as D

pytestmark = pytest.mark.gpu

DIRECT, IGEMM, DW = 1, 2, 3


def _check(d, **kw):
    from tests.parity_util import check_case
    nbad, total, name, detail = check_case(d, **kw)
    assert nbad == 0, f"{name}: {detail}"
    return name


# ---- CUDA-core direct kernel: any shape -----------------------------------------------------------------
DIRECT_CASES = [
    D(n=2, h=9, w=7, c=8, k=12, r=3, s=3, pad_h=1, pad_w=1, relu=1),
    D(n=1, h=12, w=12, c=16, k=8, r=3, s=3, stride_h=2, stride_w=2, pad_h=1, pad_w=1),
    D(n=2, h=20, w=20, c=3, k=16, r=7, s=7, stride_h=2, stride_w=2, pad_h=3, pad_w=3, relu=1),   # stem-like
    D(n=1, h=6, w=5, c=32, k=24, r=1, s=1, relu=1),
    D(n=1, h=10, w=10, c=8, k=8, r=3, s=3, pad_h=2, pad_w=2, dil_h=2, dil_w=2, groups=2, relu=1),
    D(n=1, h=5, w=9, c=4, k=6, r=3, s=2, pad_h=0, pad_w=1, stride_h=1, stride_w=2),
    D(n=1, h=7, w=7, c=5, k=3, r=3, s=3, pad_h=1, pad_w=1),                                      # odd C, K
    D(n=2, h=9, w=9, c=6, k=6, r=3, s=3, pad_h=1, pad_w=1, groups=6, relu=1),                    # depthwise, C%4 != 0
]


@pytest.mark.parametrize("d", DIRECT_CASES, ids=lambda d: f"c{d.c}k{d.k}r{d.r}s{d.stride_h}g{d.groups}")
@pytest.mark.parametrize("out_mode", [0, 1])
def test_direct_kernel(d, out_mode):
    _check(D(**{**d.__dict__, "out_mode": out_mode}), force=DIRECT)


# ---- depthwise kernel ---------------------------------------------------------------------------------
DW_CASES = [
    D(n=2, h=9, w=9, c=24, k=24, r=3, s=3, pad_h=1, pad_w=1, groups=24, relu=1),
    D(n=1, h=14, w=14, c=96, k=96, r=3, s=3, stride_h=2, stride_w=2, pad_h=1, pad_w=1, groups=96, relu=1),
    D(n=1, h=7, w=7, c=960, k=960, r=3, s=3, pad_h=1, pad_w=1, groups=960),
    # 3x3 fast path: odd output height (half-used row pair), strips with a clipped tail, W != H, no padding
    D(n=2, h=13, w=37, c=32, k=32, r=3, s=3, pad_h=1, pad_w=1, groups=32, relu=1),
    D(n=1, h=11, w=19, c=16, k=16, r=3, s=3, groups=16, relu=1),
    D(n=2, h=23, w=41, c=8, k=8, r=3, s=3, stride_h=2, stride_w=2, pad_h=1, pad_w=1, groups=8, relu=1),
    D(n=1, h=112, w=112, c=32, k=32, r=3, s=3, pad_h=1, pad_w=1, groups=32, relu=1),
    D(n=1, h=56, w=56, c=144, k=144, r=3, s=3, stride_h=2, stride_w=2, pad_h=1, pad_w=1, groups=144, relu=1),
    # generic depthwise kernel: 5x5, dilation, stride 3
    D(n=1, h=15, w=15, c=12, k=12, r=5, s=5, pad_h=2, pad_w=2, groups=12, relu=1),
    D(n=1, h=15, w=17, c=8, k=8, r=3, s=3, pad_h=2, pad_w=2, dil_h=2, dil_w=2, groups=8),
    D(n=1, h=16, w=16, c=8, k=8, r=3, s=3, stride_h=3, stride_w=3, pad_h=1, pad_w=1, groups=8),
]


@pytest.mark.parametrize("d", DW_CASES, ids=lambda d: f"h{d.h}w{d.w}c{d.c}r{d.r}s{d.stride_h}d{d.dil_h}")
@pytest.mark.parametrize("out_mode", [0, 1])
def test_depthwise_kernel(d, out_mode):
    assert _check(D(**{**d.__dict__, "out_mode": out_mode})) == "depthwise"


# ---- tcgen05 implicit GEMM ----------------------------------------------------------------------------
IGEMM_CASES = [
    # pure GEMM (tiled TMA): 128B / 64B / 32B swizzle, N tiles, M tail
    D(n=1, h=16, w=16, c=128, k=128, r=1, s=1),
    D(n=2, h=14, w=14, c=64, k=256, r=1, s=1, relu=1),
    D(n=1, h=9, w=11, c=32, k=64, r=1, s=1),
    D(n=1, h=7, w=7, c=512, k=2048, r=1, s=1, relu=1),          # tiles_n = 8, many k-blocks
    D(n=3, h=5, w=5, c=16, k=16, r=1, s=1),                     # C padded 16 -> 32, smallest N
    D(n=1, h=12, w=12, c=144, k=48, r=1, s=1, relu=1),          # C=144 -> 160 zero-padded, bn = 48
    D(n=1, h=8, w=8, c=256, k=320, r=1, s=1),                   # K_out tail tile (320 = 256 + 64)
    # im2col TMA: 3x3 pad 1, stride 2, 1x1 stride 2, 7x7, dilation, rectangular
    D(n=1, h=56, w=56, c=64, k=64, r=3, s=3, pad_h=1, pad_w=1, relu=1),       # BASELINE config 1
    D(n=2, h=14, w=14, c=256, k=256, r=3, s=3, pad_h=1, pad_w=1, relu=1),
    D(n=2, h=28, w=28, c=128, k=128, r=3, s=3, stride_h=2, stride_w=2, pad_h=1, pad_w=1, relu=1),
    D(n=2, h=28, w=28, c=256, k=512, r=1, s=1, stride_h=2, stride_w=2),
    D(n=1, h=17, w=13, c=32, k=32, r=3, s=3, pad_h=1, pad_w=1),
    D(n=1, h=20, w=20, c=16, k=32, r=7, s=7, stride_h=2, stride_w=2, pad_h=3, pad_w=3, relu=1),
    D(n=1, h=15, w=15, c=64, k=64, r=3, s=3, pad_h=2, pad_w=2, dil_h=2, dil_w=2),
    D(n=1, h=12, w=10, c=64, k=32, r=3, s=1, pad_h=1, pad_w=0),
    D(n=1, h=10, w=10, c=64, k=64, r=3, s=3),                                 # VALID (reference style)
    # shifted-window path: tiles of whole rows, column tiles, ragged edges, N tiles, 16-byte pixels
    D(n=3, h=28, w=28, c=128, k=128, r=3, s=3, pad_h=1, pad_w=1, relu=1),
    D(n=5, h=7, w=7, c=512, k=512, r=3, s=3, pad_h=1, pad_w=1),
    D(n=1, h=20, w=224, c=64, k=64, r=3, s=3, pad_h=1, pad_w=1, relu=1),
    D(n=2, h=9, w=150, c=32, k=48, r=3, s=3, pad_h=1, pad_w=1),
    D(n=2, h=19, w=23, c=32, k=32, r=5, s=5, pad_h=2, pad_w=2, relu=1),
    D(n=2, h=14, w=14, c=128, k=512, r=3, s=3, pad_h=1, pad_w=1, relu=1),
    D(n=2, h=30, w=40, c=16, k=32, r=3, s=3, pad_h=1, pad_w=1, relu=1),        # C=16: paired taps, phantom 4th tap
    D(n=2, h=33, w=115, c=16, k=64, r=4, s=4, relu=1),                         # the space-to-depth'd 7x7 stem shape
    D(n=2, h=20, w=57, c=16, k=32, r=2, s=2),
    D(n=1, h=8, w=40, c=64, k=64, r=1, s=3, pad_h=0, pad_w=1),
    D(n=1, h=6, w=112, c=128, k=128, r=3, s=3, pad_h=1, pad_w=1, relu=1),     # VGG conv2_2 rows: windows too wide for pair mode
]


@pytest.mark.parametrize("d", IGEMM_CASES, ids=lambda d: f"n{d.n}h{d.h}c{d.c}k{d.k}r{d.r}s{d.stride_h}p{d.pad_h}")
@pytest.mark.parametrize("out_mode", [1, 0])
def test_igemm_tc_kernel(d, out_mode):
    assert _check(D(**{**d.__dict__, "out_mode": out_mode}), force=IGEMM) == "igemm_tc"


# pointwise layers whose C or K is not a multiple of 16: pixel-group rewrite (f pixels per GEMM row, block-diagonal filter)
PIXEL_GROUP_CASES = [
    D(n=2, h=14, w=14, c=24, k=144, r=1, s=1, relu=1),      # MobileNetV2 expand 24 -> 144 (f = 2, N tiles 2 x 144)
    D(n=2, h=14, w=14, c=144, k=24, r=1, s=1),              # project 144 -> 24
    D(n=1, h=56, w=56, c=96, k=24, r=1, s=1),               # project 96 -> 24
    D(n=1, h=10, w=10, c=24, k=24, r=1, s=1, relu=1),
    D(n=1, h=6, w=10, c=12, k=20, r=1, s=1, relu=1),        # f = 4
    D(n=3, h=4, w=4, c=40, k=8, r=1, s=1),                  # f = 2, K' = 16
]


@pytest.mark.parametrize("d", PIXEL_GROUP_CASES, ids=lambda d: f"n{d.n}h{d.h}c{d.c}k{d.k}")
@pytest.mark.parametrize("out_mode", [1, 0])
def test_igemm_pixel_group_rewrite(d, out_mode):
    assert _check(D(**{**d.__dict__, "out_mode": out_mode})) == "igemm_tc"


def test_pixel_group_rewrite_needs_divisible_pixels():
    """An odd pixel count cannot be grouped in pairs: the planner falls back to the CUDA-core kernel."""
    assert _check(D(n=1, h=5, w=5, c=24, k=24, r=1, s=1, relu=1)) == "direct"


@pytest.mark.parametrize("d", [c for c in IGEMM_CASES if c.stride_h == 1 and c.r > 1][:6],
                         ids=lambda d: f"n{d.n}h{d.h}c{d.c}k{d.k}r{d.r}")
def test_igemm_im2col_path_on_stride1_shapes(d):
    """The same stride-1 layers with the window planner disabled: the TMA im2col path stays covered."""
    assert _check(D(**{**d.__dict__, "out_mode": 0}), force=IGEMM, options={"window": 0}) == "igemm_tc"


@pytest.mark.parametrize("d", [c for c in IGEMM_CASES if c.c >= 32], ids=lambda d: f"n{d.n}h{d.h}c{d.c}k{d.k}r{d.r}s{d.stride_h}")
def test_igemm_cta_pairs_forced(d):
    """Every igemm shape that can run in CTA pairs (two-CTA clusters, cta_group::2 MMAs, half of the filter rows per CTA)
    does so here regardless of the planner's preference, with the resident-filter variant switched off so that small
    filter matrices stream too."""
    assert _check(D(**{**d.__dict__, "out_mode": 0}), force=IGEMM, options={"cta_pairs": 1, "resident_filter": 0}) == "igemm_tc"


@pytest.mark.parametrize("d", [c for c in IGEMM_CASES if c.r == 1 or c.stride_h > 1],
                         ids=lambda d: f"n{d.n}h{d.h}c{d.c}k{d.k}r{d.r}s{d.stride_h}")
def test_igemm_per_warp_stores_forced(d):
    """Ring-mode layers with the per-warp epilogue (own 32-row staging buffer and TMA store per warp, no team barrier)
    forced on for every N tile it supports (256 / 128 / 64 / 32 columns), and forced off."""
    assert _check(D(**{**d.__dict__, "out_mode": 0}), force=IGEMM, options={"warp_store": 1}) == "igemm_tc"
    assert _check(D(**{**d.__dict__, "out_mode": 0}), force=IGEMM, options={"warp_store": 0}) == "igemm_tc"


FOLD_EXTRA_CASES = [
    D(n=2, h=28, w=28, c=64, k=256, r=1, s=1, relu=1),                         # the 64->256 expansion (per-warp stores)
    D(n=2, h=28, w=28, c=64, k=64, r=3, s=3, pad_h=1, pad_w=1, relu=1),        # window, two MMA warps, two tiles per iteration
    D(n=2, h=32, w=32, c=3, k=64, r=7, s=7, stride_h=2, stride_w=2, pad_h=3, pad_w=3, relu=1),   # stem (16-byte pixels)
    D(n=2, h=28, w=28, c=24, k=144, r=1, s=1, relu=1),                         # pixel groups: bias index = channel % 144
    D(n=1, h=14, w=14, c=128, k=80, r=1, s=1),                                 # 80-column tile: unswizzled staging panel
    D(n=2, h=28, w=28, c=128, k=512, r=1, s=1, relu=1),                        # resident filter matrix, 2 N tiles (128->512)
    D(n=2, h=14, w=14, c=64, k=384, r=1, s=1, relu=1),                         # resident, 2 N tiles of 192
    D(n=3, h=14, w=14, c=32, k=600, r=1, s=1),                                 # resident, 3 N tiles, ragged last tile (600 of 624)
    D(n=1, h=20, w=20, c=16, k=320, r=3, s=3, pad_h=1, pad_w=1, relu=1),       # 16-byte pixels, resident, 2 N tiles
]


@pytest.mark.parametrize("d", IGEMM_CASES + FOLD_EXTRA_CASES, ids=lambda d: f"n{d.n}h{d.h}c{d.c}k{d.k}r{d.r}s{d.stride_h}")
def test_igemm_bias_folded_into_mma(d):
    """Resident-filter layers with the bias fed through the first MMA of every tile (constant A block x bias digits)
    instead of the epilogue's add, forced on for every N tile width, int8 and raw int32 outputs.  _check's biases are a few
    thousand; the second call uses biases beyond the digit range, which the kernel must detect and add the classic way."""
    on, off = {"fold_bias": 1}, {"fold_bias": 0}
    assert _check(D(**{**d.__dict__, "out_mode": 0}), options=on) in ("igemm_tc", "stem_tc")
    assert _check(D(**{**d.__dict__, "out_mode": 1}), options=on) in ("igemm_tc", "stem_tc")
    assert _check(D(**{**d.__dict__, "out_mode": 0}), bias_range=3_000_000, options=on) in ("igemm_tc", "stem_tc")
    assert _check(D(**{**d.__dict__, "out_mode": 1}), bias_range=3_000_000, options=on) in ("igemm_tc", "stem_tc")
    assert _check(D(**{**d.__dict__, "out_mode": 0}), options=off) in ("igemm_tc", "stem_tc")


MULTI_TILE_CASES = [
    D(n=4, h=28, w=28, c=64, k=64, r=3, s=3, pad_h=1, pad_w=1, relu=1),        # window, resident B, two MMA warps
    D(n=2, h=28, w=28, c=64, k=256, r=1, s=1, relu=1),                         # tiled, resident B, 256-wide tile
    D(n=2, h=28, w=28, c=256, k=64, r=1, s=1, relu=1),                         # tiled, resident B, two MMA warps
    D(n=2, h=14, w=14, c=256, k=256, r=3, s=3, pad_h=1, pad_w=1, relu=1),      # window, streaming B, 2 channel chunks
    D(n=2, h=14, w=14, c=512, k=1024, r=1, s=1),                               # tiled, streaming B, 4 N tiles
    D(n=2, h=28, w=28, c=128, k=128, r=3, s=3, stride_h=2, stride_w=2, pad_h=1, pad_w=1, relu=1),   # im2col
    D(n=2, h=40, w=40, c=3, k=64, r=7, s=7, stride_h=2, stride_w=2, pad_h=3, pad_w=3, relu=1),      # stem
    D(n=2, h=28, w=28, c=24, k=144, r=1, s=1, relu=1),                         # pixel groups, 2 N tiles
    D(n=3, h=28, w=28, c=128, k=128, r=3, s=3, pad_h=1, pad_w=1, relu=1),      # paired tiles, odd tile count (padding image)
    D(n=3, h=7, w=7, c=64, k=512, r=3, s=3, pad_h=1, pad_w=1),                 # paired tiles, 4 N tiles, 1 tile per image
    D(n=5, h=14, w=14, c=512, k=256, r=3, s=3, pad_h=1, pad_w=1, relu=1),      # paired tiles, 4 channel chunks
    D(n=3, h=14, w=14, c=256, k=320, r=1, s=1, relu=1),                        # CTA pairs: odd M tile count, 160-wide N tiles
    D(n=3, h=15, w=15, c=128, k=96, r=3, s=3, stride_h=2, stride_w=2, pad_h=1, pad_w=1),   # CTA pairs: im2col, M tail, 48 B rows per CTA
    D(n=1, h=9, w=9, c=2048, k=512, r=1, s=1, relu=1),                         # CTA pairs: one real M tile + its padding twin
    D(n=2, h=28, w=28, c=128, k=512, r=1, s=1, relu=1),                        # resident filter matrix with 2 N tiles
    D(n=3, h=14, w=14, c=32, k=600, r=1, s=1),                                 # resident, 3 N tiles, ragged last tile
]


@pytest.mark.parametrize("grid", [1, 3])
@pytest.mark.parametrize("d", MULTI_TILE_CASES, ids=lambda d: f"n{d.n}h{d.h}c{d.c}k{d.k}r{d.r}s{d.stride_h}")
def test_many_tiles_per_cta(d, grid):
    """Cap the persistent grid so every CTA walks many tiles: ring/window phases, TMEM accumulator reuse, the
    alternating MMA warps and the staging ring all wrap several times (a full-size layer does this on 148 SMs)."""
    g = {"max_grid": grid}
    # the opt-in paired-tile mode stays covered (cases marked "paired tiles")
    assert _check(D(**{**d.__dict__, "out_mode": 0}), options={**g, "paired_tiles": 1}) in ("igemm_tc", "stem_tc")
    assert _check(D(**{**d.__dict__, "out_mode": 0}), options=g) in ("igemm_tc", "stem_tc")
    # CTA pairs (cta_group::2) forced on wherever the layer streams its filter matrix, then forced off: the planner's own
    # choice between the two only depends on the K-loop length
    assert _check(D(**{**d.__dict__, "out_mode": 0}), options={**g, "cta_pairs": 1}) in ("igemm_tc", "stem_tc")
    if d.r == 3 and d.c >= 128:
        assert _check(D(**{**d.__dict__, "out_mode": 1}), options={**g, "cta_pairs": 1}) == "igemm_tc"
    # ... and the per-warp epilogue with and without CTA pairs
    assert _check(D(**{**d.__dict__, "out_mode": 0}), options={**g, "cta_pairs": 0, "warp_store": 1}) in ("igemm_tc", "stem_tc")
    assert _check(D(**{**d.__dict__, "out_mode": 0}), options={**g, "cta_pairs": 1, "warp_store": 1}) in ("igemm_tc", "stem_tc")
    # reversed traversal (what a network's alternate layers do)
    assert _check(D(**{**d.__dict__, "out_mode": 0}), options={**g, "reverse": 1}) in ("igemm_tc", "stem_tc")
    if grid == 1 and d.r == 3 and d.c >= 64:
        assert _check(D(**{**d.__dict__, "out_mode": 1}), options=g) == "igemm_tc"    # raw accumulators through the same walk


# ---- column-split epilogue and N-stationary filter tiles (r02) ------------------------------------------------
SPLIT_CASES = [c for c in IGEMM_CASES + FOLD_EXTRA_CASES + MULTI_TILE_CASES if c.k > 128]     # N tiles > 128 columns


@pytest.mark.parametrize("d", SPLIT_CASES, ids=lambda d: f"n{d.n}h{d.h}c{d.c}k{d.k}r{d.r}s{d.stride_h}")
def test_igemm_column_split_epilogue(d):
    """N tiles with two TMEM accumulator stages: both epilogue teams drain every tile (alternate panels in int8 mode,
    column halves in int32 mode; four 64-byte panels with per-warp stores), forced on and off, with every way the
    operands can arrive (resident / streamed / CTA pairs) and with few CTAs so that the accumulator stages wrap."""
    for opts in ({"epi_split": 1}, {"epi_split": 0}, {"epi_split": 1, "warp_store": 1}, {"epi_split": 1, "warp_store": 0},
                 {"epi_split": 1, "cta_pairs": 1, "resident_filter": 0}, {"epi_split": 1, "max_grid": 2},
                 {"epi_split": 1, "max_grid": 3, "warp_store": 1, "fold_bias": 1}, {"epi_split": 1, "max_grid": 2, "reverse": 1}):
        assert _check(D(**{**d.__dict__, "out_mode": 0}), options=opts) == "igemm_tc", opts
    for opts in ({"epi_split": 1}, {"epi_split": 1, "max_grid": 2}, {"epi_split": 1, "cta_pairs": 1, "resident_filter": 0}):
        assert _check(D(**{**d.__dict__, "out_mode": 1}), options=opts) == "igemm_tc", opts


N_STATIONARY_CASES = [
    D(n=2, h=14, w=14, c=256, k=1024, r=1, s=1, relu=1),                       # ResNet-50 l3.x.conv3: 4 N tiles of 64 KB
    D(n=3, h=7, w=7, c=512, k=2048, r=1, s=1, relu=1),                         # l4.x.conv3: 8 N tiles of 128 KB
    D(n=2, h=28, w=28, c=256, k=512, r=1, s=1, stride_h=2, stride_w=2),        # l2.0.downsample: im2col A, 2 N tiles
    D(n=2, h=14, w=14, c=512, k=1024, r=1, s=1, stride_h=2, stride_w=2),       # l3.0.downsample: 4 N tiles of 128 KB
    D(n=2, h=14, w=14, c=256, k=608, r=1, s=1, relu=1),                        # ragged last N tile (3 x 208 columns for 608)
    D(n=2, h=14, w=14, c=32, k=640, r=3, s=3, pad_h=1, pad_w=1, relu=1),       # window A, 3 N tiles of 224 columns
    D(n=2, h=9, w=9, c=320, k=1280, r=1, s=1, relu=1),                         # MobileNetV2 `last`: 64-byte K chunks
]


@pytest.mark.parametrize("d", N_STATIONARY_CASES, ids=lambda d: f"n{d.n}h{d.h}c{d.c}k{d.k}r{d.r}s{d.stride_h}")
def test_igemm_n_stationary_filter_tiles(d):
    """Filter matrices too large to be resident as a whole: every CTA keeps the N tile it works on (grid a multiple of
    tiles_n), with and without the folded bias, with the grid capped to one and to several CTAs per N tile, and with
    grids that cannot be N-stationary at all (fewer CTAs than N tiles -> the planner must stream)."""
    import lowbitdnn_project_b200 as lbc
    from tests.parity_util import lbc_desc
    ns = {"n_stationary": 1}     # (forced: on its own the planner keeps tiles of up to 64 KB, the option allows 128 KB)
    plan = lbc.ConvPlan(lbc_desc(D(**{**d.__dict__, "out_mode": 0})), options=ns)
    assert "n-stationary" in plan.describe(), plan.describe()
    plan.close()
    tiles_n = -(-d.k // 256)
    for opts in ({}, {"n_stationary": 0}, {"fold_bias": 0}, {"fold_bias": 1}, {"max_grid": tiles_n}, {"max_grid": 2 * tiles_n + 1},
                 {"max_grid": tiles_n - 1}, {"max_grid": 3 * tiles_n, "reverse": 1}, {"epi_split": 1, "max_grid": 2 * tiles_n},
                 {"warp_store": 1, "max_grid": 2 * tiles_n}, {"warp_store": 0}):
        assert _check(D(**{**d.__dict__, "out_mode": 0}), options={**ns, **opts}) == "igemm_tc", opts
    assert _check(D(**{**d.__dict__, "out_mode": 0})) == "igemm_tc"            # the planner's own choice
    for opts in ({}, {"max_grid": 2 * tiles_n}):
        assert _check(D(**{**d.__dict__, "out_mode": 1}), options={**ns, **opts}) == "igemm_tc", opts
    assert _check(D(**{**d.__dict__, "out_mode": 0}), bias_range=3_000_000, options={**ns, "fold_bias": 1}) == "igemm_tc"


# ---- narrow N tiles with per-warp stores (r02): one tile of 32 / 64 columns, every A mode ---------------------
NARROW_CASES = [
    # window tiles: pitch padded to 32 / 64 / 128, runs of 32 pixels + a ragged last run, halo-only warps
    D(n=3, h=56, w=56, c=64, k=64, r=3, s=3, pad_h=1, pad_w=1, relu=1),        # 2 x 56 per tile, pitch 64, ragged run of 24
    D(n=5, h=28, w=28, c=64, k=64, r=3, s=3, pad_h=1, pad_w=1, relu=1),        # 4 x 28, pitch 32: every run ragged (28)
    D(n=2, h=9, w=112, c=64, k=64, r=3, s=3, pad_h=1, pad_w=1),                # 1 x 112, pitch 128, ragged run of 16
    D(n=1, h=7, w=224, c=32, k=64, r=3, s=3, pad_h=1, pad_w=1, relu=1),        # two column tiles of 112
    D(n=2, h=11, w=150, c=32, k=32, r=3, s=3, pad_h=1, pad_w=1, relu=1),       # column tiles of 75: run of 11, a halo-only warp
    D(n=2, h=13, w=96, c=64, k=64, r=3, s=3, pad_h=1, pad_w=1),                # 1 x 96: three full runs, no ragged one
    D(n=3, h=7, w=48, c=32, k=32, r=3, s=3, pad_h=1, pad_w=1, relu=1),         # 2 x 48, pitch 64, odd P: ragged row tile
    D(n=2, h=15, w=100, c=32, k=64, r=5, s=5, pad_h=2, pad_w=2, relu=1),       # pitch 128, run of 4
    D(n=2, h=30, w=60, c=16, k=64, r=3, s=3, pad_h=1, pad_w=1, relu=1),        # 16-byte pixels (paired taps)
    D(n=2, h=45, w=224, c=3, k=64, r=7, s=7, stride_h=2, stride_w=2, pad_h=3, pad_w=3, relu=1),   # ResNet stem rows (1 x 112)
    D(n=2, h=40, w=224, c=3, k=32, r=3, s=3, stride_h=2, stride_w=2, pad_h=1, pad_w=1, relu=1),   # MobileNetV2 stem rows
    D(n=1, h=12, w=224, c=3, k=64, r=3, s=3, pad_h=1, pad_w=1, relu=1),        # VGG conv1_1 rows: two column tiles
    # ring modes
    D(n=3, h=28, w=28, c=64, k=64, r=1, s=1, relu=1),                          # tiled, resident filter
    D(n=2, h=19, w=19, c=256, k=64, r=1, s=1),                                 # tiled, M tail (361 rows per image)
    D(n=2, h=14, w=14, c=384, k=64, r=1, s=1, relu=1),
    D(n=2, h=28, w=28, c=192, k=32, r=1, s=1),
    D(n=3, h=29, w=29, c=64, k=64, r=3, s=3, stride_h=2, stride_w=2, pad_h=1, pad_w=1, relu=1),   # im2col
    D(n=2, h=14, w=14, c=1024, k=64, r=1, s=1),                                # streaming filter matrix (128 KB)
]


@pytest.mark.parametrize("d", NARROW_CASES, ids=lambda d: f"n{d.n}h{d.h}w{d.w}c{d.c}k{d.k}r{d.r}s{d.stride_h}")
def test_narrow_tiles_with_per_warp_stores(d):
    """The narrow warp-store path on its own, then with capped grids (every CTA walks many tiles: TMEM stages and both
    staging buffers wrap), one staging buffer per warp, reversed traversal; last the planner's default for the layer."""
    import lowbitdnn_project_b200 as lbc
    ws = {"warp_store": 1}                 # the path is opt-in (measured neutral to slightly slower, see the planner)
    plan = lbc.ConvPlan(lbc.ConvDesc(**d.__dict__), options=ws)
    try:
        assert "narrow-warp-stores" in plan.describe(), plan.describe()
    finally:
        plan.close()
    for options in ({}, {"max_grid": 1}, {"max_grid": 3, "reverse": 1}, {"max_grid": 2, "stage_bufs": 1}, {"max_grid": 5, "two_mma_warps": 0}):
        assert _check(d, options={**ws, **options}) in ("igemm_tc", "stem_tc")
    assert _check(d) in ("igemm_tc", "stem_tc")           # the planner's own choice (team path)


# ---- CTA pairs with the filter matrix resident as two halves (r02): window A, one N tile -----------------------
PAIR_RESIDENT_CASES = [
    D(n=3, h=28, w=28, c=128, k=128, r=3, s=3, pad_h=1, pad_w=1, relu=1),      # ResNet-50 l2.x.conv2; odd tile count
    D(n=5, h=14, w=14, c=128, k=128, r=3, s=3, pad_h=1, pad_w=1),
    D(n=2, h=6, w=112, c=128, k=128, r=3, s=3, pad_h=1, pad_w=1, relu=1),      # VGG conv2_2 rows: 44 KB windows
    D(n=2, h=19, w=23, c=64, k=64, r=5, s=5, pad_h=2, pad_w=2, relu=1),        # 100 KB matrix, 64-column tile (32 rows per CTA)
    D(n=3, h=20, w=20, c=128, k=96, r=3, s=3, pad_h=1, pad_w=1),               # 96-column tile: 48 rows per CTA
    D(n=2, h=20, w=20, c=128, k=128, r=3, s=3, pad_h=2, pad_w=2, dil_h=2, dil_w=2, relu=1),   # dilated taps
    # ring modes: only A streams
    D(n=3, h=14, w=14, c=512, k=256, r=1, s=1, relu=1),                        # ResNet-50 l3.0.conv1: 64 KB per CTA
    D(n=3, h=9, w=9, c=1024, k=256, r=1, s=1),                                 # l3.x.conv1: 128 KB per CTA, M tail
    D(n=3, h=27, w=27, c=128, k=128, r=3, s=3, stride_h=2, stride_w=2, pad_h=1, pad_w=1, relu=1),   # im2col (l2.0.conv2)
    D(n=2, h=12, w=12, c=640, k=192, r=1, s=1, relu=1),                        # 192-column tile: 96 rows per CTA
]


@pytest.mark.parametrize("d", PAIR_RESIDENT_CASES, ids=lambda d: f"n{d.n}h{d.h}w{d.w}c{d.c}k{d.k}r{d.r}")
def test_cta_pairs_with_resident_filter_halves(d):
    import lowbitdnn_project_b200 as lbc
    # window layers get there by the planner's own choice; the ring modes (measured slower, see the planner) with
    # resident_filter = 3, and at these small sizes only with the 256-wide N tile kept (max_bn)
    base = {} if d.stride_h == 1 and d.r > 1 else {"resident_filter": 3, "max_bn": 256}
    plan = lbc.ConvPlan(lbc.ConvDesc(**d.__dict__), options=base or None)
    try:
        assert "b=resident,cta-pair" in plan.describe(), plan.describe()
    finally:
        plan.close()
    assert _check(d, options=base or None) == "igemm_tc"
    assert _check(D(**{**d.__dict__, "out_mode": 1}), options=base or None) == "igemm_tc"
    for options in ({"max_grid": 2}, {"max_grid": 4, "reverse": 1}, {"max_grid": 6, "stage_bufs": 1}):
        assert _check(d, options={**base, **options}) == "igemm_tc"
    assert _check(d, options={"resident_filter": 2}) == "igemm_tc"


# ---- split last round of a CTA-pair launch (r02): leftover pair-steps run as half-width tiles ------------------
TAIL_SPLIT_CASES = [
    # (layer, options): grid capped so that the pair-steps do not divide by the pairs; max_bn keeps the 256-wide tile
    (D(n=5, h=16, w=16, c=256, k=256, r=3, s=3, pad_h=1, pad_w=1, relu=1), {"max_grid": 8}),          # 5 steps, 4 pairs
    (D(n=5, h=16, w=16, c=256, k=256, r=3, s=3, pad_h=1, pad_w=1, relu=1), {"max_grid": 8, "reverse": 1}),
    (D(n=7, h=14, w=14, c=1024, k=256, r=1, s=1), {"max_grid": 8}),                                   # tiled; 11 M tiles + padding twin
    (D(n=3, h=33, w=33, c=256, k=256, r=3, s=3, stride_h=2, stride_w=2, pad_h=1, pad_w=1, relu=1), {"max_grid": 6}),
    (D(n=15, h=16, w=16, c=512, k=256, r=1, s=1, relu=1), {"max_grid": 12}),                          # 15 steps, 6 pairs: 3 leftover
]


@pytest.mark.parametrize("case", TAIL_SPLIT_CASES, ids=lambda c: f"n{c[0].n}h{c[0].h}c{c[0].c}r{c[0].r}s{c[0].stride_h}g{c[1]['max_grid']}")
def test_cta_pair_tail_split(case):
    import lowbitdnn_project_b200 as lbc
    d, opt = case
    opt = {**opt, "max_bn": 256, "cta_pairs": 1}
    plan = lbc.ConvPlan(lbc.ConvDesc(**d.__dict__), options=opt)
    try:
        assert "tail-split" in plan.describe(), plan.describe()
    finally:
        plan.close()
    assert _check(d, options=opt) == "igemm_tc"
    assert _check(d, options={**opt, "tail_split": 0}) == "igemm_tc"


def test_tail_split_at_full_size_equals_the_unsplit_launch():
    """ResNet-50 stage 3 at the benchmark batch (392 pair-steps on 74 pairs): the launch with the split last round writes
    exactly the bytes of the launch without it."""
    import torch
    import lowbitdnn_project_b200 as lbc
    dev = torch.device("cuda:0")
    for d in (lbc.ConvDesc(n=512, h=14, w=14, c=256, k=256, r=3, s=3, pad_h=1, pad_w=1, relu=1),
              lbc.ConvDesc(n=512, h=14, w=14, c=1024, k=256, r=1, s=1, relu=1)):
        g = torch.Generator(device="cpu").manual_seed(7)
        x = torch.randint(-128, 128, (d.n, d.h, d.w, d.c), dtype=torch.int8, generator=g).to(dev)
        w = torch.randint(-127, 128, (d.k * d.r * d.s * d.c,), dtype=torch.int8, generator=g).to(dev)
        bias = torch.randint(-30000, 30000, (d.k,), dtype=torch.int32, generator=g).to(dev)
        scale = (torch.rand((d.k,), generator=g) * 1.5 + 0.5).mul(2.0**-7 / (d.r * d.s * d.c) ** 0.5).to(torch.float32).to(dev)
        outs = []
        for opt in ({"tail_split": 1}, {"tail_split": 0}, {"tail_split": 1, "reverse": 1}):
            plan = lbc.ConvPlan(d, options=opt)
            assert ("tail-split" in plan.describe()) == (opt["tail_split"] == 1), plan.describe()
            y = plan.empty_output(dev)
            plan.run(x, plan.prepack(w), bias, scale, out=y)
            plan.run(x, plan.prepack(w), bias, scale, out=y)
            torch.cuda.synchronize()
            plan.check_status()
            outs.append(y.clone())
            plan.close()
        assert torch.equal(outs[0], outs[1]) and torch.equal(outs[2], outs[1])


# ---- small-C tensor-core path (zero-pad + space-to-depth into 16-channel pixels) -------------------------
STEM_CASES = [
    D(n=2, h=32, w=32, c=3, k=64, r=7, s=7, stride_h=2, stride_w=2, pad_h=3, pad_w=3, relu=1),    # ResNet stem
    D(n=1, h=37, w=45, c=3, k=32, r=7, s=7, stride_h=2, stride_w=2, pad_h=3, pad_w=3),            # odd sizes
    D(n=2, h=30, w=30, c=3, k=32, r=3, s=3, stride_h=2, stride_w=2, pad_h=1, pad_w=1, relu=1),    # MobileNetV2 stem
    D(n=1, h=24, w=40, c=3, k=64, r=3, s=3, pad_h=1, pad_w=1, relu=1),                            # VGG conv1_1
    D(n=1, h=300, w=300, c=3, k=16, r=3, s=3, pad_h=1, pad_w=1),                                  # wide rows: column tiles
    D(n=2, h=20, w=20, c=4, k=16, r=5, s=5, stride_h=2, stride_w=2, pad_h=2, pad_w=2),
    D(n=2, h=17, w=19, c=8, k=32, r=3, s=3, pad_h=1, pad_w=1, relu=1),
    D(n=1, h=40, w=40, c=1, k=16, r=5, s=5, pad_h=2, pad_w=2),
]


@pytest.mark.parametrize("d", STEM_CASES, ids=lambda d: f"h{d.h}w{d.w}c{d.c}k{d.k}r{d.r}s{d.stride_h}")
@pytest.mark.parametrize("out_mode", [0, 1])
def test_stem_tc_kernel(d, out_mode):
    assert _check(D(**{**d.__dict__, "out_mode": out_mode})) == "stem_tc"


def test_stem_tc_oihw_weights():
    d = D(n=1, h=32, w=32, c=3, k=64, r=7, s=7, stride_h=2, stride_w=2, pad_h=3, pad_w=3, relu=1)
    assert _check(d, w_layout="oihw") == "stem_tc"


def test_igemm_reference_style_inputs_and_oihw_weights():
    """{0,1}-valued inputs (check.cu:43-44,69-75), weights handed over in the reference's OIHW order."""
    d = D(n=2, h=34, w=34, c=128, k=128, r=3, s=3, out_mode=1)
    _check(d, style="ref", force=IGEMM, use_bias=False, w_layout="oihw")


def test_igemm_matches_reference_golden():
    """The tensor-core path against outputs of the reference itself (tests/golden)."""
    from tests.test_oracle import iter_golden
    from tests.parity_util import run_gpu
    n_run = 0
    for key, shape, x, w, y in iter_golden():
        b, ic, ih, iw, oc, oh, ow, kh, kw = shape
        if ic % 16 or oc % 16:
            continue
        d = D(n=b, h=ih, w=iw, c=ic, k=oc, r=kh, s=kw, out_mode=1)
        got, name, _ = run_gpu(d, np.ascontiguousarray(x.transpose(0, 2, 3, 1)),
                               np.ascontiguousarray(w.transpose(0, 2, 3, 1)), None, None, force=IGEMM)
        assert name == "igemm_tc"
        assert np.array_equal(got.transpose(0, 3, 1, 2), y), key
        n_run += 1
    assert n_run >= 12


def test_direct_matches_reference_golden():
    from tests.test_oracle import iter_golden
    from tests.parity_util import run_gpu
    for key, shape, x, w, y in iter_golden():
        b, ic, ih, iw, oc, oh, ow, kh, kw = shape
        d = D(n=b, h=ih, w=iw, c=ic, k=oc, r=kh, s=kw, out_mode=1)
        got, _, _ = run_gpu(d, np.ascontiguousarray(x.transpose(0, 2, 3, 1)),
                            np.ascontiguousarray(w.transpose(0, 2, 3, 1)), None, None, force=DIRECT)
        assert np.array_equal(got.transpose(0, 3, 1, 2), y), key


def test_planner_choices():
    import lowbitdnn_project_b200 as lbc
    assert lbc.ConvPlan(lbc.ConvDesc(n=1, h=56, w=56, c=64, k=64, r=3, s=3, pad_h=1, pad_w=1)).kernel == "igemm_tc"
    assert lbc.ConvPlan(lbc.ConvDesc(n=1, h=56, w=56, c=64, k=256, r=1, s=1)).kernel == "igemm_tc"
    assert lbc.ConvPlan(lbc.ConvDesc(n=1, h=28, w=28, c=192, k=192, r=3, s=3, pad_h=1, pad_w=1, groups=192)).kernel == "depthwise"
    assert lbc.ConvPlan(lbc.ConvDesc(n=1, h=224, w=224, c=3, k=64, r=7, s=7, stride_h=2, stride_w=2, pad_h=3, pad_w=3)).kernel == "stem_tc"
    assert lbc.ConvPlan(lbc.ConvDesc(n=1, h=28, w=28, c=64, k=64, r=3, s=3, pad_h=1, pad_w=1, groups=4)).kernel == "direct"


def test_requant_extremes_on_device():
    """Saturation, ties and wraparound through the fused epilogue (bias/scale chosen to hit them)."""
    from tests.parity_util import run_gpu
    d = D(n=1, h=4, w=8, c=16, k=16, r=1, s=1, out_mode=0)
    rng = np.random.default_rng(5)
    x = rng.integers(-128, 128, size=(1, 4, 8, 16), dtype=np.int8)
    w = np.zeros((16, 1, 1, 16), dtype=np.int8)
    for k in range(16):
        w[k, 0, 0, k] = 1                                  # acc[k] = x[..., k]
    bias = np.array([0, 1, -1, 127, -128, 2**31 - 1, -2**31, 5, 0, 0, 0, 0, 3, 3, 3, 3], dtype=np.int32)
    scale = np.array([0.5, 0.5, 0.5, 1.0, 1.0, 1.0, 1.0, 1e-9, 1e9, -1.0, np.inf, np.nan, 0.25, 1.5, 2.5, 0.75],
                     dtype=np.float32)
    for relu in (0, 1):
        dd = D(**{**d.__dict__, "relu": relu})
        want = oracle.conv_nhwc(dd, x, w, bias, scale)
        for force in (DIRECT, IGEMM):
            got, _, _ = run_gpu(dd, x, w, bias, scale, force=force)
            assert np.array_equal(got, want), (relu, force)


def test_folded_bias_digit_boundaries():
    """The bias-in-MMA path on a layer where the planner enables it by default (resident filter matrix, 128-column tile):
    biases at the edges of the digit decomposition (multiples of 127, +-63/64 remainders, the +-500000 range limit and
    just beyond it, where the launch must fall back to the epilogue add) - raw int32 accumulators and requantised int8."""
    from tests.parity_util import run_gpu
    import lowbitdnn_project_b200 as lbc
    base = D(n=2, h=8, w=8, c=32, k=128, r=1, s=1)
    plan = lbc.ConvPlan(lbc.ConvDesc(**{**base.__dict__, "out_mode": 0}))
    assert "bias-in-mma" in plan.describe(), plan.describe()
    plan.close()
    rng = np.random.default_rng(11)
    x = rng.integers(-128, 128, size=(2, 8, 8, 32), dtype=np.int8)
    w = rng.integers(-127, 128, size=(128, 1, 1, 32), dtype=np.int8)
    edge = np.array([0, 1, -1, 63, 64, -63, -64, 126, 127, 128, -127, -128, 127 * 31, 127 * 31 + 63, 127 * 3937, 127 * 3937 + 63,
                     -127 * 3937 - 63, 499999, 500000, -500000, 190, 191, -190, -191, 32767, -32768, 16129, -16129, 8001, -8001,
                     254, -254], dtype=np.int32)
    scale = np.full(128, 2.0**-9, dtype=np.float32)
    for name, bias in (("in-range", np.resize(edge, 128)),
                       ("beyond", np.where(np.arange(128) == 77, 500001, np.resize(edge, 128)).astype(np.int32)),
                       ("int32-extremes", np.where(np.arange(128) % 2 == 0, 2**31 - 1, -2**31).astype(np.int32))):
        for out_mode, relu in ((1, 0), (0, 0), (0, 1)):
            dd = D(**{**base.__dict__, "out_mode": out_mode, "relu": relu})
            want = oracle.conv_nhwc(dd, x, w, bias, scale)
            got, kern, _ = run_gpu(dd, x, w, bias, scale if out_mode == 0 else None)
            assert kern == "igemm_tc" and np.array_equal(got, want), (name, out_mode, relu)


@pytest.mark.parametrize("d", [
    D(n=4, h=28, w=28, c=64, k=256, r=1, s=1, relu=1),                         # resident, per-warp stores, folded bias
    D(n=4, h=14, w=14, c=256, k=256, r=3, s=3, pad_h=1, pad_w=1, relu=1),      # CTA pairs
    D(n=4, h=28, w=28, c=64, k=64, r=3, s=3, pad_h=1, pad_w=1, relu=1),        # two MMA warps, two tiles per epilogue iteration
    D(n=4, h=28, w=28, c=128, k=512, r=1, s=1, relu=1),                        # resident filter matrix over two N tiles
], ids=lambda d: f"c{d.c}k{d.k}r{d.r}")
def test_repeated_launches_are_identical(d):
    """The same plan launched ten times back to back (no host synchronisation in between, so launches overlap through
    programmatic dependent launch) on a grid capped to 8 CTAs: every run must reproduce the oracle bit for bit - a race in
    the TMEM / staging / ring reuse or across the PDL boundary would show up as a run-to-run difference."""
    import torch
    import lowbitdnn_project_b200 as lbc
    from tests.parity_util import lbc_desc
    od = D(**{**d.__dict__, "out_mode": 0})
    x, w, bias, scale = oracle.synth(od, layer=3)
    want = oracle.conv_nhwc(od, x, w, bias, scale)
    dev = torch.device("cuda:0")
    plan = lbc.ConvPlan(lbc_desc(od), options={"max_grid": 8})
    wp = plan.prepack(torch.from_numpy(w).to(dev).reshape(-1), lbc.W_KRSC)
    xt, bt, st = torch.from_numpy(x).to(dev), torch.from_numpy(bias).to(dev), torch.from_numpy(scale).to(dev)
    outs = [plan.empty_output(dev) for _ in range(10)]
    for y in outs:
        plan.run(xt, wp, bt, st, out=y)
    torch.cuda.synchronize()
    plan.check_status()
    for i, y in enumerate(outs):
        assert np.array_equal(y.cpu().numpy(), want), f"launch {i} differs"
    plan.close()


# ---- fused bottleneck tail: conv(R x S -> 64) -> conv(1x1 -> 256) in one launch (r02) -------------------------
FUSED_CASES = [
    # (n, h, w, c_in, r, pad, relu_a, relu_b)
    (2, 56, 56, 64, 3, 1, 1, 1),        # ResNet-50 l1.x.conv2 -> conv3
    (3, 28, 28, 64, 3, 1, 1, 0),        # no ReLU after the 1x1 (the full-graph bottleneck)
    (2, 14, 14, 128, 3, 1, 1, 1),       # 128-byte K chunks, 8-row window tiles
    (1, 20, 224, 32, 3, 1, 1, 1),       # column tiles, 32-byte K chunks
    (3, 19, 23, 32, 5, 2, 0, 1),        # 5x5, ragged edges, no ReLU in between
    (2, 9, 40, 64, 1, 0, 1, 1),         # conv A itself a 1 x 3 filter below
]


@pytest.mark.parametrize("case", FUSED_CASES, ids=lambda c: "n%dh%dw%dc%dr%d" % c[:5])
def test_fused_tail_equals_the_two_layer_chain(case):
    """The fused launch must reproduce conv A -> int8 -> conv B bit for bit (oracle chain), for every window geometry the
    first convolution can have, repeated launches included."""
    import torch
    import lowbitdnn_project_b200 as lbc
    n, h, w, c, r, pad, relu_a, relu_b = case
    s = 3 if (r == 1) else r
    da = D(n=n, h=h, w=w, c=c, k=64, r=r, s=s, pad_h=pad, pad_w=(1 if r == 1 else pad), relu=relu_a)
    db = D(n=n, h=da.p, w=da.q, c=64, k=256, r=1, s=1, relu=relu_b)
    xa, wa, ba, sa = oracle.synth(da, layer=11)
    _, wb, bb, sb = oracle.synth(db, layer=12)
    mid = oracle.conv_nhwc(da, xa, wa, ba, sa)
    want = oracle.conv_nhwc(db, mid, wb, bb, sb)
    from tests.parity_util import lbc_desc
    dev = torch.device("cuda:0")
    plan = lbc.FusedTailPlan(lbc_desc(da), lbc_desc(db))
    wpa, wpb = plan.prepack(torch.from_numpy(wa).to(dev).reshape(-1), torch.from_numpy(wb).to(dev).reshape(-1))
    t = lambda a: torch.from_numpy(a).to(dev)
    x, tba, tsa, tbb, tsb = t(xa), t(ba), t(sa), t(bb), t(sb)
    outs = [plan.run(x, wpa, tba, tsa, wpb, tbb, tsb) for _ in range(3)]
    y, ms = plan.run(x, wpa, tba, tsa, wpb, tbb, tsb, timed=True)
    torch.cuda.synchronize()
    for i, o in enumerate(outs + [y]):
        got = o.cpu().numpy()
        if not np.array_equal(got, want):
            from tests.parity_util import mismatch_report
            raise AssertionError(f"launch {i}:\n" + mismatch_report(got, want))
    assert ms > 0
    plan.close()


def test_fused_tail_refuses_other_pairs():
    import lowbitdnn_project_b200 as lbc
    a = lbc.ConvDesc(n=1, h=14, w=14, c=64, k=64, r=3, s=3, pad_h=1, pad_w=1, relu=1)
    for bad_a, bad_b in [
        (a.replace(k=128), lbc.ConvDesc(n=1, h=14, w=14, c=128, k=256, r=1, s=1)),          # 128 channels in between
        (a, lbc.ConvDesc(n=1, h=14, w=14, c=64, k=128, r=1, s=1)),                           # 1x1 to 128
        (a.replace(stride_h=2, stride_w=2), lbc.ConvDesc(n=1, h=7, w=7, c=64, k=256, r=1, s=1)),   # stride 2: no window mode
        (a, lbc.ConvDesc(n=1, h=14, w=14, c=64, k=256, r=3, s=3, pad_h=1, pad_w=1)),         # second conv not 1x1
    ]:
        with pytest.raises(lbc.LbcError) as e:
            lbc.FusedTailPlan(bad_a, bad_b)
        assert e.value.status == 2
