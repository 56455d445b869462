// depthwise.cu — CUDA-core depthwise convolution (groups == C == K), NHWC int8.
//
// A depthwise layer is a per-channel stencil: 9 MACs and 2 bytes of HBM traffic per output, no contraction a
// 128-row MMA could use.  It is HBM-bound if the integer pipe keeps up, so the 3x3 kernel is built around the
// instruction count per output:
//
//   * a thread owns 4 consecutive channels (one 32-bit word per pixel; the lanes of a warp cover 128 consecutive
//     bytes of the NHWC row, so every load and store is a full line) and walks a strip of output pixels along W;
//   * for every input column it loads the 3 (stride 2) or 4 (stride 1) input rows and transposes them with 6-8
//     PRMTs into per-channel words whose BYTES are the vertically adjacent pixels (r0, r1, r2[, r3]);
//   * one __dp4a of such a word with the matching column of the filter (w[0][s], w[1][s], w[2][s], 0) does the
//     three vertical taps of one channel at once: 3 dp4a per output instead of 9 multiply-adds plus byte
//     extraction.  With stride 1 the fourth byte carries the next input row, and the same word dotted with
//     (0, w0, w1, w2) gives the output row below: two output rows per thread from one set of loads;
//   * the three live columns roll through registers (each input column is loaded once per thread);
//   * bias / scale of the thread's 4 channels sit in registers for the whole strip (no shared memory at all).
//
// Anything that is not 3x3 / dilation 1 / stride 1 or 2 runs the generic kernel at the bottom.
//
// Replaces nothing in the reference (it has no depthwise path: `groups` is accepted but never forwarded,
// python/qtorch/cpp/conv2d.cuh:96,140); the shape family comes from BASELINE.json config 5 (MobileNetV2).
#include "common.cuh"
#include "ptx.cuh"

#include <algorithm>
#include <mutex>

namespace lbc {

namespace {

struct DwParams {
    int32_t n, h, w, c, p, q;
    int32_t stride, pad_h, pad_w;
    int32_t cq;           // C / 4
    int32_t tw, strips;   // strip width (output pixels per thread along W), strips per output row
    int32_t row_groups;   // ceil(P / ROWS)
    int32_t relu, out_mode;
};

__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel)
{
    uint32_t r;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(sel));
    return r;
}

// rows (a, b, c, d) x 4 channels  ->  4 channel words of bytes (a_ch, b_ch, c_ch, d_ch)
template <bool FOUR>
__device__ __forceinline__ void rows_to_channels(uint32_t a, uint32_t b, uint32_t c, uint32_t d, uint32_t (&v)[4])
{
    const uint32_t t0 = prmt(a, b, 0x5140);   // a0 b0 a1 b1
    const uint32_t t1 = prmt(a, b, 0x7362);   // a2 b2 a3 b3
    if (FOUR) {
        const uint32_t u0 = prmt(c, d, 0x5140);   // c0 d0 c1 d1
        const uint32_t u1 = prmt(c, d, 0x7362);   // c2 d2 c3 d3
        v[0] = prmt(t0, u0, 0x5410);
        v[1] = prmt(t0, u0, 0x7632);
        v[2] = prmt(t1, u1, 0x5410);
        v[3] = prmt(t1, u1, 0x7632);
    } else {   // fourth byte: don't care (the filter word has a zero there)
        v[0] = prmt(t0, c, 0x4410);
        v[1] = prmt(t0, c, 0x5532);
        v[2] = prmt(t1, c, 0x6610);
        v[3] = prmt(t1, c, 0x7732);
    }
}

// STRIDE 1: two output rows per thread (4 input rows per column); STRIDE 2: one output row (3 input rows).
//
// Instruction diet (the kernel is issue-bound, not HBM-bound, until this is right):
//   * input rows outside the image are not predicated in the loop: their pointer is clamped to a valid row and the
//     matching BYTE of every filter word is zeroed once, so whatever is loaded there multiplies by zero;
//   * all addresses are 32-bit word offsets from the (warp-uniform) tensor base: one IMAD.WIDE per access;
//   * the bias is the initial value of the dp4a chain (same int32 wraparound as acc + bias);
//   * the strip loop is unrolled by the column-reuse period (3 for stride 1, 2 for stride 2) so the rolling window
//     is a renaming of registers, not a set of moves.
template <int STRIDE>
__global__ void __launch_bounds__(256) depthwise3x3_kernel(const DwParams g, const uint32_t* __restrict__ x32,
                                                           const int8_t* __restrict__ w_rsc,
                                                           const int32_t* __restrict__ bias,
                                                           const float* __restrict__ scale, void* __restrict__ y)
{
    constexpr int ROWS = (STRIDE == 1) ? 2 : 1;
    constexpr int IN_ROWS = (STRIDE == 1) ? 4 : 3;
    uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t cqi = idx % (uint32_t)g.cq;
    idx /= (uint32_t)g.cq;
    const uint32_t strip = idx % (uint32_t)g.strips;
    idx /= (uint32_t)g.strips;
    const uint32_t rg = idx % (uint32_t)g.row_groups;
    const uint32_t n = idx / (uint32_t)g.row_groups;
    if (n >= (uint32_t)g.n) return;
    const int32_t c0 = (int32_t)cqi * 4;
    const int32_t p0 = (int32_t)rg * ROWS;
    const int32_t q0 = (int32_t)strip * g.tw;
    const int32_t q1 = min(q0 + g.tw, g.q);

    // ---- input rows: clamped word offsets + a byte mask that removes out-of-image rows from the filter
    const int32_t ih0 = p0 * STRIDE - g.pad_h;
    uint32_t rowoff[IN_ROWS];
    uint32_t keep_mask = 0;
#pragma unroll
    for (int r = 0; r < IN_ROWS; ++r) {
        const int32_t ih = ih0 + r;
        const bool ok = ih >= 0 && ih < g.h;
        if (ok) keep_mask |= 0xffu << (8 * r);
        rowoff[r] = ((n * (uint32_t)g.h + (uint32_t)(ok ? ih : 0)) * (uint32_t)g.w) * (uint32_t)g.cq + cqi;
    }

    // ---- filter columns: wa[s][ch] = bytes (w[0][s][ch], w[1][s][ch], w[2][s][ch], 0) for output row p0 and
    //      wb = the same shifted up one byte for output row p0 + 1 (stride 1 only)
    uint32_t wa[3][4], wb[3][4];
#pragma unroll
    for (int s = 0; s < 3; ++s) {
        const uint32_t a = __ldg(reinterpret_cast<const uint32_t*>(w_rsc + (0 * 3 + s) * g.c + c0));
        const uint32_t b = __ldg(reinterpret_cast<const uint32_t*>(w_rsc + (1 * 3 + s) * g.c + c0));
        const uint32_t c = __ldg(reinterpret_cast<const uint32_t*>(w_rsc + (2 * 3 + s) * g.c + c0));
        rows_to_channels<false>(a, b, c, 0u, wa[s]);
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
            const uint32_t w3 = wa[s][ch] & 0x00ffffffu;
            wa[s][ch] = w3 & keep_mask;
            wb[s][ch] = (w3 << 8) & keep_mask;
        }
    }

    // ---- epilogue parameters of the 4 channels
    const int32_t lo = g.relu ? 0 : -128;
    int32_t bi[4] = {0, 0, 0, 0};
    float sc[4] = {0.f, 0.f, 0.f, 0.f};
    if (bias) {
        const int4 b = __ldg(reinterpret_cast<const int4*>(bias + c0));
        bi[0] = b.x; bi[1] = b.y; bi[2] = b.z; bi[3] = b.w;
    }
    if (g.out_mode == LBC_OUT_INT8) {
        const float4 s4 = __ldg(reinterpret_cast<const float4*>(scale + c0));
        sc[0] = s4.x; sc[1] = s4.y; sc[2] = s4.z; sc[3] = s4.w;
    }
    const bool row1 = (ROWS == 2) && (p0 + 1 < g.p);
    // word offset (4 channels = one 32-bit word of int8 output) of (n, p0, 0, c0)
    const uint32_t yoff0 = ((n * (uint32_t)g.p + (uint32_t)p0) * (uint32_t)g.q) * (uint32_t)g.cq + cqi;
    const uint32_t yoff1 = yoff0 + (uint32_t)g.q * (uint32_t)g.cq;
    uint32_t* y8 = reinterpret_cast<uint32_t*>(y);
    int4* y32 = reinterpret_cast<int4*>(y);

    // raw row words of one input column (zero outside the image); transposed later, one iteration after the load
    // was issued, so the memory latency is covered by the previous column's arithmetic
    auto load_raw = [&](int32_t iw, uint32_t(&rw)[4]) {
#pragma unroll
        for (int r = 0; r < 4; ++r) rw[r] = 0u;
        if ((uint32_t)iw < (uint32_t)g.w) {
            const uint32_t co = (uint32_t)iw * (uint32_t)g.cq;
#pragma unroll
            for (int r = 0; r < IN_ROWS; ++r) rw[r] = __ldg(x32 + (rowoff[r] + co));
        }
    };
    auto xpose = [&](const uint32_t(&rw)[4], uint32_t(&v)[4]) {
        rows_to_channels<(IN_ROWS == 4)>(rw[0], rw[1], rw[2], rw[3], v);
    };
    auto emit = [&](const int32_t(&acc)[4], uint32_t o) {
        if (g.out_mode == LBC_OUT_INT32) {
            y32[o] = make_int4(acc[0], acc[1], acc[2], acc[3]);
        } else {
            y8[o] = requant4_pack(acc, sc, lo == 0);
        }
    };
    auto compute = [&](const uint32_t(&va)[4], const uint32_t(&vb)[4], const uint32_t(&vc)[4], int32_t q) {
        int32_t a0[4], a1[4];
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
            a0[ch] = __dp4a((int32_t)va[ch], (int32_t)wa[0][ch], bi[ch]);
            a0[ch] = __dp4a((int32_t)vb[ch], (int32_t)wa[1][ch], a0[ch]);
            a0[ch] = __dp4a((int32_t)vc[ch], (int32_t)wa[2][ch], a0[ch]);
            if (ROWS == 2) {
                a1[ch] = __dp4a((int32_t)va[ch], (int32_t)wb[0][ch], bi[ch]);
                a1[ch] = __dp4a((int32_t)vb[ch], (int32_t)wb[1][ch], a1[ch]);
                a1[ch] = __dp4a((int32_t)vc[ch], (int32_t)wb[2][ch], a1[ch]);
            }
        }
        const uint32_t qo = (uint32_t)q * (uint32_t)g.cq;
        emit(a0, yoff0 + qo);
        if (ROWS == 2 && row1) emit(a1, yoff1 + qo);
    };

    uint32_t v0[4], v1[4], v2[4], ra[4], rb[4];
    int32_t iw = q0 * STRIDE - g.pad_w;
    if (STRIDE == 1) {
        load_raw(iw, ra);
        load_raw(iw + 1, rb);
        xpose(ra, v0);
        load_raw(iw + 2, ra);
        xpose(rb, v1);
        // invariant at the top of a step for output q: ra holds the raw words of column iw + 2
        for (int32_t q = q0; q < q1; q += 3, iw += 3) {
            load_raw(iw + 3, rb);
            xpose(ra, v2);
            compute(v0, v1, v2, q);
            if (q + 1 < q1) {
                load_raw(iw + 4, ra);
                xpose(rb, v0);
                compute(v1, v2, v0, q + 1);
            }
            if (q + 2 < q1) {
                load_raw(iw + 5, rb);
                xpose(ra, v1);
                compute(v2, v0, v1, q + 2);
#pragma unroll
                for (int r = 0; r < 4; ++r) ra[r] = rb[r];
            }
        }
    } else {
        uint32_t rc[4], rd[4];
        load_raw(iw, ra);
        load_raw(iw + 1, rb);
        load_raw(iw + 2, rc);
        xpose(ra, v0);
        // invariant: rb, rc hold the raw words of columns iw + 1, iw + 2
        for (int32_t q = q0; q < q1; ++q, iw += 2) {
            load_raw(iw + 3, ra);
            load_raw(iw + 4, rd);
            xpose(rb, v1);
            xpose(rc, v2);
            compute(v0, v1, v2, q);
#pragma unroll
            for (int r = 0; r < 4; ++r) { v0[r] = v2[r]; rb[r] = ra[r]; rc[r] = rd[r]; }
        }
    }
}

// ---- TMA-staged 3x3 kernel ---------------------------------------------------------------------------------
// The direct kernel above is latency-bound (ncu, r01: 48 % issue utilisation, every loop iteration waits on its own
// global loads).  Here a CTA stages one input tile - nb images x (th*S+2) rows x (tq*S+2) columns x cc channels - in
// shared memory with ONE rank-4 TMA load (out-of-image elements arrive as zeros, so there is no padding logic at all),
// several CTAs per SM keep ~100 KB of loads in flight, and the arithmetic is the same row-packed dp4a scheme reading
// shared memory with constant offsets.
struct DwTiledParams {
    int32_t n, p, q, cq_total;     // output extents, C / 4
    int32_t cc, ccq;               // channels of a tile, / 4
    int32_t th, tq, nb, tw;        // output rows / cols / images per tile, strip width
    int32_t in_h, in_w;            // input rows / cols of a tile
    int32_t tiles_c, tiles_q, tiles_p;
    int32_t pad_h, pad_w;
    int32_t strips, row_groups;    // per tile
    int32_t relu, out_mode;
    uint32_t tile_bytes;
    int32_t reverse;               // visit the tiles last-to-first (see IgemmLaunch::reverse)
    int* flag;                     // the plan's device watchdog word
};

template <int STRIDE>
__global__ void __launch_bounds__(256) depthwise3x3_tiled_kernel(const __grid_constant__ CUtensorMap tm_x, const DwTiledParams g,
                                                                 const int8_t* __restrict__ w_rsc,
                                                                 const int32_t* __restrict__ bias,
                                                                 const float* __restrict__ scale, void* __restrict__ y)
{
    constexpr int ROWS = (STRIDE == 1) ? 2 : 1;
    constexpr int IN_ROWS = (STRIDE == 1) ? 4 : 3;
    extern __shared__ __align__(128) uint8_t tile[];
    __shared__ uint64_t bar;

    // tile coordinates (blocks are dispatched in index order, so mirroring the index mirrors the traversal)
    uint32_t t = g.reverse ? gridDim.x - 1u - blockIdx.x : blockIdx.x;
    const int32_t tc = (int32_t)(t % (uint32_t)g.tiles_c); t /= (uint32_t)g.tiles_c;
    const int32_t tq_i = (int32_t)(t % (uint32_t)g.tiles_q); t /= (uint32_t)g.tiles_q;
    const int32_t tp_i = (int32_t)(t % (uint32_t)g.tiles_p); t /= (uint32_t)g.tiles_p;
    const int32_t n0 = (int32_t)t * g.nb, p0 = tp_i * g.th, q0 = tq_i * g.tq, c0 = tc * g.cc;

    if (threadIdx.x == 0) {
        ptx::mbar_init(&bar, 1);
        ptx::fence_barrier_init();
        ptx::mbar_expect_tx(&bar, g.tile_bytes);
        ptx::tma_load_4d(tile, &tm_x, &bar, c0, q0 * STRIDE - g.pad_w, p0 * STRIDE - g.pad_h, n0);
    }
    __syncthreads();
    // a watchdog trip (tile never landed) must not store results computed from unfilled shared memory
    if (!ptx::mbar_wait(&bar, 0, g.flag)) return;

    const uint32_t* tile32 = reinterpret_cast<const uint32_t*>(tile);
    const int32_t lo = g.relu ? 0 : -128;
    uint32_t* y8 = reinterpret_cast<uint32_t*>(y);
    int4* y32 = reinterpret_cast<int4*>(y);
    const int32_t items = g.nb * g.row_groups * g.strips * g.ccq;
    for (int32_t item = (int32_t)threadIdx.x; item < items; item += (int32_t)blockDim.x) {
        uint32_t r = (uint32_t)item;
        const int32_t cqi = (int32_t)(r % (uint32_t)g.ccq); r /= (uint32_t)g.ccq;
        const int32_t strip = (int32_t)(r % (uint32_t)g.strips); r /= (uint32_t)g.strips;
        const int32_t rg = (int32_t)(r % (uint32_t)g.row_groups);
        const int32_t img = (int32_t)(r / (uint32_t)g.row_groups);
        const int32_t n = n0 + img, pl = rg * ROWS, ql0 = strip * g.tw;
        const int32_t pg = p0 + pl;                       // global output row of the (first) row of this item
        if (n >= g.n || pg >= g.p || q0 + ql0 >= g.q) continue;
        const int32_t cg4 = tc * g.ccq + cqi;             // global channel quad
        const int32_t ch0 = cg4 * 4;

        // filter columns (see depthwise3x3_kernel); reloaded per item: keeping them across items (a thread bound to
        // one channel quad) measured slower on the stride-2 layers (r01)
        uint32_t wa[3][4], wb[3][4];
#pragma unroll
        for (int s = 0; s < 3; ++s) {
            const uint32_t a = __ldg(reinterpret_cast<const uint32_t*>(w_rsc + (0 * 3 + s) * (g.cq_total * 4) + ch0));
            const uint32_t b = __ldg(reinterpret_cast<const uint32_t*>(w_rsc + (1 * 3 + s) * (g.cq_total * 4) + ch0));
            const uint32_t c = __ldg(reinterpret_cast<const uint32_t*>(w_rsc + (2 * 3 + s) * (g.cq_total * 4) + ch0));
            rows_to_channels<false>(a, b, c, 0u, wa[s]);
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) {
                wa[s][ch] &= 0x00ffffffu;
                wb[s][ch] = wa[s][ch] << 8;
            }
        }
        int32_t bi[4] = {0, 0, 0, 0};
        float sc[4] = {0.f, 0.f, 0.f, 0.f};
        if (bias) {
            const int4 b = __ldg(reinterpret_cast<const int4*>(bias + ch0));
            bi[0] = b.x; bi[1] = b.y; bi[2] = b.z; bi[3] = b.w;
        }
        if (g.out_mode == LBC_OUT_INT8) {
            const float4 s4 = __ldg(reinterpret_cast<const float4*>(scale + ch0));
            sc[0] = s4.x; sc[1] = s4.y; sc[2] = s4.z; sc[3] = s4.w;
        }
        // shared-memory word offset of (img, input row pl*S, input col ql0*S, this channel quad)
        const uint32_t row_words = (uint32_t)(g.in_w * g.ccq);
        const uint32_t base = ((uint32_t)(img * g.in_h + pl * STRIDE) * (uint32_t)g.in_w + (uint32_t)(ql0 * STRIDE)) * (uint32_t)g.ccq + (uint32_t)cqi;
        auto load_col = [&](int32_t col, uint32_t(&v)[4]) {
            const uint32_t o = base + (uint32_t)col * (uint32_t)g.ccq;
            uint32_t rw[4] = {0u, 0u, 0u, 0u};
#pragma unroll
            for (int rr = 0; rr < IN_ROWS; ++rr) rw[rr] = tile32[o + (uint32_t)rr * row_words];
            rows_to_channels<(IN_ROWS == 4)>(rw[0], rw[1], rw[2], rw[3], v);
        };
        const bool row1 = (ROWS == 2) && (pg + 1 < g.p) && (pl + 1 < g.th);
        const uint32_t yoff0 = ((uint32_t)(n * g.p + pg) * (uint32_t)g.q + (uint32_t)(q0 + ql0)) * (uint32_t)g.cq_total + (uint32_t)cg4;
        const uint32_t yoff1 = yoff0 + (uint32_t)g.q * (uint32_t)g.cq_total;
        auto emit = [&](const int32_t(&acc)[4], uint32_t o) {
            if (g.out_mode == LBC_OUT_INT32) y32[o] = make_int4(acc[0], acc[1], acc[2], acc[3]);
            else y8[o] = requant4_pack(acc, sc, lo == 0);
        };
        auto compute = [&](const uint32_t(&va)[4], const uint32_t(&vb)[4], const uint32_t(&vc)[4], int32_t j) {
            int32_t a0[4], a1[4];
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) {
                a0[ch] = __dp4a((int32_t)va[ch], (int32_t)wa[0][ch], bi[ch]);
                a0[ch] = __dp4a((int32_t)vb[ch], (int32_t)wa[1][ch], a0[ch]);
                a0[ch] = __dp4a((int32_t)vc[ch], (int32_t)wa[2][ch], a0[ch]);
                if (ROWS == 2) {
                    a1[ch] = __dp4a((int32_t)va[ch], (int32_t)wb[0][ch], bi[ch]);
                    a1[ch] = __dp4a((int32_t)vb[ch], (int32_t)wb[1][ch], a1[ch]);
                    a1[ch] = __dp4a((int32_t)vc[ch], (int32_t)wb[2][ch], a1[ch]);
                }
            }
            const uint32_t qo = (uint32_t)j * (uint32_t)g.cq_total;
            emit(a0, yoff0 + qo);
            if (ROWS == 2 && row1) emit(a1, yoff1 + qo);
        };
        const int32_t jn = min(g.tw, g.q - (q0 + ql0));     // output columns of this strip inside the image
        uint32_t v0[4], v1[4], v2[4];
        load_col(0, v0);
        if (STRIDE == 1) {
            load_col(1, v1);
            for (int32_t j = 0; j < jn; j += 3) {
                load_col(j + 2, v2);
                compute(v0, v1, v2, j);
                if (j + 1 < jn) { load_col(j + 3, v0); compute(v1, v2, v0, j + 1); }
                if (j + 2 < jn) { load_col(j + 4, v1); compute(v2, v0, v1, j + 2); }
            }
        } else {
            for (int32_t j = 0; j < jn; j += 2) {
                load_col(2 * j + 1, v1);
                load_col(2 * j + 2, v2);
                compute(v0, v1, v2, j);
                if (j + 1 < jn) {
                    load_col(2 * j + 3, v1);
                    load_col(2 * j + 4, v0);
                    compute(v2, v1, v0, j + 1);
                }
            }
        }
    }
}

// ---- generic depthwise (any R, S, stride, dilation): one thread = one output pixel x 4 channels ------------
struct DwGenericParams {
    int32_t n, h, w, c, r, s, stride_h, stride_w, pad_h, pad_w, dil_h, dil_w, p, q, cq;
    int32_t relu, out_mode;
    int64_t total;
};

__device__ __forceinline__ int32_t sbyte(uint32_t v, int j) { return (int32_t)(int8_t)(v >> (8 * j)); }

__global__ void __launch_bounds__(256) depthwise_generic_kernel(const DwGenericParams g, const int8_t* __restrict__ x,
                                                                const int8_t* __restrict__ w_rsc,
                                                                const int32_t* __restrict__ bias,
                                                                const float* __restrict__ scale, void* __restrict__ y)
{
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= g.total) return;
    const int32_t c0 = (int32_t)(idx % g.cq) * 4;
    const int64_t m = idx / g.cq;
    const int32_t q = (int32_t)(m % g.q);
    const int32_t p = (int32_t)((m / g.q) % g.p);
    const int32_t n = (int32_t)(m / ((int64_t)g.q * g.p));

    int32_t acc[4] = {0, 0, 0, 0};
    for (int32_t r = 0; r < g.r; ++r) {
        const int32_t ih = p * g.stride_h - g.pad_h + r * g.dil_h;
        if (ih < 0 || ih >= g.h) continue;
        for (int32_t s = 0; s < g.s; ++s) {
            const int32_t iw = q * g.stride_w - g.pad_w + s * g.dil_w;
            if (iw < 0 || iw >= g.w) continue;
            const uint32_t xv = *reinterpret_cast<const uint32_t*>(x + (((int64_t)n * g.h + ih) * g.w + iw) * g.c + c0);
            const uint32_t wv = __ldg(reinterpret_cast<const uint32_t*>(w_rsc + ((int64_t)r * g.s + s) * g.c + c0));
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[j] += sbyte(xv, j) * sbyte(wv, j);
        }
    }
    const int64_t o = m * g.c + c0;
    const int32_t lo = g.relu ? 0 : -128;
    if (g.out_mode == LBC_OUT_INT32) {
        int4 v;
        v.x = acc[0] + (bias ? bias[c0 + 0] : 0);
        v.y = acc[1] + (bias ? bias[c0 + 1] : 0);
        v.z = acc[2] + (bias ? bias[c0 + 2] : 0);
        v.w = acc[3] + (bias ? bias[c0 + 3] : 0);
        *reinterpret_cast<int4*>(reinterpret_cast<int32_t*>(y) + o) = v;
    } else {
        int32_t b[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) b[j] = requant_s32(acc[j], bias ? bias[c0 + j] : 0, scale[c0 + j], lo);
        *reinterpret_cast<uint32_t*>(reinterpret_cast<int8_t*>(y) + o) = pack4_sat_s8(b[0], b[1], b[2], b[3]);
    }
}

// strip width: whole rows when short, else a divisor of Q in [12, 28], else 16 (the last strip is clipped)
int32_t pick_strip(int32_t q)
{
    if (q <= 28) return q;
    for (int32_t t = 28; t >= 12; --t)
        if (q % t == 0) return t;
    return 16;
}

}  // namespace

// Tile geometry of the TMA-staged kernel; tiled == 0 when the shape does not qualify.
lbc_status depthwise_encode(const ConvGeom& g, const lbc_plan_options& opt, const int8_t* x, DwLaunch* out)
{
    const lbc_conv_desc& d = g.d;
    *out = DwLaunch{};
    const bool fast = d.r == 3 && d.s == 3 && d.dil_h == 1 && d.dil_w == 1 && d.stride_h == d.stride_w &&
                      (d.stride_h == 1 || d.stride_h == 2) && d.c % 16 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 &&
                      (int64_t)d.n * d.h * d.w * d.c < (1ll << 32) && g.m_total * d.c < (1ll << 32) && opt.dw_tiled != 0;
    if (!fast) return LBC_OK;
    const int S = d.stride_h, rows = S == 1 ? 2 : 1;
    DwLaunch l{};
    // channels per tile: the largest multiple of 16 that divides C, up to 128
    l.cc = 16;
    for (int c = 128; c >= 16; c -= 16)
        if (d.c % c == 0) { l.cc = c; break; }
    const int ccq = l.cc / 4;
    // strips of 7 or 8 output columns, up to 4 strips (28..32 columns) per tile
    l.tw = (g.q % 7 == 0) ? 7 : (g.q < 8 ? g.q : 8);
    const int strips = std::min(4, (g.q + l.tw - 1) / l.tw);
    l.tq = strips * l.tw;
    // row groups: aim at ~3 passes of the 256 threads over the tile's work items, within the image
    const int rg_img = (g.p + rows - 1) / rows;
    int rg = std::max(1, std::min(rg_img, 768 / (strips * ccq)));
    l.th = rg * rows;
    l.in_w = (l.tq - 1) * S + 3;
    l.in_h = (l.th - 1) * S + 3;
    // small images: several images per tile so that the CTA has work for all its threads
    l.nb = 1;
    if (rg >= rg_img && strips * l.tw >= g.q) {
        const int items = rg * strips * ccq;
        l.nb = std::max(1, std::min(std::min(d.n, 8), 512 / std::max(1, items)));
    }
    while (l.nb > 1 && (size_t)l.nb * l.in_h * l.in_w * l.cc > 96u * 1024u) --l.nb;
    while (rg > 1 && (size_t)l.nb * l.in_h * l.in_w * l.cc > 96u * 1024u) {
        --rg;
        l.th = rg * rows;
        l.in_h = (l.th - 1) * S + 3;
    }
    l.tile_bytes = (uint32_t)((size_t)l.nb * l.in_h * l.in_w * l.cc);
    if (l.tile_bytes > 96u * 1024u || l.in_w > 256 || l.in_h > 256) return LBC_OK;   // direct kernel
    l.tiles_c = d.c / l.cc;
    l.tiles_q = (g.q + l.tq - 1) / l.tq;
    l.tiles_p = (g.p + l.th - 1) / l.th;
    l.tiles_n = (d.n + l.nb - 1) / l.nb;
    if ((int64_t)l.tiles_c * l.tiles_q * l.tiles_p * l.tiles_n >= (1ll << 31)) return LBC_OK;
    const uint64_t dims[4] = {(uint64_t)d.c, (uint64_t)d.w, (uint64_t)d.h, (uint64_t)d.n};
    const uint32_t box[4] = {(uint32_t)l.cc, (uint32_t)l.in_w, (uint32_t)l.in_h, (uint32_t)l.nb};
    lbc_status st = encode_tiled_u8_4d(&l.tm_x, x, dims, box);
    if (st != LBC_OK) return st;
    l.tiled = 1;
    *out = l;
    return LBC_OK;
}

lbc_status launch_depthwise(const ConvGeom& g, const int8_t* x, const int8_t* w_rsc, const EpilogueParams& ep,
                            void* y, const DwLaunch* dw, int* flag, cudaStream_t stream)
{
    const lbc_conv_desc& d = g.d;
    if (dw && dw->tiled) {
        LBC_REQUIRE((reinterpret_cast<uintptr_t>(y) & 15) == 0 && (reinterpret_cast<uintptr_t>(w_rsc) & 3) == 0, LBC_ERR_INVALID_ARG,
                    "depthwise: w must be 4-byte aligned and y 16-byte aligned");
        DwTiledParams p{};
        p.n = d.n; p.p = g.p; p.q = g.q; p.cq_total = d.c / 4;
        p.cc = dw->cc; p.ccq = dw->cc / 4;
        p.th = dw->th; p.tq = dw->tq; p.nb = dw->nb; p.tw = dw->tw; p.in_h = dw->in_h; p.in_w = dw->in_w;
        p.tiles_c = dw->tiles_c; p.tiles_q = dw->tiles_q; p.tiles_p = dw->tiles_p;
        p.pad_h = d.pad_h; p.pad_w = d.pad_w;
        p.strips = dw->tq / dw->tw;
        p.row_groups = dw->th / (d.stride_h == 1 ? 2 : 1);
        p.relu = ep.relu; p.out_mode = ep.out_mode;
        p.tile_bytes = dw->tile_bytes;
        p.reverse = dw->reverse ? 1 : 0;
        p.flag = flag;
        const unsigned grid = (unsigned)(dw->tiles_c * dw->tiles_q * dw->tiles_p * dw->tiles_n);
        const size_t smem = dw->tile_bytes;
        const unsigned block = 256;
        {   // the attribute is per device
            static std::mutex mu;
            static uint64_t done[4] = {0, 0, 0, 0};
            int dev_ord = 0;
            LBC_CUDA_TRY(cudaGetDevice(&dev_ord));
            std::lock_guard<std::mutex> lk(mu);
            if (first_use_on_device(done, dev_ord)) {
                LBC_CUDA_TRY(cudaFuncSetAttribute(depthwise3x3_tiled_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
                LBC_CUDA_TRY(cudaFuncSetAttribute(depthwise3x3_tiled_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
            }
        }
        if (d.stride_h == 1)
            depthwise3x3_tiled_kernel<1><<<grid, block, smem, stream>>>(dw->tm_x, p, w_rsc, ep.bias, ep.scale, y);
        else
            depthwise3x3_tiled_kernel<2><<<grid, block, smem, stream>>>(dw->tm_x, p, w_rsc, ep.bias, ep.scale, y);
        LBC_CUDA_TRY(cudaGetLastError());
        return LBC_OK;
    }
    LBC_REQUIRE(d.groups == d.c && d.k == d.c && (d.c % 4) == 0, LBC_ERR_UNSUPPORTED,
                "depthwise kernel needs groups == C == K and C %% 4 == 0");
    LBC_REQUIRE((reinterpret_cast<uintptr_t>(x) & 3) == 0 && (reinterpret_cast<uintptr_t>(y) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(w_rsc) & 3) == 0,
                LBC_ERR_INVALID_ARG, "depthwise: x / w must be 4-byte aligned and y 16-byte aligned");
    // the fast kernel addresses x and y with 32-bit word offsets
    const bool small = (int64_t)d.n * d.h * d.w * d.c < (1ll << 32) && g.m_total * d.c < (1ll << 32);
    const bool fast = d.r == 3 && d.s == 3 && d.dil_h == 1 && d.dil_w == 1 && d.stride_h == d.stride_w &&
                      (d.stride_h == 1 || d.stride_h == 2) && small;
    const int block = 256;
    if (fast) {
        DwParams p{};
        p.n = d.n; p.h = d.h; p.w = d.w; p.c = d.c; p.p = g.p; p.q = g.q;
        p.stride = d.stride_h; p.pad_h = d.pad_h; p.pad_w = d.pad_w;
        p.cq = d.c / 4;
        p.tw = pick_strip(g.q);
        p.strips = (g.q + p.tw - 1) / p.tw;
        const int rows = d.stride_h == 1 ? 2 : 1;
        p.row_groups = (g.p + rows - 1) / rows;
        p.relu = ep.relu; p.out_mode = ep.out_mode;
        const int64_t threads = (int64_t)d.n * p.row_groups * p.strips * p.cq;
        LBC_REQUIRE(threads < (1ll << 31), LBC_ERR_UNSUPPORTED, "depthwise: problem too large (%lld threads)", (long long)threads);
        const unsigned grid = (unsigned)((threads + block - 1) / block);
        if (d.stride_h == 1)
            depthwise3x3_kernel<1><<<grid, block, 0, stream>>>(p, reinterpret_cast<const uint32_t*>(x), w_rsc, ep.bias, ep.scale, y);
        else
            depthwise3x3_kernel<2><<<grid, block, 0, stream>>>(p, reinterpret_cast<const uint32_t*>(x), w_rsc, ep.bias, ep.scale, y);
    } else {
        DwGenericParams p{};
        p.n = d.n; p.h = d.h; p.w = d.w; p.c = d.c; p.r = d.r; p.s = d.s;
        p.stride_h = d.stride_h; p.stride_w = d.stride_w; p.pad_h = d.pad_h; p.pad_w = d.pad_w;
        p.dil_h = d.dil_h; p.dil_w = d.dil_w; p.p = g.p; p.q = g.q; p.cq = d.c / 4;
        p.relu = ep.relu; p.out_mode = ep.out_mode;
        p.total = g.m_total * p.cq;
        const int64_t grid = (p.total + block - 1) / block;
        LBC_REQUIRE(grid <= 0x7fffffffLL, LBC_ERR_UNSUPPORTED, "depthwise: grid too large");
        depthwise_generic_kernel<<<(unsigned)grid, block, 0, stream>>>(p, x, w_rsc, ep.bias, ep.scale, y);
    }
    LBC_CUDA_TRY(cudaGetLastError());
    return LBC_OK;
}

}  // namespace lbc
