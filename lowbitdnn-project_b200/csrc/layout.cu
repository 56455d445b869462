// layout.cu — weight pre-packing and the reference's tensor-format converters.
//
//   to_vect_c / from_vect_c : cpp/int8conv/utils.cuh:11-26  ([N,C,H,W] <-> [N,C/V,H,W,V]); the reference
//   leaves these as views and pays a hidden .contiguous() transpose inside the op
//   (conv2DForward3x3TensorCores.cuh:715-716); here they are explicit, stream-ordered kernels.
#include "common.cuh"

namespace lbc {

namespace {

struct Permute5 {
    int32_t ddim[5];      // destination extents
    int64_t sstride[5];   // source stride (in elements) of the source dim feeding destination dim j
};

template <typename T>
__global__ void __launch_bounds__(256) permute5_kernel(Permute5 pm, const T* __restrict__ src, T* __restrict__ dst,
                                                       int64_t total)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t rem = i, so = 0;
#pragma unroll
        for (int j = 4; j >= 0; --j) {
            const int64_t idx = rem % pm.ddim[j];
            rem /= pm.ddim[j];
            so += idx * pm.sstride[j];
        }
        dst[i] = src[so];
    }
}

// dst [K][R][S][c_pad] <- src KRSC [K][R][S][cg] or OIHW [K][cg][R][S]; channels >= cg are zero.
__global__ void __launch_bounds__(256) prepack_krsc_kernel(const int8_t* __restrict__ src, int32_t oihw,
                                                           int8_t* __restrict__ dst, int32_t k, int32_t r, int32_t s,
                                                           int32_t cg, int32_t c_pad)
{
    const int64_t total = (int64_t)k * r * s * c_pad;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int32_t c = (int32_t)(i % c_pad);
        int64_t t = i / c_pad;
        const int32_t is = (int32_t)(t % s); t /= s;
        const int32_t ir = (int32_t)(t % r); t /= r;
        const int32_t ik = (int32_t)t;
        int8_t v = 0;
        if (c < cg) {
            const int64_t so = oihw ? ((((int64_t)ik * cg + c) * r + ir) * s + is)
                                    : ((((int64_t)ik * r + ir) * s + is) * cg + c);
            v = src[so];
        }
        dst[i] = v;
    }
}

// tcgen05 filter matrix (see common.cuh::launch_prepack_igemm).
__global__ void __launch_bounds__(256) prepack_igemm_kernel(const int8_t* __restrict__ src, int32_t oihw,
                                                            int8_t* __restrict__ dst, int32_t k, int32_t r, int32_t s,
                                                            int32_t cg, int32_t s_pad, int32_t bkc, int32_t cblocks,
                                                            int32_t chunk_outer)
{
    const int32_t taps = r * s_pad;
    const int64_t row_bytes = (int64_t)taps * cblocks * bkc;
    const int64_t total = (int64_t)k * row_bytes;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int32_t ik = (int32_t)(i / row_bytes);
        const int32_t rest = (int32_t)(i - (int64_t)ik * row_bytes);
        const int32_t cin = rest % bkc;
        int32_t tap, cb;
        if (chunk_outer) { cb = rest / (taps * bkc); tap = (rest / bkc) % taps; }
        else             { tap = rest / (cblocks * bkc); cb = (rest / bkc) % cblocks; }
        const int32_t c = cb * bkc + cin;
        const int32_t ir = tap / s_pad, is = tap % s_pad;
        int8_t v = 0;
        if (c < cg && is < s) {
            const int64_t so = oihw ? ((((int64_t)ik * cg + c) * r + ir) * s + is)
                                    : ((((int64_t)ik * r + ir) * s + is) * cg + c);
            v = src[so];
        }
        dst[i] = v;
    }
}

// Pixel-group rewrite of a pointwise convolution (see api.cu): the [K][C] filter becomes the block-diagonal
// [f*K][f*C] matrix  W'[j*K + k][i*C + c] = (i == j) ? W[k][c] : 0, so that f consecutive pixels, seen as ONE GEMM row
// of f*C channels, produce their f*K outputs side by side - which is exactly the NHWC order of the f pixels.
__global__ void __launch_bounds__(256) blockdiag_kernel(const int8_t* __restrict__ src, int8_t* __restrict__ dst, int32_t k,
                                                        int32_t c, int32_t f)
{
    const int64_t total = (int64_t)f * k * f * c;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int32_t col = (int32_t)(i % ((int64_t)f * c));
        const int32_t row = (int32_t)(i / ((int64_t)f * c));
        const int32_t jb = row / k, kk = row - jb * k, ib = col / c, cc = col - ib * c;
        dst[i] = (ib == jb) ? src[(int64_t)kk * c + cc] : (int8_t)0;
    }
}

// Small-C ("stem") path.  A conv with C*sh*sw <= 16 input channels and stride (sh, sw) in {1, 2} is rewritten as a
// stride-1, pad-0 conv over 16-channel pixels: the input is zero-padded and (for stride 2) space-to-depth'd,
//   X'[n][hs][ws][(dr*sw + ds)*C + c] = Xpad[n][hs*sh + dr][ws*sw + ds][c],
// and the filter likewise, W'[k][r2][s2][(dr*sw + ds)*C + c] = W[k][r2*sh + dr][s2*sw + ds][c] (0 outside R x S).
struct StemXformParams {
    int32_t n, h, w, c, hs, ws, sh, sw, pad_h, pad_w;
};

// grid (ceil(Ws / 128), ceil(Hs / kXformRows), N): one thread = one column of kXformRows 16-byte output pixels; no
// divisions, and the loads of several output rows are in flight together (a one-pixel-per-thread version was bound by
// block turnover: every block was a single load -> store dependency chain).  A warp reads sw*C*32 consecutive bytes
// of each input row it needs (192 for the 7x7/stride-2 RGB stem) and writes 512 consecutive bytes per output row.
constexpr int kXformRows = 8;

// C, SH, SW > 0: compile-time channel count and strides (every loop unrolls, every byte slot is a constant: ~50
// instructions per output pixel instead of ~470 with run-time bounds); C == 0: generic run-time version.
template <int C, int SH, int SW>
__global__ void __launch_bounds__(128) stem_xform_kernel(StemXformParams p, const int8_t* __restrict__ x,
                                                         uint4* __restrict__ out)
{
    // the convolution that follows is launched with programmatic stream serialisation: let its CTAs set up (and fetch their
    // filter matrix) while this pass runs; its griddepcontrol.wait still orders every read of `out` after this grid
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const int32_t ws = (int32_t)(blockIdx.x * blockDim.x + threadIdx.x);
    if (ws >= p.ws) return;
    const int32_t hs0 = (int32_t)blockIdx.y * kXformRows, n = (int32_t)blockIdx.z;
    const int32_t pc = C ? C : p.c, psh = C ? SH : p.sh, psw = C ? SW : p.sw;
    const int32_t iw0 = ws * psw - p.pad_w;
    const uint8_t* img = reinterpret_cast<const uint8_t*>(x) + (int64_t)n * p.h * p.w * pc;
    uint4* dst = out + ((int64_t)n * p.hs + hs0) * p.ws + ws;
#pragma unroll 4
    for (int j = 0; j < kXformRows; ++j) {
        const int32_t hs = hs0 + j;
        if (hs >= p.hs) break;
        if (C) {
            uint32_t v[4] = {0, 0, 0, 0};
#pragma unroll
            for (int dr = 0; dr < (C ? SH : 1); ++dr) {
                const int32_t ih = hs * SH + dr - p.pad_h;
                const bool row_in = ih >= 0 && ih < p.h;
                const uint8_t* row = img + (row_in ? ih : 0) * (p.w * C);
#pragma unroll
                for (int ds = 0; ds < (C ? SW : 1); ++ds) {
                    const int32_t iw = iw0 + ds;
                    const bool in = row_in && iw >= 0 && iw < p.w;
                    const uint8_t* src = row + (in ? iw : 0) * C;
#pragma unroll
                    for (int c = 0; c < (C ? C : 1); ++c) {
                        const int slot = (dr * SW + ds) * C + c;
                        const uint32_t b = in ? (uint32_t)__ldg(src + c) : 0u;
                        v[slot >> 2] |= b << (8 * (slot & 3));
                    }
                }
            }
            dst[(int64_t)j * p.ws] = make_uint4(v[0], v[1], v[2], v[3]);
        } else {
            // the 16 output bytes are assembled in two 64-bit registers (an indexed array would live in local memory)
            uint64_t lo = 0, hi = 0;
            int slot = 0;
            for (int dr = 0; dr < psh; ++dr) {
                const int32_t ih = hs * psh + dr - p.pad_h;
                const bool row_in = ih >= 0 && ih < p.h;
                const uint8_t* row = img + (row_in ? ih : 0) * (p.w * pc);
                for (int ds = 0; ds < psw; ++ds) {
                    const int32_t iw = iw0 + ds;
                    const bool in = row_in && iw >= 0 && iw < p.w;
                    const uint8_t* src = row + (in ? iw : 0) * pc;
                    for (int c = 0; c < pc; ++c, ++slot) {
                        const uint64_t b = in ? (uint64_t)__ldg(src + c) : 0ull;
                        if (slot < 8) lo |= b << (8 * slot);
                        else hi |= b << (8 * (slot - 8));
                    }
                }
            }
            dst[(int64_t)j * p.ws] = make_uint4((uint32_t)lo, (uint32_t)(lo >> 32), (uint32_t)hi, (uint32_t)(hi >> 32));
        }
    }
}

// dst [K][r2][s_pad][16] <- src KRSC [K][R][S][C] or OIHW [K][C][R][S]
__global__ void __launch_bounds__(256) prepack_stem_kernel(const int8_t* __restrict__ src, int32_t oihw,
                                                           int8_t* __restrict__ dst, int32_t k, int32_t r, int32_t s,
                                                           int32_t c, int32_t sh, int32_t sw, int32_t r2, int32_t s_pad)
{
    const int64_t total = (int64_t)k * r2 * s_pad * 16;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int32_t ch = (int32_t)(i % 16);
        int64_t t = i / 16;
        const int32_t is2 = (int32_t)(t % s_pad); t /= s_pad;
        const int32_t ir2 = (int32_t)(t % r2); t /= r2;
        const int32_t ik = (int32_t)t;
        int8_t v = 0;
        if (ch < sh * sw * c) {
            const int32_t sub = ch / c, cc = ch % c;
            const int32_t dr = sub / sw, ds = sub % sw;
            const int32_t ir = ir2 * sh + dr, is = is2 * sw + ds;
            if (ir < r && is < s) {
                const int64_t so = oihw ? ((((int64_t)ik * c + cc) * r + ir) * s + is)
                                        : ((((int64_t)ik * r + ir) * s + is) * c + cc);
                v = src[so];
            }
        }
        dst[i] = v;
    }
}

// depthwise: dst [R][S][C] <- src [C][R][S] (KRSC with cg==1 and OIHW with I==1 coincide).
__global__ void __launch_bounds__(256) prepack_dw_kernel(const int8_t* __restrict__ src, int8_t* __restrict__ dst,
                                                         int32_t c, int32_t rs)
{
    const int64_t total = (int64_t)c * rs;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int32_t ic = (int32_t)(i % c);
        const int32_t t = (int32_t)(i / c);
        dst[i] = src[(int64_t)ic * rs + t];
    }
}

inline unsigned grid_for(int64_t total)
{
    int64_t b = (total + 255) / 256;
    if (b < 1) b = 1;
    if (b > 148 * 32) b = 148 * 32;
    return (unsigned)b;
}

}  // namespace

lbc_status launch_prepack_krsc(const int8_t* src, int32_t src_layout, int8_t* dst, int32_t k, int32_t r, int32_t s,
                               int32_t cg, int32_t c_pad, cudaStream_t stream)
{
    const int64_t total = (int64_t)k * r * s * c_pad;
    prepack_krsc_kernel<<<grid_for(total), 256, 0, stream>>>(src, src_layout == LBC_W_OIHW, dst, k, r, s, cg, c_pad);
    LBC_CUDA_TRY(cudaGetLastError());
    return LBC_OK;
}

lbc_status launch_prepack_igemm(const int8_t* src, int32_t src_layout, int8_t* dst, int32_t k, int32_t r, int32_t s,
                                int32_t cg, int32_t s_pad, int32_t bkc, int32_t cblocks, int32_t chunk_outer,
                                cudaStream_t stream)
{
    const int64_t total = (int64_t)k * r * s_pad * cblocks * bkc;
    prepack_igemm_kernel<<<grid_for(total), 256, 0, stream>>>(src, src_layout == LBC_W_OIHW, dst, k, r, s, cg, s_pad,
                                                              bkc, cblocks, chunk_outer);
    LBC_CUDA_TRY(cudaGetLastError());
    return LBC_OK;
}

lbc_status launch_blockdiag(const int8_t* src, int8_t* dst, int32_t k, int32_t c, int32_t f, cudaStream_t stream)
{
    blockdiag_kernel<<<grid_for((int64_t)f * k * f * c), 256, 0, stream>>>(src, dst, k, c, f);
    LBC_CUDA_TRY(cudaGetLastError());
    return LBC_OK;
}

lbc_status launch_stem_xform(const int8_t* x, void* out, int32_t n, int32_t h, int32_t w, int32_t c, int32_t hs,
                             int32_t ws, int32_t sh, int32_t sw, int32_t pad_h, int32_t pad_w, cudaStream_t stream)
{
    StemXformParams p{n, h, w, c, hs, ws, sh, sw, pad_h, pad_w};
    LBC_REQUIRE(hs <= 65535 && n <= 65535, LBC_ERR_UNSUPPORTED, "small-C path: image height / batch beyond the grid limits");
    const dim3 grid((unsigned)((ws + 127) / 128), (unsigned)((hs + kXformRows - 1) / kXformRows), (unsigned)n);
    uint4* o = reinterpret_cast<uint4*>(out);
    if (c == 3 && sh == 2 && sw == 2) stem_xform_kernel<3, 2, 2><<<grid, 128, 0, stream>>>(p, x, o);
    else if (c == 3 && sh == 1 && sw == 1) stem_xform_kernel<3, 1, 1><<<grid, 128, 0, stream>>>(p, x, o);
    else if (c == 4 && sh == 2 && sw == 2) stem_xform_kernel<4, 2, 2><<<grid, 128, 0, stream>>>(p, x, o);
    else if (c == 1 && sh == 1 && sw == 1) stem_xform_kernel<1, 1, 1><<<grid, 128, 0, stream>>>(p, x, o);
    else stem_xform_kernel<0, 0, 0><<<grid, 128, 0, stream>>>(p, x, o);
    LBC_CUDA_TRY(cudaGetLastError());
    return LBC_OK;
}

lbc_status launch_prepack_stem(const int8_t* src, int32_t src_layout, int8_t* dst, int32_t k, int32_t r, int32_t s,
                               int32_t c, int32_t sh, int32_t sw, int32_t r2, int32_t s_pad, cudaStream_t stream)
{
    prepack_stem_kernel<<<grid_for((int64_t)k * r2 * s_pad * 16), 256, 0, stream>>>(src, src_layout == LBC_W_OIHW, dst, k,
                                                                                    r, s, c, sh, sw, r2, s_pad);
    LBC_CUDA_TRY(cudaGetLastError());
    return LBC_OK;
}

lbc_status launch_prepack_depthwise(const int8_t* src, int32_t /*src_layout*/, int8_t* dst, int32_t c, int32_t r,
                                    int32_t s, cudaStream_t stream)
{
    prepack_dw_kernel<<<grid_for((int64_t)c * r * s), 256, 0, stream>>>(src, dst, c, r * s);
    LBC_CUDA_TRY(cudaGetLastError());
    return LBC_OK;
}

// dst dim j takes source dim perm[j]; `dims` are the SOURCE extents.
// dst [C][R][S][K] <- src [K][R][S][C] rotated by 180 degrees in (R, S): the filter of the data-gradient convolution
__global__ void __launch_bounds__(256) dgrad_weights_kernel(const int8_t* __restrict__ src, int8_t* __restrict__ dst, int32_t k,
                                                            int32_t r, int32_t s, int32_t c)
{
    const int64_t total = (int64_t)k * r * s * c;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int32_t ik = (int32_t)(i % k);
        int64_t t = i / k;
        const int32_t is = (int32_t)(t % s); t /= s;
        const int32_t ir = (int32_t)(t % r); t /= r;
        const int32_t ic = (int32_t)t;
        dst[i] = src[(((int64_t)ik * r + (r - 1 - ir)) * s + (s - 1 - is)) * c + ic];
    }
}

lbc_status launch_dgrad_weights(const int8_t* src, int8_t* dst, int32_t k, int32_t r, int32_t s, int32_t c, cudaStream_t stream)
{
    dgrad_weights_kernel<<<grid_for((int64_t)k * r * s * c), 256, 0, stream>>>(src, dst, k, r, s, c);
    LBC_CUDA_TRY(cudaGetLastError());
    return LBC_OK;
}

lbc_status launch_permute5(const void* src, void* dst, const int32_t dims[5], const int32_t perm[5], int32_t elt,
                           cudaStream_t stream)
{
    LBC_REQUIRE(src && dst, LBC_ERR_INVALID_ARG, "permute: null buffer");
    LBC_REQUIRE(elt == 1 || elt == 4, LBC_ERR_INVALID_ARG, "permute: element size must be 1 or 4 bytes");
    int64_t sstride_src[5];
    int64_t acc = 1;
    for (int j = 4; j >= 0; --j) {
        LBC_REQUIRE(dims[j] > 0, LBC_ERR_INVALID_ARG, "permute: non-positive extent");
        sstride_src[j] = acc;
        acc *= dims[j];
    }
    Permute5 pm;
    for (int j = 0; j < 5; ++j) {
        pm.ddim[j] = dims[perm[j]];
        pm.sstride[j] = sstride_src[perm[j]];
    }
    if (elt == 1)
        permute5_kernel<int8_t><<<grid_for(acc), 256, 0, stream>>>(pm, (const int8_t*)src, (int8_t*)dst, acc);
    else
        permute5_kernel<int32_t><<<grid_for(acc), 256, 0, stream>>>(pm, (const int32_t*)src, (int32_t*)dst, acc);
    LBC_CUDA_TRY(cudaGetLastError());
    return LBC_OK;
}

}  // namespace lbc
