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

lbc_status launch_prepack_depthwise(const int8_t* src, int32_t /*src_layout*/, int8_t* dst, int32_t c, int32_t r,
                                    int32_t s, cudaStream_t stream)
{
    prepack_dw_kernel<<<grid_for((int64_t)c * r * s), 256, 0, stream>>>(src, dst, c, r * s);
    LBC_CUDA_TRY(cudaGetLastError());
    return LBC_OK;
}

// dst dim j takes source dim perm[j]; `dims` are the SOURCE extents.
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
