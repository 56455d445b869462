// pool_add.cu — the int8 ops that sit BETWEEN convolutions in the reference's chains: max-pool, residual
// add (+ ReLU) and global average pool, all on NHWC int8 so that a whole network stays int8 on the device.
//
//   max-pool      replaces max_pool2d(input, kernel, stride, padding) of python/qtorch/cpp/pool2d.cuh:54-92 (cuDNN
//                 CUDNN_POOLING_MAX_DETERMINISTIC on int8 NCHW_VECT_C): out = 1 + (in + 2*pad - window) / stride, padding
//                 elements never win (-inf), plain integer max.  Used as conv -> pool -> relu in python/tmp.py:43-56.
//   add (+ReLU)   y = sat_int8(a + b), optional clamp at 0: the residual join of a bottleneck.  The reference has no int8
//                 residual add (it never built a ResNet); the rule is the saturating-add convention of its quantizer
//                 (round, then saturate to [-128, 127]: conv2DForward3x3WinogradFused.cuh:39-46) applied to an exact sum.
//   global pool   y = requant(sum over H*W) with one fp32 scale: same epilogue rule as the convolutions (requant_s32).
//
// All three are pure HBM streams (1-2 bytes of traffic per output byte, a handful of byte-wise integer ops): 16-byte
// vectors per thread, coalesced along C, grids sized in whole multiples of the SM count.  CUDA cores only.
#include "common.cuh"

#include <algorithm>

namespace lbc {

namespace {

__device__ __forceinline__ uint32_t vmax4(uint32_t a, uint32_t b) { return __vmaxs4(a, b); }

// One thread: 16 channels (one uint4) of one output pixel.
__global__ void __launch_bounds__(256) maxpool_v16_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, int32_t n, int32_t h,
                                                          int32_t w, int32_t c16, int32_t p, int32_t q, int32_t kh, int32_t kw,
                                                          int32_t sh, int32_t sw, int32_t ph, int32_t pw, int64_t total)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t r = i;
        const int32_t cv = (int32_t)(r % c16); r /= c16;
        const int32_t qq = (int32_t)(r % q); r /= q;
        const int32_t pp = (int32_t)(r % p);
        const int32_t img = (int32_t)(r / p);
        const int32_t h0 = pp * sh - ph, w0 = qq * sw - pw;
        uint4 m = make_uint4(0x80808080u, 0x80808080u, 0x80808080u, 0x80808080u);      // -128: loses against everything
        for (int32_t a = 0; a < kh; ++a) {
            const int32_t hh = h0 + a;
            if (hh < 0 || hh >= h) continue;
            for (int32_t b = 0; b < kw; ++b) {
                const int32_t ww = w0 + b;
                if (ww < 0 || ww >= w) continue;
                const uint4 v = __ldg(x + (((int64_t)img * h + hh) * w + ww) * c16 + cv);
                m.x = vmax4(m.x, v.x); m.y = vmax4(m.y, v.y); m.z = vmax4(m.z, v.z); m.w = vmax4(m.w, v.w);
            }
        }
        y[i] = m;
    }
}

// any channel count: one thread per output element
__global__ void __launch_bounds__(256) maxpool_scalar_kernel(const int8_t* __restrict__ x, int8_t* __restrict__ y, int32_t n, int32_t h,
                                                             int32_t w, int32_t c, int32_t p, int32_t q, int32_t kh, int32_t kw,
                                                             int32_t sh, int32_t sw, int32_t ph, int32_t pw, int64_t total)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t r = i;
        const int32_t ch = (int32_t)(r % c); r /= c;
        const int32_t qq = (int32_t)(r % q); r /= q;
        const int32_t pp = (int32_t)(r % p);
        const int32_t img = (int32_t)(r / p);
        int32_t m = -128;
        for (int32_t a = 0; a < kh; ++a) {
            const int32_t hh = pp * sh - ph + a;
            if (hh < 0 || hh >= h) continue;
            for (int32_t b = 0; b < kw; ++b) {
                const int32_t ww = qq * sw - pw + b;
                if (ww < 0 || ww >= w) continue;
                m = max(m, (int32_t)x[(((int64_t)img * h + hh) * w + ww) * c + ch]);
            }
        }
        y[i] = (int8_t)m;
    }
}

// y = sat(a + b), optional ReLU; 16 bytes per thread per step, scalar tail
__global__ void __launch_bounds__(256) add_relu_kernel(const int8_t* __restrict__ a, const int8_t* __restrict__ b, int8_t* __restrict__ y,
                                                       size_t n, int32_t relu, int32_t vec_ok)
{
    const size_t nv = vec_ok ? n / 16 : 0;
    const uint4* a4 = reinterpret_cast<const uint4*>(a);
    const uint4* b4 = reinterpret_cast<const uint4*>(b);
    uint4* y4 = reinterpret_cast<uint4*>(y);
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += stride) {
        const uint4 u = __ldg(a4 + i), v = __ldg(b4 + i);
        uint4 r = make_uint4(__vaddss4(u.x, v.x), __vaddss4(u.y, v.y), __vaddss4(u.z, v.z), __vaddss4(u.w, v.w));
        if (relu) { r.x = __vmaxs4(r.x, 0u); r.y = __vmaxs4(r.y, 0u); r.z = __vmaxs4(r.z, 0u); r.w = __vmaxs4(r.w, 0u); }
        y4[i] = r;
    }
    for (size_t i = nv * 16 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        int32_t s = (int32_t)a[i] + (int32_t)b[i];
        s = min(127, max(relu ? 0 : -128, s));
        y[i] = (int8_t)s;
    }
}

// one block per (image, group of 64 channels): threads x = channel quad (16), y = pixel lanes (16); int32 sums, then the
// convolution epilogue's requantisation with a single scale
__global__ void __launch_bounds__(256) global_avgpool_kernel(const int8_t* __restrict__ x, int8_t* __restrict__ y, int32_t hw, int32_t c,
                                                             float scale)
{
    __shared__ int32_t part[16][65];
    const int32_t img = blockIdx.y, c0 = blockIdx.x * 64 + threadIdx.x * 4;
    int32_t s[4] = {0, 0, 0, 0};
    if (c0 < c) {
        for (int32_t px = threadIdx.y; px < hw; px += 16) {
            const int8_t* src = x + ((int64_t)img * hw + px) * c + c0;
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (c0 + j < c) s[j] += (int32_t)src[j];
        }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) part[threadIdx.y][threadIdx.x * 4 + j] = s[j];
    __syncthreads();
    const int32_t t = threadIdx.y * 16 + threadIdx.x;
    if (t < 64 && blockIdx.x * 64 + t < c) {
        int32_t tot = 0;
#pragma unroll
        for (int r = 0; r < 16; ++r) tot += part[r][t];
        y[(int64_t)img * c + blockIdx.x * 64 + t] = requant_s8(tot, 0, scale, -128);
    }
}

int grid_for(int64_t work_items, int sm_count)
{
    const int64_t blocks = (work_items + 255) / 256;
    const int64_t cap = (int64_t)(sm_count > 0 ? sm_count : 148) * 16;      // whole multiples of the SM count
    if (blocks >= cap) return (int)cap;
    return (int)std::max<int64_t>(1, blocks);
}

}  // namespace

lbc_status launch_maxpool(const lbc_pool_desc& d, int32_t p, int32_t q, const int8_t* x, int8_t* y, int sm_count, cudaStream_t stream)
{
    const bool v16 = d.c % 16 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(y) & 15) == 0;
    if (v16) {
        const int64_t total = (int64_t)d.n * p * q * (d.c / 16);
        maxpool_v16_kernel<<<grid_for(total, sm_count), 256, 0, stream>>>(reinterpret_cast<const uint4*>(x), reinterpret_cast<uint4*>(y),
                                                                           d.n, d.h, d.w, d.c / 16, p, q, d.kh, d.kw, d.stride_h, d.stride_w,
                                                                           d.pad_h, d.pad_w, total);
    } else {
        const int64_t total = (int64_t)d.n * p * q * d.c;
        maxpool_scalar_kernel<<<grid_for(total, sm_count), 256, 0, stream>>>(x, y, d.n, d.h, d.w, d.c, p, q, d.kh, d.kw, d.stride_h,
                                                                              d.stride_w, d.pad_h, d.pad_w, total);
    }
    LBC_CUDA_TRY(cudaGetLastError());
    return LBC_OK;
}

lbc_status launch_add_relu(const int8_t* a, const int8_t* b, int8_t* y, size_t n, int32_t relu, int sm_count, cudaStream_t stream)
{
    const int vec_ok = ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(y)) & 15) == 0;
    add_relu_kernel<<<grid_for((int64_t)((n + 15) / 16), sm_count), 256, 0, stream>>>(a, b, y, n, relu, vec_ok);
    LBC_CUDA_TRY(cudaGetLastError());
    return LBC_OK;
}

lbc_status launch_global_avgpool(const int8_t* x, int8_t* y, int32_t n, int32_t hw, int32_t c, float scale, cudaStream_t stream)
{
    const dim3 grid((unsigned)((c + 63) / 64), (unsigned)n), block(16, 16);
    global_avgpool_kernel<<<grid, block, 0, stream>>>(x, y, hw, c, scale);
    LBC_CUDA_TRY(cudaGetLastError());
    return LBC_OK;
}

}  // namespace lbc
