// direct_conv.cu — CUDA-core direct convolution (any R,S,stride,pad,dilation,groups): the always-available
// correct path for every descriptor (grouped convolutions, channel counts the tensor-core path cannot tile).
// The depthwise kernel lives in depthwise.cu.
//
// Replaces, functionally: CUDAConv2DForward3x3CudaV1 (cpp/int8conv/conv2DForward3x3.cuh:602-676), which is
// 3x3/stride-1 only and atomically accumulates partial sums; here each thread owns its outputs.
#include "common.cuh"

namespace lbc {

namespace {

struct DirectParams {
    int32_t n, h, w, c, k, r, s, stride_h, stride_w, pad_h, pad_w, dil_h, dil_w;
    int32_t p, q, cg, kg;
    int64_t m_total;
    int32_t kquads;   // ceil(K / KT)
    int32_t relu, out_mode;
};

// One thread = one output pixel x KT consecutive output channels (all in one group when KT == 4).
template <int KT, bool VEC4>
__global__ void __launch_bounds__(256) direct_conv_kernel(DirectParams g, const int8_t* __restrict__ x,
                                                          const int8_t* __restrict__ wgt,
                                                          const int32_t* __restrict__ bias,
                                                          const float* __restrict__ scale, void* __restrict__ y)
{
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= g.m_total * g.kquads) return;
    const int32_t kq = (int32_t)(idx % g.kquads);
    const int64_t m = idx / g.kquads;
    const int32_t k0 = kq * KT;
    const int32_t q = (int32_t)(m % g.q);
    const int32_t p = (int32_t)((m / g.q) % g.p);
    const int32_t n = (int32_t)(m / ((int64_t)g.q * g.p));
    const int32_t grp = k0 / g.kg;

    int32_t acc[KT];
#pragma unroll
    for (int j = 0; j < KT; ++j) acc[j] = 0;

    const int64_t wk_stride = (int64_t)g.r * g.s * g.cg;
    for (int32_t r = 0; r < g.r; ++r) {
        const int32_t ih = p * g.stride_h - g.pad_h + r * g.dil_h;
        if (ih < 0 || ih >= g.h) continue;   // zero padding
        for (int32_t s = 0; s < g.s; ++s) {
            const int32_t iw = q * g.stride_w - g.pad_w + s * g.dil_w;
            if (iw < 0 || iw >= g.w) continue;
            const int8_t* xp = x + (((int64_t)n * g.h + ih) * g.w + iw) * g.c + (int64_t)grp * g.cg;
            const int8_t* wp = wgt + ((int64_t)k0 * g.r + r) * g.s * g.cg + (int64_t)s * g.cg;
            if (VEC4) {
                for (int32_t c = 0; c < g.cg; c += 4) {
                    const int32_t xv = *reinterpret_cast<const int32_t*>(xp + c);
#pragma unroll
                    for (int j = 0; j < KT; ++j) {
                        if (k0 + j < g.k) {
                            const int32_t wv = __ldg(reinterpret_cast<const int32_t*>(wp + j * wk_stride + c));
                            acc[j] = __dp4a(xv, wv, acc[j]);
                        }
                    }
                }
            } else {
                for (int32_t c = 0; c < g.cg; ++c) {
                    const int32_t xv = xp[c];
#pragma unroll
                    for (int j = 0; j < KT; ++j)
                        if (k0 + j < g.k) acc[j] += xv * (int32_t)__ldg(wp + j * wk_stride + c);
                }
            }
        }
    }

    const int64_t o = m * g.k + k0;
    const int32_t lo = g.relu ? 0 : -128;
    if (g.out_mode == LBC_OUT_INT32) {
        int32_t* yo = reinterpret_cast<int32_t*>(y) + o;
#pragma unroll
        for (int j = 0; j < KT; ++j)
            if (k0 + j < g.k) yo[j] = acc[j] + (bias ? bias[k0 + j] : 0);
    } else {
        int8_t* yo = reinterpret_cast<int8_t*>(y) + o;
        if (KT == 4 && (g.k & 3) == 0) {
            int32_t b[4];
#pragma unroll
            for (int j = 0; j < 4; ++j)
                b[j] = requant_s32(acc[j], bias ? bias[k0 + j] : 0, scale[k0 + j], lo);
            *reinterpret_cast<uint32_t*>(yo) = pack4_sat_s8(b[0], b[1], b[2], b[3]);
        } else {
#pragma unroll
            for (int j = 0; j < KT; ++j)
                if (k0 + j < g.k)
                    yo[j] = requant_s8(acc[j], bias ? bias[k0 + j] : 0, scale[k0 + j], lo);
        }
    }
}

DirectParams to_params(const ConvGeom& g, int kt, const EpilogueParams& ep)
{
    DirectParams p{};
    const lbc_conv_desc& d = g.d;
    p.n = d.n; p.h = d.h; p.w = d.w; p.c = d.c; p.k = d.k; p.r = d.r; p.s = d.s;
    p.stride_h = d.stride_h; p.stride_w = d.stride_w; p.pad_h = d.pad_h; p.pad_w = d.pad_w;
    p.dil_h = d.dil_h; p.dil_w = d.dil_w;
    p.p = g.p; p.q = g.q; p.cg = g.cg; p.kg = g.kg; p.m_total = g.m_total;
    p.kquads = (d.k + kt - 1) / kt;
    p.relu = ep.relu; p.out_mode = ep.out_mode;
    return p;
}

}  // namespace

lbc_status launch_direct_conv(const ConvGeom& g, const int8_t* x, const int8_t* w, const EpilogueParams& ep, void* y,
                              cudaStream_t stream)
{
    const bool kt4 = (g.kg % 4 == 0);
    const bool vec4 = (g.cg % 4 == 0) && (g.d.c % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) & 3) == 0) &&
                      ((reinterpret_cast<uintptr_t>(w) & 3) == 0);
    const int kt = kt4 ? 4 : 1;
    DirectParams p = to_params(g, kt, ep);
    const int64_t threads = g.m_total * p.kquads;
    const int block = 256;
    const int64_t grid = (threads + block - 1) / block;
    LBC_REQUIRE(grid <= 0x7fffffffLL, LBC_ERR_UNSUPPORTED, "direct conv: grid too large (%lld blocks)", (long long)grid);
    if (kt4 && vec4)
        direct_conv_kernel<4, true><<<(unsigned)grid, block, 0, stream>>>(p, x, w, ep.bias, ep.scale, y);
    else if (kt4)
        direct_conv_kernel<4, false><<<(unsigned)grid, block, 0, stream>>>(p, x, w, ep.bias, ep.scale, y);
    else if (vec4)
        direct_conv_kernel<1, true><<<(unsigned)grid, block, 0, stream>>>(p, x, w, ep.bias, ep.scale, y);
    else
        direct_conv_kernel<1, false><<<(unsigned)grid, block, 0, stream>>>(p, x, w, ep.bias, ep.scale, y);
    LBC_CUDA_TRY(cudaGetLastError());
    return LBC_OK;
}

}  // namespace lbc
