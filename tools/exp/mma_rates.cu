// mma_rates.cu — cycles per tcgen05.mma.kind::i8 (M=128, K=32) as a function of N, operand row width (swizzle
// mode) and the A descriptor's starting row (the shifted-window mode starts A on arbitrary pixel rows).
// No loads, no epilogue: operands are whatever is in shared memory.  Build: see tools/exp/README or
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I lowbitdnn-project_b200/csrc -o tools/bin/mma_rates tools/exp/mma_rates.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "ptx.cuh"
using namespace lbc;

__device__ int g_flag = 0;

// row_bytes: 128/64/32 swizzled K-major rows; 16 = unswizzled 16-byte pixels (LBO = 16, SBO = 128)
__global__ void __launch_bounds__(128, 1) rate_kernel(int n, int row_bytes, int a_row_shift, int ksteps, int iters, long long* cyc)
{
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* a = smem;                // up to (128 + 64 shift rows) x 128 B = 24 KB
    uint8_t* b = smem + 32 * 1024;    // 256 x 128 B = 32 KB
    __shared__ uint64_t done_bar;
    __shared__ uint32_t tmem_base_s;
    for (int i = threadIdx.x; i < 64 * 1024 / 16; i += blockDim.x) reinterpret_cast<int4*>(smem)[i] = make_int4(i, i * 3, i * 5, i * 7);
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) { ptx::mbar_init(&done_bar, 1); ptx::fence_barrier_init(); }
    if (warp == 1) { ptx::tmem_alloc(&tmem_base_s, 256); ptx::tmem_relinquish(); }
    ptx::fence_proxy_async();
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_d = tmem_base_s;
    if (warp == 0 && lane == 0) {
        const uint32_t idesc = ptx::make_idesc_i8(128, (uint32_t)n);
        uint64_t da, db;
        if (row_bytes == 16) {
            da = ptx::make_kmajor_desc_nosw(ptx::smem_u32(a) + a_row_shift * 16, 16, 128);
            db = ptx::make_kmajor_desc_nosw(ptx::smem_u32(b), 16, 128);
        } else {
            da = ptx::make_kmajor_desc(ptx::smem_u32(a), row_bytes) + (uint64_t)((a_row_shift * row_bytes) >> 4);
            db = ptx::make_kmajor_desc(ptx::smem_u32(b), row_bytes);
        }
        const long long t0 = clock64();
        // k-steps inside one row (2 x 16 B per step); single-step rows re-issue the same operands
        const uint64_t kmask = (row_bytes >= 64) ? 1ull : 0ull;
        ptx::mma_i8_ss(tmem_d, da, db, idesc, 0u);
        for (int32_t it = 0; it < iters; ++it) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
                ptx::mma_i8_ss(tmem_d, da + 2ull * ((uint64_t)k & kmask), db + 2ull * ((uint64_t)k & kmask), idesc, 1u);
        }
        ptx::mma_commit(&done_bar);
        ptx::mbar_wait(&done_bar, 0, &g_flag, 4000000000ull);
        const long long t1 = clock64();
        cyc[blockIdx.x] = t1 - t0;
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) { ptx::tc_fence_after(); ptx::tmem_dealloc(tmem_d, 256); }
}

int main()
{
    long long* cyc;
    cudaMalloc(&cyc, 148 * 8);
    const size_t smem = 65 * 1024 + 1024;
    cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int iters = 2000, ksteps = 4;
    printf("%5s %9s %9s | cycles per MMA (128 x N x 32)   ideal = N/2\n", "N", "row B", "A shift");
    for (int n : {32, 64, 96, 128, 192, 256})
        for (int rb : {128, 64, 32, 16})
            for (int sh : {0, 1, 9, 58}) {
                if (rb == 16 && sh == 9) continue;
                rate_kernel<<<148, 128, smem>>>(n, rb, sh, ksteps, iters, cyc);
                cudaError_t e = cudaDeviceSynchronize();
                long long h[148];
                cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
                double c = 0;
                for (int i = 0; i < 148; ++i) c += h[i];
                c /= 148;
                printf("%5d %9d %9d | %7.1f   (%s)\n", n, rb, sh, c / ((double)iters * ksteps), cudaGetErrorString(e));
            }
    return 0;
}
