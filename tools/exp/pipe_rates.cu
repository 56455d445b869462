// pipe_rates.cu — per-SM throughput of the instructions the requantise epilogue is made of (sm_100a).
// 148 CTAs x 512 threads; every thread runs 8 independent dependency chains of one instruction kind.
// Prints lanes/clk/SM (128 = one warp instruction per SMSP per cycle).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

enum { K_IADD, K_I2F, K_F2I, K_FMUL, K_FMNMX, K_I2IP, K_PRMT, K_FADD, K_LDS128, K_FFMA, K_IMNMX, K_I2F_F2I, K_VMINU4, K_COUNT };
static const char* names[] = {"IADD", "I2F.S32", "F2I.RNI", "FMUL", "FMNMX", "I2IP(cvt.pack.sat)", "PRMT", "FADD", "LDS.128",
                              "FFMA", "IMNMX", "I2F+F2I mix", "VMINU4"};

template <int K>
__device__ __forceinline__ uint32_t op(uint32_t x, uint32_t y, const uint4* sm)
{
    uint32_t r = x;
    if (K == K_IADD) asm volatile("add.s32 %0, %1, %2;" : "=r"(r) : "r"(x), "r"(y));
    if (K == K_I2F) asm volatile("cvt.rn.f32.s32 %0, %1;" : "=r"(r) : "r"(x));
    if (K == K_F2I) asm volatile("cvt.rni.s32.f32 %0, %1;" : "=r"(r) : "r"(x));
    if (K == K_FMUL) asm volatile("mul.rn.f32 %0, %1, %2;" : "=r"(r) : "r"(x), "r"(y));
    if (K == K_FMNMX) asm volatile("max.f32 %0, %1, %2;" : "=r"(r) : "r"(x), "r"(y));
    if (K == K_I2IP) asm volatile("cvt.pack.sat.s8.s32.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(x), "r"(y), "r"(x));
    if (K == K_PRMT) asm volatile("prmt.b32 %0, %1, %2, 0x0040;" : "=r"(r) : "r"(x), "r"(y));
    if (K == K_FADD) asm volatile("add.rn.f32 %0, %1, %2;" : "=r"(r) : "r"(x), "r"(y));
    if (K == K_LDS128) { uint32_t a0, a1, a2, a3; asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a0), "=r"(a1), "=r"(a2), "=r"(a3) : "r"((uint32_t)__cvta_generic_to_shared(sm) + ((x & 31) << 4))); r = a0; }
    if (K == K_FFMA) asm volatile("fma.rn.f32 %0, %1, %2, %1;" : "=r"(r) : "r"(x), "r"(y));
    if (K == K_IMNMX) asm volatile("max.s32 %0, %1, %2;" : "=r"(r) : "r"(x), "r"(y));
    if (K == K_I2F_F2I) { asm volatile("cvt.rn.f32.s32 %0, %1;" : "=r"(r) : "r"(x)); asm volatile("cvt.rni.s32.f32 %0, %1;" : "=r"(r) : "r"(r)); }
    if (K == K_VMINU4) asm volatile("vmin4.u32.u32.u32 %0, %1, %2, %0;" : "+r"(r) : "r"(x), "r"(y));
    return r;
}

template <int K>
__global__ void __launch_bounds__(512, 1) bench(uint32_t* out, int iters, uint32_t seed, long long* cyc)
{
    __shared__ uint4 sm[64];
    if (threadIdx.x < 64) sm[threadIdx.x] = make_uint4(threadIdx.x, seed, 3, 4);
    __syncthreads();
    uint32_t a[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] = seed * (j + 1) + threadIdx.x;
    const uint32_t y = seed | 1;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int j = 0; j < 8; ++j) a[j] = op<K>(a[j], y, sm);
        }
    }
    long long t1 = clock64();
    uint32_t x = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) x ^= a[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = x;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int K>
void run()
{
    uint32_t* out; long long* cyc;
    cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&cyc, 148 * 8);
    const int iters = 4000;
    bench<K><<<148, 512>>>(out, iters, 12345u, cyc);
    cudaDeviceSynchronize();
    bench<K><<<148, 512>>>(out, iters, 12345u, cyc);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
    double c = 0; for (int i = 0; i < 148; ++i) c += h[i]; c /= 148;
    const double n = (double)iters * 64 * 512 * (K == K_I2F_F2I ? 2 : 1);
    printf("%-20s %7.1f lanes/clk/SM   (%.2f cyc per warp-instr per SMSP)  %s\n", names[K], n / c, c / (n / 128), cudaGetErrorString(e));
    cudaFree(out); cudaFree(cyc);
}

int main()
{
    run<K_IADD>(); run<K_I2F>(); run<K_F2I>(); run<K_FMUL>(); run<K_FMNMX>(); run<K_I2IP>(); run<K_PRMT>(); run<K_FADD>();
    run<K_LDS128>(); run<K_FFMA>(); run<K_IMNMX>(); run<K_I2F_F2I>(); run<K_VMINU4>();
    return 0;
}
