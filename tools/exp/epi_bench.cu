// epi_bench.cu — which requantise instruction sequence is fastest on sm_100a?  Each thread converts 64 int32
// accumulators per iteration (bias/scale from shared memory, like the real epilogue) and packs to int8.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t pack4(uint32_t a, uint32_t b, uint32_t c, uint32_t d)
{
    return __byte_perm(__byte_perm(a, b, 0x0040), __byte_perm(c, d, 0x0040), 0x5410);
}

template <int V>
__device__ __forceinline__ uint32_t q4(const int* acc, const int4 b, const float4 s, float lo)
{
    if (V == 0) {   // current: float clamp + magic add
        auto f = [&](int a, int bb, float sc) {
            float x = __fmul_rn(__int2float_rn(a + bb), sc);
            x = fminf(fmaxf(x, lo), 127.f);
            return __float_as_uint(__fadd_rn(x, 12582912.f));
        };
        return pack4(f(acc[0], b.x, s.x), f(acc[1], b.y, s.y), f(acc[2], b.z, s.z), f(acc[3], b.w, s.w));
    } else if (V == 1) {   // cvt.rni.sat.s8.f32 (+ relu max in float)
        auto f = [&](int a, int bb, float sc) {
            float x = fmaxf(__fmul_rn(__int2float_rn(a + bb), sc), lo);
            int r;
            asm("cvt.rni.sat.s8.f32 %0, %1;" : "=r"(r) : "f"(x));
            return (uint32_t)r;
        };
        return pack4(f(acc[0], b.x, s.x), f(acc[1], b.y, s.y), f(acc[2], b.z, s.z), f(acc[3], b.w, s.w));
    } else if (V == 2) {   // F2I rni (s32) + cvt.pack.sat.s8.s32 (I2IP)
        auto f = [&](int a, int bb, float sc) {
            float x = fmaxf(__fmul_rn(__int2float_rn(a + bb), sc), lo);
            return __float2int_rn(x);
        };
        int q0 = f(acc[0], b.x, s.x), q1 = f(acc[1], b.y, s.y), q2 = f(acc[2], b.z, s.z), q3 = f(acc[3], b.w, s.w);
        uint32_t hi, r;
        asm("cvt.pack.sat.s8.s32.b32 %0, %1, %2, 0;" : "=r"(hi) : "r"(q3), "r"(q2));
        asm("cvt.pack.sat.s8.s32.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(q1), "r"(q0), "r"(hi));
        return r;
    } else if (V == 4) {   // V2 without the float clamp (non-ReLU, NaN ignored)
        auto f = [&](int a, int bb, float sc) { return __float2int_rn(__fmul_rn(__int2float_rn(a + bb), sc)); };
        int q0 = f(acc[0], b.x, s.x), q1 = f(acc[1], b.y, s.y), q2 = f(acc[2], b.z, s.z), q3 = f(acc[3], b.w, s.w);
        uint32_t hi, r;
        asm("cvt.pack.sat.s8.s32.b32 %0, %1, %2, 0;" : "=r"(hi) : "r"(q3), "r"(q2));
        asm("cvt.pack.sat.s8.s32.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(q1), "r"(q0), "r"(hi));
        return r;
    } else if (V == 5) {   // ReLU on the packed word: pack.sat.s8 then signed-byte max with 0 = clear bytes whose sign bit is set
        auto f = [&](int a, int bb, float sc) { return __float2int_rn(__fmul_rn(__int2float_rn(a + bb), sc)); };
        int q0 = f(acc[0], b.x, s.x), q1 = f(acc[1], b.y, s.y), q2 = f(acc[2], b.z, s.z), q3 = f(acc[3], b.w, s.w);
        uint32_t hi, r;
        asm("cvt.pack.sat.s8.s32.b32 %0, %1, %2, 0;" : "=r"(hi) : "r"(q3), "r"(q2));
        asm("cvt.pack.sat.s8.s32.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(q1), "r"(q0), "r"(hi));
        const uint32_t neg = (r >> 7) & 0x01010101u;          // 1 per negative byte
        return r & ~(neg * 0xffu);                             // IMAD + LOP3
    } else if (V == 6) {   // V4 with scale/bias in registers (no LDS) - lower bound
        auto f = [&](int a, int bb, float sc) { return __float2int_rn(__fmul_rn(__int2float_rn(a + bb), sc)); };
        int q0 = f(acc[0], lo > 1.f ? b.x : 77, lo > 1.f ? s.x : 0.013f), q1 = f(acc[1], lo > 1.f ? b.y : 78, lo > 1.f ? s.y : 0.014f),
            q2 = f(acc[2], lo > 1.f ? b.z : 79, lo > 1.f ? s.z : 0.015f), q3 = f(acc[3], lo > 1.f ? b.w : 80, lo > 1.f ? s.w : 0.016f);
        uint32_t hi, r;
        asm("cvt.pack.sat.s8.s32.b32 %0, %1, %2, 0;" : "=r"(hi) : "r"(q3), "r"(q2));
        asm("cvt.pack.sat.s8.s32.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(q1), "r"(q0), "r"(hi));
        return r;
    } else {   // V == 3: float clamp + magic add, pack via I2IP-free shifts: (a&255)|(b&255)<<8 ... using LOP3/prmt alt
        auto f = [&](int a, int bb, float sc) {
            float x = __fmul_rn(__int2float_rn(a + bb), sc);
            x = fminf(fmaxf(x, lo), 127.f);
            return __float_as_uint(__fadd_rn(x, 12582912.f));
        };
        uint32_t q0 = f(acc[0], b.x, s.x), q1 = f(acc[1], b.y, s.y), q2 = f(acc[2], b.z, s.z), q3 = f(acc[3], b.w, s.w);
        uint32_t hi, r;
        asm("cvt.pack.sat.s8.s32.b32 %0, %1, %2, 0;" : "=r"(hi) : "r"((int)(q3 << 24) >> 24), "r"((int)(q2 << 24) >> 24));
        asm("cvt.pack.sat.s8.s32.b32 %0, %1, %2, %3;" : "=r"(r) : "r"((int)(q1 << 24) >> 24), "r"((int)(q0 << 24) >> 24), "r"(hi));
        return r;
    }
}

template <int V>
__global__ void __launch_bounds__(512, 1) bench(const int* in, uint32_t* out, int iters, float lo, long long* cyc)
{
    __shared__ __align__(16) float sc[256];
    __shared__ __align__(16) int bi[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) { sc[i] = 0.01f + i * 1e-4f; bi[i] = i * 7 - 900; }
    __syncthreads();
    int acc[64];
    for (int j = 0; j < 64; ++j) acc[j] = in[(threadIdx.x * 64 + j) & 1023];
    uint32_t x = 0;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int g = 0; g < 16; ++g) {
            const int c = ((it & 3) * 64 + g * 4);
            x ^= q4<V>(acc + g * 4, *reinterpret_cast<const int4*>(bi + c), *reinterpret_cast<const float4*>(sc + c), lo);
        }
#pragma unroll
        for (int j = 0; j < 64; ++j) acc[j] += x & 3;
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = x;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int V>
void run(const char* name, int threads)
{
    int* in; uint32_t* out; long long* cyc;
    cudaMalloc(&in, 4096); cudaMemset(in, 3, 4096);
    cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&cyc, 148 * 8);
    const int iters = 2000;
    bench<V><<<148, threads>>>(in, out, iters, 0.f, cyc);
    cudaDeviceSynchronize();
    bench<V><<<148, threads>>>(in, out, iters, 0.f, cyc);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
    double c = 0; for (int i = 0; i < 148; ++i) c += h[i]; c /= 148;
    double elts = (double)iters * 64 * threads;
    printf("%-28s threads %3d: %.3f cycles/element/SM  -> 128x256 tile = %.0f cycles  (%s)\n", name, threads, c / elts,
           c / elts * 32768, cudaGetErrorString(e));
}

int main()
{
    for (int th : {256, 512}) {
        run<0>("V0 fclamp+magic+prmt", th);
        run<1>("V1 cvt.rni.sat.s8.f32+prmt", th);
        run<2>("V2 f2i + cvt.pack.sat", th);
        run<3>("V3 fclamp+magic+cvt.pack", th);
        run<4>("V4 V2 minus fmax", th);
        run<5>("V5 relu on packed word", th);
    }
    return 0;
}
