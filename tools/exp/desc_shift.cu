// desc_shift.cu — experiment: can a K-major swizzled UMMA A-operand descriptor start at an arbitrary ROW of a
// TMA-layout tile (row-shifted window), and what must the descriptor's base_offset field be?
// A tile: ROWS rows x rowbytes (32/64/128), stored with the TMA swizzle (16B chunk index XOR (addr>>7)&mask).
// A[i][k] = (i*7 + k*3) % 251 - 125 ; B = [16 rows][32 B] with B[n][k] = (k == n) -> D[i][n] = A[i+shift][n].
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../lowbitdnn-project_b200/csrc/ptx.cuh"
using namespace lbc;

constexpr int ROWS = 192;

__device__ __host__ inline int aval(int i, int k) { return (i * 7 + k * 3) % 251 - 125; }

__global__ void __launch_bounds__(128, 1) k(int rowbytes, int shift, int base_off_mode, int* out /*128x16*/)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* a = smem;                 // ROWS*rowbytes  (<= 24 KB)
    uint8_t* b = smem + 32768;         // 16 rows x rowbytes
    __shared__ uint64_t bar;
    __shared__ uint32_t tbase;
    const int bits = rowbytes == 128 ? 3 : rowbytes == 64 ? 2 : 1;
    const uint32_t mask = (1u << bits) - 1;
    for (int idx = threadIdx.x; idx < ROWS * rowbytes; idx += blockDim.x) {
        int i = idx / rowbytes, kk = idx % rowbytes;
        uint32_t off = i * rowbytes + kk;
        uint32_t phys = off ^ (((off >> 7) & mask) << 4);
        a[phys] = (uint8_t)(int8_t)aval(i, kk);
    }
    for (int idx = threadIdx.x; idx < 16 * rowbytes; idx += blockDim.x) {
        int n = idx / rowbytes, kk = idx % rowbytes;
        uint32_t off = n * rowbytes + kk;
        uint32_t phys = off ^ (((off >> 7) & mask) << 4);
        b[phys] = (kk == n) ? 1 : 0;
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { ptx::mbar_init(&bar, 1); ptx::fence_barrier_init(); }
    if (warp == 1) { ptx::tmem_alloc(&tbase, 32); ptx::tmem_relinquish(); }
    ptx::fence_proxy_async();
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t td = tbase;
    __shared__ int tflag;
    if (threadIdx.x == 0) {
        tflag = 0;
        uint32_t a_addr = ptx::smem_u32(a) + shift * rowbytes;
        uint64_t da = ptx::make_kmajor_desc(a_addr, rowbytes);
        uint64_t db = ptx::make_kmajor_desc(ptx::smem_u32(b), rowbytes);
        uint64_t bo = 0;
        if (base_off_mode == 1) bo = (a_addr >> 7) & 7;
        if (base_off_mode == 2) bo = (a_addr >> 7) & mask;
        da |= bo << 49;
        ptx::mma_i8_ss(td, da, db, ptx::make_idesc_i8(128, 16), 0);
        ptx::mma_commit(&bar);
        ptx::mbar_wait(&bar, 0, &tflag);
    }
    __syncthreads();
    ptx::tc_fence_after();
    uint32_t v[16];
    ptx::tmem_ld_32x32b_x16(td + ((uint32_t)(warp * 32) << 16), v);
    ptx::tmem_ld_wait();
    for (int j = 0; j < 16; ++j) out[(warp * 32 + lane) * 16 + j] = (int)v[j];
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) ptx::tmem_dealloc(td, 32);
}

int main()
{
    int* d;
    cudaMalloc(&d, 128 * 16 * 4);
    static int h[128 * 16];
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 49152);
    for (int rb : {128, 64, 32}) {
        for (int mode = 0; mode < 3; ++mode) {
            printf("rowbytes %3d base_off_mode %d :", rb, mode);
            for (int shift = 0; shift <= 17; ++shift) {
                cudaMemset(d, 0xff, sizeof h);
                k<<<1, 128, 49152>>>(rb, shift, mode, d);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf(" ERR(%s)", cudaGetErrorString(e)); break; }
                cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
                int bad = 0;
                for (int i = 0; i < 128; ++i)
                    for (int n = 0; n < 16; ++n)
                        if (h[i * 16 + n] != aval(i + shift, n)) ++bad;
                printf(" %s", bad ? "x" : "OK");
            }
            printf("\n");
        }
    }
    return 0;
}
