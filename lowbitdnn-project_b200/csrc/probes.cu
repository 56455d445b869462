// probes.cu — on-device peak probes used as roofline denominators by the harness
// (the cpp/libbenchmark role: cpp/libbenchmark/benchmark.cu:36-184 times library calls; here the probes
// measure what the silicon can do so per-layer numbers can be quoted as fractions).
#include "common.cuh"
#include "ptx.cuh"

#include <mutex>

namespace lbc {

namespace {

constexpr int kProbeN = 256;
constexpr int kProbeThreads = 128;

__device__ int g_probe_timeout = 0;

// Every CTA: 128x256x128-byte operand tiles in smem (zeros), one thread issues `iters` x 4 MMAs
// (M=128, N=256, K=32 int8 each) back to back into one TMEM accumulator (pseudo-random operands).  No loads, no epilogue.
__global__ void __launch_bounds__(kProbeThreads, 1) mma_i8_peak_kernel(int32_t iters)
{
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* a = smem;                       // 128 rows x 128 B
    uint8_t* b = smem + 128 * 128;           // 256 rows x 128 B
    __shared__ uint64_t done_bar;
    __shared__ uint32_t tmem_base_s;

    // pseudo-random operand bytes: all-zero operands draw far less power than real data and would let the clocks sit
    // at boost, overstating what a kernel working on real activations can reach
    for (int i = threadIdx.x; i < (128 + kProbeN) * 128 / 16; i += blockDim.x) {
        uint32_t h = (uint32_t)i * 2654435761u + blockIdx.x * 40503u + 12345u;
        int4 v;
        h ^= h >> 15; h *= 2246822519u; v.x = (int)h;
        h ^= h >> 13; h *= 3266489917u; v.y = (int)h;
        h ^= h >> 16; h *= 668265263u;  v.z = (int)h;
        h ^= h >> 15; h *= 374761393u;  v.w = (int)h;
        reinterpret_cast<int4*>(smem)[i] = v;
    }
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        ptx::mbar_init(&done_bar, 1);
        ptx::fence_barrier_init();
    }
    if (warp == 1) {
        ptx::tmem_alloc(&tmem_base_s, 256);
        ptx::tmem_relinquish();
    }
    ptx::fence_proxy_async();   // generic-proxy smem writes -> visible to the tensor-core (async) proxy
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_d = tmem_base_s;
    if (warp == 0 && lane == 0) {
        const uint32_t idesc = ptx::make_idesc_i8(128, kProbeN);
        const uint64_t da = ptx::make_kmajor_desc(ptx::smem_u32(a), 128);
        const uint64_t db = ptx::make_kmajor_desc(ptx::smem_u32(b), 128);
        for (int32_t it = 0; it < iters; ++it) {
#pragma unroll
            for (uint32_t k = 0; k < 4; ++k) ptx::mma_i8_ss(tmem_d, da + 2ull * k, db + 2ull * k, idesc, (it | k) ? 1u : 0u);
        }
        ptx::mma_commit(&done_bar);
        ptx::mbar_wait(&done_bar, 0, &g_probe_timeout, 4000000000ull);
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_d, 256);
    }
}

__global__ void __launch_bounds__(256) copy_kernel(const int4* __restrict__ src, int4* __restrict__ dst, size_t n)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        dst[i] = src[i];
}

void* g_flush_buf = nullptr;
size_t g_flush_bytes = 0;
std::mutex g_flush_mu;

}  // namespace

lbc_status probe_int8_mma_peak(int32_t iters, double* tops, cudaStream_t stream)
{
    DeviceInfo dev;
    lbc_status st = current_device(&dev);
    if (st != LBC_OK) return st;
    LBC_REQUIRE(iters > 0 && tops, LBC_ERR_INVALID_ARG, "probe: bad arguments");
    const size_t smem = 1024 + (128 + kProbeN) * 128;
    LBC_CUDA_TRY(cudaFuncSetAttribute(mma_i8_peak_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaEvent_t e0, e1;
    LBC_CUDA_TRY(cudaEventCreate(&e0));
    LBC_CUDA_TRY(cudaEventCreate(&e1));
    mma_i8_peak_kernel<<<dev.sm_count, kProbeThreads, smem, stream>>>(iters / 8 + 1);   // warm-up
    LBC_CUDA_TRY(cudaEventRecord(e0, stream));
    mma_i8_peak_kernel<<<dev.sm_count, kProbeThreads, smem, stream>>>(iters);
    LBC_CUDA_TRY(cudaEventRecord(e1, stream));
    LBC_CUDA_TRY(cudaEventSynchronize(e1));
    LBC_CUDA_TRY(cudaGetLastError());
    float ms = 0.f;
    LBC_CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    int flag = 0;
    LBC_CUDA_TRY(cudaMemcpyFromSymbol(&flag, g_probe_timeout, sizeof(int)));
    LBC_REQUIRE(flag == 0, LBC_ERR_KERNEL_TIMEOUT, "mma peak probe: watchdog fired");
    const double ops = 2.0 * 128 * kProbeN * 32 * 4.0 * (double)iters * dev.sm_count;
    *tops = ops / (ms * 1e-3) / 1e12;
    return LBC_OK;
}

lbc_status probe_hbm_copy(size_t bytes, int32_t iters, double* gbs, cudaStream_t stream)
{
    DeviceInfo dev;
    lbc_status st = current_device(&dev);
    if (st != LBC_OK) return st;
    LBC_REQUIRE(bytes >= (1u << 20) && iters > 0 && gbs, LBC_ERR_INVALID_ARG, "probe: bad arguments");
    bytes &= ~size_t(15);
    void *a = nullptr, *b = nullptr;
    if (cudaMalloc(&a, bytes) != cudaSuccess || cudaMalloc(&b, bytes) != cudaSuccess) {
        cudaFree(a);
        set_error("probe_hbm_copy: cudaMalloc of 2 x %zu bytes failed", bytes);
        return LBC_ERR_ALLOC;
    }
    cudaMemsetAsync(a, 1, bytes, stream);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const unsigned grid = dev.sm_count * 16;
    copy_kernel<<<grid, 256, 0, stream>>>((const int4*)a, (int4*)b, bytes / 16);
    cudaEventRecord(e0, stream);
    for (int i = 0; i < iters; ++i) copy_kernel<<<grid, 256, 0, stream>>>((const int4*)a, (int4*)b, bytes / 16);
    cudaEventRecord(e1, stream);
    cudaError_t e = cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(a);
    cudaFree(b);
    LBC_CUDA_TRY(e);
    *gbs = 2.0 * (double)bytes * iters / (ms * 1e-3) / 1e9;
    return LBC_OK;
}

lbc_status flush_l2(cudaStream_t stream)
{
    std::lock_guard<std::mutex> lk(g_flush_mu);
    if (!g_flush_buf) {
        g_flush_bytes = 256u << 20;   // > 126 MB L2
        if (cudaMalloc(&g_flush_buf, g_flush_bytes) != cudaSuccess) {
            g_flush_buf = nullptr;
            set_error("flush_l2: cudaMalloc failed");
            return LBC_ERR_ALLOC;
        }
    }
    LBC_CUDA_TRY(cudaMemsetAsync(g_flush_buf, 0, g_flush_bytes, stream));
    return LBC_OK;
}

}  // namespace lbc
