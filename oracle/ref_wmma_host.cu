// ref_wmma_host.cu — TEST INFRASTRUCTURE ONLY (see cpu_ref.c): a torch-free host for the REFERENCE's own tensor-core
// kernel, so that "the reference on this GPU" has a measured line next to liblowbit-cnn (bench.py --impl reference-gpu).
//
// The kernel body is NOT in this repository: oracle/Makefile (target `ref_gpu`) cuts
//     cpp/int8conv/conv2DForward3x3TensorCores.cuh:18,24-26   (namespace alias + WMMA_M/N/K)
//     cpp/int8conv/conv2DForward3x3TensorCores.cuh:537-693    (CUDAConv2DForward3x3TensorCoures<...>, wmma m32n8k16)
// out of /root/reference where it lies into oracle/_ref/ref_wmma_kernel.inc (git-ignored, generated) and compiles it
// unmodified for sm_100a.  What this file restates is only the reference's host wrapper
// (conv2DForward3x3<batch,inC,outC,inH,inW,outH,outW>, same file :695-751): the launch geometry
//     grid (outH/32 * outW/32, batch, outC/16), block (32, 16), static shared memory, cudaEvent timing
// without at::Tensor.  Shape: the reference's own benchmark shape (check.cu:31-41).
#include <cuda_runtime.h>
#include <mma.h>
#include <stdint.h>

#include "_ref/ref_wmma_kernel.inc"

namespace {
constexpr uint32_t kBatch = 16, kInC = 128, kOutC = 128, kInH = 130, kInW = 130, kOutH = 128, kOutW = 128;
}

extern "C" {

// the reference's constants, for the caller's buffers
void ref_wmma_shape(int32_t* s)
{
    s[0] = kBatch; s[1] = kInC; s[2] = kInH; s[3] = kInW; s[4] = kOutC; s[5] = kOutH; s[6] = kOutW;
}

// data  int8 [N][C/16][H][W][16], kernel int8 [K][C/16][3][3][16], result int32 [N][K/16][P][Q][16] (device pointers).
// Runs `iters` launches between two events; returns the mean milliseconds per launch (< 0 on a CUDA error).
float ref_wmma_conv3x3(const int8_t* data, const int8_t* kernel, int32_t* result, int iters)
{
    static_assert(kOutH % 32 == 0 && kOutW % 32 == 0 && kInC % 16 == 0 && kOutC % 16 == 0, "reference tile constraints (:724)");
    const dim3 grid(kOutH / 32 * (kOutW / 32), kBatch, kOutC / 16), block(32, 16);
    cudaEvent_t a, b;
    if (cudaEventCreate(&a) != cudaSuccess || cudaEventCreate(&b) != cudaSuccess) return -1.f;
    cudaEventRecord(a);
    for (int i = 0; i < iters; ++i)
        CUDAConv2DForward3x3TensorCoures<kBatch, kInC / 16, kOutC / 16, 3, 3, kInH, kInW, kOutH, kOutW, 32, 32>
            <<<grid, block>>>(data, kernel, result);
    cudaEventRecord(b);
    float ms = -1.f;
    if (cudaEventSynchronize(b) == cudaSuccess && cudaGetLastError() == cudaSuccess) {
        cudaEventElapsedTime(&ms, a, b);
        ms /= (float)iters;
    }
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    return ms;
}

}  // extern "C"
