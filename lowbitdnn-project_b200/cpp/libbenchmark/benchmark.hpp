// benchmark.hpp — timing library (cpp/libbenchmark/benchmark.cuh:34-84 role) on top of liblowbit-cnn.
//
// benchmark_convolution keeps the reference entry point's shape arguments and its "one call = set everything up,
// run ONE convolution between two device synchronisations, return wall-clock microseconds" contract
// (cpp/libbenchmark/benchmark.cu:36-184), but the convolution is liblowbit-cnn's int8 kernel instead of
// cudnnConvolutionForward, and failures are reported as exceptions instead of "0 us".
// benchmark_cpu_baseline times a caller-supplied CPU implementation of the same layer (in the reference tree:
// refConv2DForward, cpp/int8conv/refConv2DForward.hpp) so the CPU number sits beside the GPU one.
#pragma once
#include <chrono>
#include <cstddef>
#include <functional>
#include <string>

namespace lowbit {

struct ConvTiming {
    std::chrono::microseconds wall{0};   // host wall clock around one launch (reference contract)
    float device_ms = 0.f;               // CUDA-event time of the kernel alone
    double ops = 0, bytes = 0;           // algorithmic work
    std::string plan;                    // planner's description of the kernel/tiling used
};

// int8 NHWC activations, int8 KRSC weights, int32 bias, per-channel fp32 scale, ReLU -> int8 (out_int32 == false)
// or raw int32 accumulators (out_int32 == true).  Throws std::runtime_error on any failure.
ConvTiming benchmark_convolution(size_t B, size_t C, size_t H, size_t W, size_t numFilters, size_t filterH,
                                 size_t filterW, size_t padH, size_t padW, size_t strideH, size_t strideW,
                                 size_t dilationH, size_t dilationW, size_t groups = 1, bool out_int32 = false,
                                 int repeats = 1, int verbose = 0);

// Wall-clock of `fn` (a CPU implementation of one layer), best of `repeats`; reports the thread count it is told.
struct CpuTiming {
    std::chrono::microseconds wall{0};
    int threads = 1;
};
CpuTiming benchmark_cpu_baseline(const std::function<void()>& fn, int threads, int repeats = 1);

}  // namespace lowbit
