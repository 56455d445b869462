#include "benchmark.hpp"

#include <cuda_runtime.h>

#include <cmath>
#include <random>
#include <stdexcept>
#include <vector>

#include "../../../include/lowbit_cnn.h"

namespace lowbit {

namespace {
void check(lbc_status st, const char* what)
{
    if (st != LBC_OK) throw std::runtime_error(std::string(what) + ": " + lbc_last_error_string());
}
void check(cudaError_t e, const char* what)
{
    if (e != cudaSuccess) throw std::runtime_error(std::string(what) + ": " + cudaGetErrorString(e));
}
struct DevBuf {
    void* p = nullptr;
    explicit DevBuf(size_t n) { check(cudaMalloc(&p, n ? n : 1), "cudaMalloc"); }
    ~DevBuf() { cudaFree(p); }
};
}  // namespace

ConvTiming benchmark_convolution(size_t B, size_t C, size_t H, size_t W, size_t numFilters, size_t filterH,
                                 size_t filterW, size_t padH, size_t padW, size_t strideH, size_t strideW,
                                 size_t dilationH, size_t dilationW, size_t groups, bool out_int32, int repeats,
                                 int verbose)
{
    lbc_conv_desc d{};
    d.n = (int32_t)B; d.h = (int32_t)H; d.w = (int32_t)W; d.c = (int32_t)C;
    d.k = (int32_t)numFilters; d.r = (int32_t)filterH; d.s = (int32_t)filterW;
    d.stride_h = (int32_t)strideH; d.stride_w = (int32_t)strideW; d.pad_h = (int32_t)padH; d.pad_w = (int32_t)padW;
    d.dil_h = (int32_t)dilationH; d.dil_w = (int32_t)dilationW; d.groups = (int32_t)groups;
    d.relu = 1; d.out_mode = out_int32 ? LBC_OUT_INT32 : LBC_OUT_INT8;

    ConvTiming t;
    int32_t p = 0, q = 0;
    check(lbc_conv_out_shape(&d, &p, &q), "lbc_conv_out_shape");
    check(lbc_conv_work(&d, &t.ops, &t.bytes), "lbc_conv_work");
    lbc_plan* plan = nullptr;
    check(lbc_conv_plan_create(&d, LBC_KERNEL_AUTO, &plan), "lbc_conv_plan_create");
    char desc[512];
    lbc_conv_plan_describe(plan, desc, sizeof desc);
    t.plan = desc;
    if (verbose) fprintf(stderr, "%s\n", desc);
    try {
        const size_t cg = C / groups;
        const size_t xbytes = B * H * W * C, wraw = numFilters * filterH * filterW * cg;
        const size_t ybytes = B * (size_t)p * q * numFilters * (out_int32 ? 4 : 1);
        size_t wpacked = 0;
        check(lbc_conv_packed_weight_bytes(plan, &wpacked), "lbc_conv_packed_weight_bytes");
        DevBuf x(xbytes), wr(wraw), wp(wpacked), y(ybytes), bias(4 * numFilters), scale(4 * numFilters);
        // synthetic data: deterministic, full int8 range
        std::mt19937 rng(1234);
        std::vector<int8_t> hx(xbytes), hw(wraw);
        for (auto& v : hx) v = (int8_t)(rng() & 0xff);
        for (auto& v : hw) v = (int8_t)((rng() % 255) - 127);
        std::vector<int32_t> hb(numFilters);
        std::vector<float> hs(numFilters, 0.0078125f / std::sqrt((float)(filterH * filterW * cg)));
        for (auto& v : hb) v = (int32_t)(rng() % 65536) - 32768;
        check(cudaMemcpy(x.p, hx.data(), xbytes, cudaMemcpyHostToDevice), "cudaMemcpy");
        check(cudaMemcpy(wr.p, hw.data(), wraw, cudaMemcpyHostToDevice), "cudaMemcpy");
        check(cudaMemcpy(bias.p, hb.data(), 4 * numFilters, cudaMemcpyHostToDevice), "cudaMemcpy");
        check(cudaMemcpy(scale.p, hs.data(), 4 * numFilters, cudaMemcpyHostToDevice), "cudaMemcpy");
        check(lbc_conv_prepack_weights(plan, (const int8_t*)wr.p, LBC_W_KRSC, wp.p, nullptr), "lbc_conv_prepack_weights");
        check(lbc_conv_run(plan, (const int8_t*)x.p, wp.p, (const int32_t*)bias.p, (const float*)scale.p, y.p, nullptr, nullptr),
              "lbc_conv_run (warm-up)");
        check(cudaDeviceSynchronize(), "cudaDeviceSynchronize");
        float best_ms = 1e30f;
        auto best_wall = std::chrono::microseconds::max();
        for (int i = 0; i < (repeats < 1 ? 1 : repeats); ++i) {
            float ms = 0.f;
            const auto t0 = std::chrono::high_resolution_clock::now();
            check(lbc_conv_run(plan, (const int8_t*)x.p, wp.p, (const int32_t*)bias.p, (const float*)scale.p, y.p, nullptr, &ms),
                  "lbc_conv_run");
            check(cudaDeviceSynchronize(), "cudaDeviceSynchronize");
            const auto t1 = std::chrono::high_resolution_clock::now();
            best_wall = std::min(best_wall, std::chrono::duration_cast<std::chrono::microseconds>(t1 - t0));
            best_ms = std::min(best_ms, ms);
        }
        t.wall = best_wall;
        t.device_ms = best_ms;
    } catch (...) {
        lbc_conv_plan_destroy(plan);
        throw;
    }
    lbc_conv_plan_destroy(plan);
    return t;
}

CpuTiming benchmark_cpu_baseline(const std::function<void()>& fn, int threads, int repeats)
{
    CpuTiming t;
    t.threads = threads;
    auto best = std::chrono::microseconds::max();
    for (int i = 0; i < (repeats < 1 ? 1 : repeats); ++i) {
        const auto t0 = std::chrono::high_resolution_clock::now();
        fn();
        const auto t1 = std::chrono::high_resolution_clock::now();
        best = std::min(best, std::chrono::duration_cast<std::chrono::microseconds>(t1 - t0));
    }
    t.wall = best;
    return t;
}

}  // namespace lowbit
