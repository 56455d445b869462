// check.cpp — the reference's self-check driver (cpp/int8conv/check.cu:62-155, checkForward3x3) against the new
// operator: for WARMUP fresh random draws assert  int32 custom conv == round(fp32 library conv)  EXACTLY
// (check.cu:114-129), then time REPEATS repeats and print the mean ms of both (check.cu:138-154).
// Same shape constants as the reference (check.cu:28-41); WARMUP / REPEATS can be overridden on the command line
// because the reference's 1000 + 1000 iterations take minutes.   usage: check [warmup] [repeats]
#include <torch/torch.h>

#include <iostream>
#include <map>
#include <numeric>
#include <vector>

#include "conv2DForward3x3.hpp"

using namespace lowbit;

constexpr uint32_t BATCH = 16, IN_CHANNELS = 128, IN_H = 130, IN_W = 130;
constexpr uint32_t OUT_CHANNELS = 128, OUT_H = 128, OUT_W = 128, KH = 3, KW = 3;
constexpr float RMUL = 1, RSUB = 0;   // values in {0,1}: the fp32 library conv is exact (check.cu:43-44)

static at::Tensor randomTensor(at::IntArrayRef shape)
{
    return torch::rand(shape, torch::device(torch::kCUDA).dtype(torch::kFloat32)).mul_(RMUL).sub_(RSUB).round_()
        .to(torch::kInt8).to(torch::kFloat32);
}

static std::tuple<at::Tensor, float> libraryConv2DFloat(const at::Tensor& data, const at::Tensor& kernel)
{   // role of cudnnConv2DFloat (cudnn2DConvolution.cuh:95-151): the independent truth, event-timed
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    auto stream = at::cuda::getCurrentCUDAStream().stream();
    cudaEventRecord(a, stream);
    auto out = at::conv2d(data, kernel);
    cudaEventRecord(b, stream);
    cudaEventSynchronize(b);
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    cudaEventDestroy(a); cudaEventDestroy(b);
    return {out, ms};
}

int main(int argc, char** argv)
{
    const int WARMUP = argc > 1 ? atoi(argv[1]) : 20;
    const int REPEATS = argc > 2 ? atoi(argv[2]) : 50;
    torch::NoGradGuard ng;
    at::globalContext().setAllowTF32CuDNN(false);
    std::cout << "FORWARD 3x3" << std::endl;
    at::Tensor kernel, data, kernel_v, data_v;
    std::cout << "WARMUP" << std::endl;
    for (int i = 0; i < WARMUP; ++i) {
        data = randomTensor({BATCH, IN_CHANNELS, IN_H, IN_W});
        kernel = randomTensor({OUT_CHANNELS, IN_CHANNELS, KH, KW});
        data_v = to_vect_c(data.to(torch::kInt8));
        kernel_v = to_vect_c(kernel.to(torch::kInt8));
        auto fres = libraryConv2DFloat(data, kernel);
        auto qres = conv2DForward3x3<BATCH, IN_CHANNELS, OUT_CHANNELS, IN_H, IN_W, OUT_H, OUT_W>(data_v, kernel_v);
        auto diff = (to_vect_c(std::get<0>(fres).round_().to(torch::kInt32)) - std::get<0>(qres)).abs();
        if (diff.max().item<int>() != 0) {
            std::cout << "MISMATCH at iteration " << i << " max diff " << diff.max().item<int>() << std::endl;
            return 1;
        }
    }
    std::map<std::string, std::vector<float>> timings;
    std::cout << "BENCHMARKING" << std::endl;
    for (int i = 0; i < REPEATS; ++i) {
        timings["float"].push_back(std::get<1>(libraryConv2DFloat(data, kernel)));
        timings["int8"].push_back(std::get<1>(
            conv2DForward3x3<BATCH, IN_CHANNELS, OUT_CHANNELS, IN_H, IN_W, OUT_H, OUT_W>(data_v, kernel_v)));
    }
    for (const auto& kv : timings) {
        double mean = std::accumulate(kv.second.begin(), kv.second.end(), 0.0) / kv.second.size();
        std::cout << kv.first << " : " << mean << std::endl;
    }
    std::cout << "CHECK OK (" << WARMUP << " exact comparisons)" << std::endl;
    return 0;
}
