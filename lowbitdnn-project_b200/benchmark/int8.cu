// int8.cu — int8 convolution benchmark driver (the file the reference left empty: benchmark/int8.cu:1-4).
//
//   int8_bench [--network resnet50|resnet18|vgg16|mobilenet_v2|single_3x3] [--batch N] [--warmup W] [--repeats R]
//              [--layer NAME]      run only that layer (short command line for an ncu capture:
//                                  ncu --set full --clock-control none -k regex:igemm -c 3 int8_bench --layer l3.1.conv2)
//
// Prints the per-layer table (mean us, TOPS, GB/s, fraction of the measured int8 MMA peak and of the measured copy
// bandwidth, planner decision) and the network line (images/s).  All compute goes through the liblowbit-cnn C ABI.
#include <cmath>
#include <cstring>

#include "conv_benchmark.cuh"

int main(int argc, char** argv)
{
    using namespace lowbit::bench;
    std::string net = "resnet50", only;
    int batch = 0, warmup = 3, repeats = 10;
    for (int i = 1; i < argc; ++i) {
        auto arg = [&](const char* f) { return !std::strcmp(argv[i], f) && i + 1 < argc; };
        if (arg("--network")) net = argv[++i];
        else if (arg("--batch")) batch = std::atoi(argv[++i]);
        else if (arg("--warmup")) warmup = std::atoi(argv[++i]);
        else if (arg("--repeats")) repeats = std::atoi(argv[++i]);
        else if (arg("--layer")) only = argv[++i];
        else { std::fprintf(stderr, "unknown argument %s\n", argv[i]); return 2; }
    }
    try {
        std::vector<Layer> L = network(net, batch);
        if (!only.empty()) {
            std::vector<Layer> one;
            for (auto& l : L)
                if (l.name == only) { one.push_back(l); one.back().input_of = -1; }
            if (one.empty()) throw std::runtime_error("no layer named " + only);
            L = one;
        }
        const Peaks pk = measure_peaks();
        double net_ms = 0;
        const auto res = run_network(L, warmup, repeats, &net_ms);
        print_table(res, pk, net_ms, L[0].d.n);
    } catch (const std::exception& e) {
        std::fprintf(stderr, "int8_bench: %s\n", e.what());
        return 1;
    }
    return 0;
}
