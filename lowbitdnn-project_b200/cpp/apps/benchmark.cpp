// benchmark.cpp — JSON-driven convolution sweep, the counterpart of the reference's cpp/apps/benchmark.cpp:109-169:
// reads config.json {repeats, configs{name: dtype/format description}, experiments[{batch[], channels[], height,
// width, filters[], filter_width, filter_height, verbose, configs[]}]}, skips filters < channels (benchmark.cpp:140),
// runs every (batch, channels, filters, config) through libbenchmark with pad = stride = dilation = 1
// (benchmark.cpp:73) and writes output.json {repeats, configs, benchmarks:[{B,C,H,W,filters,filter_width,
// filter_height,config,name,timing}]} (benchmark.cpp:84-107,162-167).  `timing` stays the mean wall-clock
// microseconds per call; device_ms / tops / gbs / plan are added next to it.
//   usage: benchmark [config.json] [output.json] [--limit N]
#include <cstring>
#include <fstream>
#include <iostream>
#include <numeric>
#include <sstream>

#include "../libbenchmark/benchmark.hpp"
#include "mini_json.hpp"

using mini_json::Value;

static std::string shortName(size_t B, size_t C, size_t H, size_t W, size_t K, size_t fh, size_t fw, const std::string& cfg)
{
    std::stringstream s;
    s << B << 'x' << C << 'x' << H << 'x' << W << " * " << K << 'x' << fh << 'x' << fw << ", " << cfg;
    return s.str();
}

int main(int argc, char** argv)
{
    std::string in = "config.json", out = "output.json";
    long limit = -1;
    int pos = 0;
    for (int i = 1; i < argc; ++i) {
        if (!std::strcmp(argv[i], "--limit") && i + 1 < argc) limit = std::atol(argv[++i]);
        else if (pos++ == 0) in = argv[i];
        else out = argv[i];
    }
    Value root;
    try {
        std::ifstream f(in);
        if (!f) throw std::runtime_error("cannot open " + in);
        std::stringstream ss; ss << f.rdbuf();
        const std::string text = ss.str();
        root = mini_json::Parser(text).parse();
    } catch (const std::exception& e) {
        std::cerr << "Config file " << in << " is not found or is not correct: " << e.what() << '\n';
        return EXIT_FAILURE;
    }
    const int repeats = (int)root.at("repeats").num;
    const Value& configs = root.at("configs");
    std::vector<std::string> rows;
    long done = 0;
    for (const Value& exp : root.at("experiments").arr) {
        const size_t H = (size_t)exp.at("height").num, W = (size_t)exp.at("width").num;
        const size_t fw = (size_t)exp.at("filter_width").num, fh = (size_t)exp.at("filter_height").num;
        const int verbose = exp.has("verbose") ? (int)exp.at("verbose").num : 0;
        for (double B : exp.at("batch").numbers())
            for (double C : exp.at("channels").numbers())
                for (double K : exp.at("filters").numbers()) {
                    if (K < C) continue;
                    for (const Value& cn : exp.at("configs").arr) {
                        if (limit >= 0 && done >= limit) goto finished;
                        const std::string name = shortName((size_t)B, (size_t)C, H, W, (size_t)K, fh, fw, cn.str);
                        try {
                            const Value& cfg = configs.at(cn.str);
                            const bool ext = cfg.at("output_data_type").str == "int32";
                            std::clog << name << "..." << std::endl;
                            std::vector<double> wall;
                            lowbit::ConvTiming t{};
                            for (int r = 0; r < repeats; ++r) {   // one fresh set-up per repeat, like the reference
                                t = lowbit::benchmark_convolution((size_t)B, (size_t)C, H, W, (size_t)K, fh, fw, 1, 1, 1, 1, 1, 1,
                                                                  1, ext, 1, verbose >= 2 && r == 0);
                                wall.push_back((double)t.wall.count());
                            }
                            const double mean = std::accumulate(wall.begin(), wall.end(), 0.0) / wall.size();
                            std::stringstream row;
                            row << "{\"B\": " << B << ", \"C\": " << C << ", \"H\": " << H << ", \"W\": " << W << ", \"filters\": " << K
                                << ", \"filter_width\": " << fw << ", \"filter_height\": " << fh << ", \"config\": "
                                << mini_json::quote(cn.str) << ", \"name\": " << mini_json::quote(name) << ", \"timing\": " << mean
                                << ", \"device_ms\": " << t.device_ms << ", \"tops\": " << t.ops / (t.device_ms * 1e9)
                                << ", \"gbs\": " << t.bytes / (t.device_ms * 1e6) << ", \"plan\": " << mini_json::quote(t.plan) << "}";
                            rows.push_back(row.str());
                            std::clog << name << " timing is " << mean << std::endl;
                            ++done;
                        } catch (const std::exception& e) {
                            std::cerr << "Failed to perform experiment: " << e.what() << '\n';
                        }
                    }
                }
    }
finished:
    std::ofstream o(out);
    o << "{\n  \"repeats\": " << repeats << ",\n  \"configs\": " << mini_json::dump(configs) << ",\n  \"benchmarks\": [\n";
    for (size_t i = 0; i < rows.size(); ++i) o << "    " << rows[i] << (i + 1 < rows.size() ? ",\n" : "\n");
    o << "  ]\n}\n";
    std::clog << "Finished." << std::endl;
    return EXIT_SUCCESS;
}
