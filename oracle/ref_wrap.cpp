// ref_wrap.cpp — builds the REFERENCE's own CPU int8 convolution, unmodified, into oracle/_ref/.
//
// TEST INFRASTRUCTURE ONLY (see oracle/cpu_ref.c header).  This translation unit is ours; the
// reference header is #included from where it lies under /root/reference and is never copied.
// It only compiles in the authoring container (the GPU box has no /root/reference); the built
// oracle/_ref/libref_conv.so travels there with the repo snapshot.
//
// refConv2DForward<batch,inC,inH,inW,outC,outH,outW,kH,kW> (cpp/int8conv/refConv2DForward.hpp:56-80)
// takes every size as a template parameter, so a fixed table of shapes is instantiated below and
// looked up at run time.  TensorOptions::is_variable() (refConv2DForward.hpp:67) was removed from
// torch after 1.x; the one-line macro shim maps it onto requires_grad(false), as SURVEY.md 8c records.
#include <ATen/ATen.h>
#include <c10/macros/Macros.h>
#include <c10/util/Logging.h>
#include <torch/extension.h>
#include <iostream>
#include <sstream>
#include <cstring>
#include <omp.h>

#define is_variable(x) requires_grad(false)
#include "/root/reference/cpp/int8conv/refConv2DForward.hpp"
#undef is_variable

// (batch, inC, inH, inW, outC, outH, outW, kH, kW)
#define REF_SHAPES(X)                       \
    X(2, 16, 10, 10, 16, 8, 8, 3, 3)        \
    X(1, 32, 9, 12, 16, 7, 10, 3, 3)        \
    X(1, 64, 6, 6, 32, 6, 6, 1, 1)          \
    X(1, 8, 11, 11, 4, 5, 5, 7, 7)          \
    X(1, 3, 15, 15, 8, 9, 9, 7, 7)          \
    X(3, 16, 6, 7, 48, 4, 5, 3, 3)          \
    X(1, 256, 4, 4, 64, 4, 4, 1, 1)         \
    X(1, 64, 10, 10, 64, 8, 8, 3, 3)        \
    X(1, 128, 6, 6, 128, 4, 4, 3, 3)        \
    X(1, 64, 18, 18, 64, 16, 16, 3, 3)      \
    X(1, 64, 58, 58, 64, 56, 56, 3, 3)

namespace {
struct CoutSilencer {  // the reference prints "\tBatch b" per image (refConv2DForward.hpp:29)
    std::streambuf* old;
    std::ostringstream sink;
    CoutSilencer() : old(std::cout.rdbuf(sink.rdbuf())) {}
    ~CoutSilencer() { std::cout.rdbuf(old); }
};
}  // namespace

extern "C" {

int ref_num_shapes(void)
{
    int n = 0;
#define X(...) ++n;
    REF_SHAPES(X)
#undef X
    return n;
}

// Writes the i-th instantiated shape into out[9]; returns 0 on success.
int ref_shape(int i, int* out)
{
    int idx = 0;
#define X(b, ic, ih, iw, oc, oh, ow, kh, kw)                                              \
    if (idx++ == i) { int s[9] = {b, ic, ih, iw, oc, oh, ow, kh, kw}; std::memcpy(out, s, sizeof s); return 0; }
    REF_SHAPES(X)
#undef X
    return 1;
}

int ref_max_threads(void) { return omp_get_max_threads(); }

// Runs the reference on NCHW int8 input / OIHW int8 kernel; int32 NCHW output.
// Returns 0 on success, 1 if the shape is not in the table, 2 on a torch exception.
int ref_conv2d_forward(int batch, int inC, int inH, int inW, int outC, int outH, int outW, int kH, int kW,
                       const int8_t* input, const int8_t* kernel, int32_t* output)
{
    try {
        auto in = at::from_blob(const_cast<int8_t*>(input), {batch, inC, inH, inW}, at::kChar);
        auto ke = at::from_blob(const_cast<int8_t*>(kernel), {outC, inC, kH, kW}, at::kChar);
        CoutSilencer quiet;
#define X(b, ic, ih, iw, oc, oh, ow, kh, kw)                                                           \
        if (batch == b && inC == ic && inH == ih && inW == iw && outC == oc && outH == oh && outW == ow && \
            kH == kh && kW == kw) {                                                                        \
            auto out = refConv2DForward<b, ic, ih, iw, oc, oh, ow, kh, kw>(in, ke);                        \
            std::memcpy(output, out.data_ptr(), sizeof(int32_t) * (size_t)out.numel());                    \
            return 0;                                                                                      \
        }
        REF_SHAPES(X)
#undef X
        return 1;
    } catch (const std::exception& e) {
        std::cerr << "ref_conv2d_forward: " << e.what() << std::endl;
        return 2;
    }
}

}  // extern "C"
