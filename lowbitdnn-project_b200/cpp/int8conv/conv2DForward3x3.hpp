// conv2DForward3x3.hpp — drop-in for the reference operator
//     std::tuple<Tensor,float> conv2DForward3x3<batch,inC,outC,inH,inW,outH,outW>(const Tensor&, const Tensor&)
// (cpp/int8conv/conv2DForward3x3TensorCores.cuh:695-751): VECT_C=16 int8 input [N][C/16][H][W][16], VECT_C=16 int8
// kernel [K][C/16][3][3][16], 3x3 / stride 1 / no padding, int32 output [N][K/16][P][Q][16], elapsed device ms as
// the second tuple element.  Same template parameters, same tensors, same (Tensor, ms) result — but the work is done
// by liblowbit-cnn's tcgen05 kernel through the C ABI, and the reference's outH,outW % 32 == 0 restriction is gone.
// Additionally exposes the fused int8 epilogue the reference never had (conv2DForward3x3Requant).
#pragma once
#include <tuple>

#include "utils.hpp"

namespace lowbit {

namespace detail {
struct PlanGuard {
    lbc_plan* p = nullptr;
    ~PlanGuard() { lbc_conv_plan_destroy(p); }
};

inline std::tuple<at::Tensor, float> run3x3(const at::Tensor& rinput, const at::Tensor& rkernel, int32_t out_mode,
                                            const at::Tensor* bias, const at::Tensor* scale, bool relu)
{
    auto input = rinput.contiguous();
    auto kernel = rkernel.contiguous();
    TORCH_CHECK(input.is_cuda() && kernel.is_cuda(), "conv2DForward3x3: CUDA tensors required (no CPU fallback)");
    TORCH_CHECK(input.scalar_type() == at::kChar && kernel.scalar_type() == at::kChar, "conv2DForward3x3: int8 tensors");
    TORCH_CHECK(input.dim() == 5 && kernel.dim() == 5 && input.size(4) == VECT_C && kernel.size(4) == VECT_C,
                "conv2DForward3x3: expects VECT_C=16 tensors");
    TORCH_CHECK(kernel.size(2) == 3 && kernel.size(3) == 3 && kernel.size(1) == input.size(1), "conv2DForward3x3: 3x3 kernel");
    const int32_t n = (int32_t)input.size(0), c = (int32_t)input.size(1) * VECT_C, h = (int32_t)input.size(2),
                  w = (int32_t)input.size(3), k = (int32_t)kernel.size(0);
    lbc_stream st = current_stream();

    // VECT_C -> NHWC (activations) / KRSC (filters): a per-pixel regroup of 16-byte channel chunks
    auto x = at::empty({n, h, w, c}, input.options());
    lbc_throw(lbc_vect_c_to_nhwc(input.data_ptr(), x.data_ptr(), n, c, h, w, VECT_C, 1, st), "lbc_vect_c_to_nhwc");
    auto wk = at::empty({k, 3, 3, c}, kernel.options());
    lbc_throw(lbc_vect_c_to_nhwc(kernel.data_ptr(), wk.data_ptr(), k, c, 3, 3, VECT_C, 1, st), "lbc_vect_c_to_nhwc");

    lbc_conv_desc d{};
    d.n = n; d.h = h; d.w = w; d.c = c; d.k = k; d.r = 3; d.s = 3;
    d.stride_h = d.stride_w = 1; d.dil_h = d.dil_w = 1; d.groups = 1;
    d.relu = relu; d.out_mode = out_mode;
    PlanGuard plan;
    lbc_throw(lbc_conv_plan_create(&d, LBC_KERNEL_AUTO, &plan.p), "lbc_conv_plan_create");
    size_t wbytes = 0;
    lbc_throw(lbc_conv_packed_weight_bytes(plan.p, &wbytes), "lbc_conv_packed_weight_bytes");
    auto wp = at::empty({(int64_t)wbytes}, kernel.options());
    lbc_throw(lbc_conv_prepack_weights(plan.p, (const int8_t*)wk.data_ptr(), LBC_W_KRSC, wp.data_ptr(), st),
              "lbc_conv_prepack_weights");
    int32_t p = 0, q = 0;
    lbc_throw(lbc_conv_out_shape(&d, &p, &q), "lbc_conv_out_shape");
    const auto odt = out_mode == LBC_OUT_INT32 ? at::kInt : at::kChar;
    auto y = at::empty({n, p, q, k}, input.options().dtype(odt));
    float ms = 0.f;
    lbc_throw(lbc_conv_run(plan.p, (const int8_t*)x.data_ptr(), wp.data_ptr(),
                           bias ? (const int32_t*)bias->data_ptr() : nullptr,
                           scale ? (const float*)scale->data_ptr() : nullptr, y.data_ptr(), st, &ms),
              "lbc_conv_run");
    auto out = at::empty({n, k / VECT_C, p, q, (int64_t)VECT_C}, y.options());
    lbc_throw(lbc_nhwc_to_vect_c(y.data_ptr(), out.data_ptr(), n, k, p, q, VECT_C, (int32_t)y.element_size(), st),
              "lbc_nhwc_to_vect_c");
    return {out, ms};
}
}  // namespace detail

template <const uint32_t batch, const uint32_t inC, const uint32_t outC, const uint32_t inH, const uint32_t inW,
          const uint32_t outH, const uint32_t outW>
std::tuple<at::Tensor, float> conv2DForward3x3(const at::Tensor& rinput, const at::Tensor& rkernel)
{
    static_assert(inC % VECT_C == 0 && outC % VECT_C == 0, "channels must be multiples of VECT_C");
    static_assert(outH == inH - 2 && outW == inW - 2, "the reference operator is 3x3, stride 1, no padding");
    TORCH_CHECK(rinput.size(0) == batch && rinput.size(1) * VECT_C == inC && rinput.size(2) == inH && rinput.size(3) == inW,
                "conv2DForward3x3: input does not match the template shape");
    TORCH_CHECK(rkernel.size(0) == outC, "conv2DForward3x3: kernel does not match the template shape");
    return detail::run3x3(rinput, rkernel, LBC_OUT_INT32, nullptr, nullptr, false);
}

// north_star extension: same tensors, fused bias + per-output-channel scale + RNE + ReLU/saturate -> int8 VECT_C
template <const uint32_t batch, const uint32_t inC, const uint32_t outC, const uint32_t inH, const uint32_t inW,
          const uint32_t outH, const uint32_t outW>
std::tuple<at::Tensor, float> conv2DForward3x3Requant(const at::Tensor& rinput, const at::Tensor& rkernel,
                                                      const at::Tensor& bias_i32, const at::Tensor& scale_f32, bool relu)
{
    TORCH_CHECK(bias_i32.numel() == outC && scale_f32.numel() == outC, "per-output-channel bias/scale expected");
    return detail::run3x3(rinput, rkernel, LBC_OUT_INT8, &bias_i32, &scale_f32, relu);
}

}  // namespace lowbit
