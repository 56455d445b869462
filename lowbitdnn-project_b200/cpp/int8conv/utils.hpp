// utils.hpp — reference-signature layout helpers on at::Tensor, backed by liblowbit-cnn.
// Mirrors cpp/int8conv/utils.cuh:8-26 of the reference (same names, same argument meaning).  The reference
// returns lazy views and pays a hidden .contiguous() transpose inside every operator
// (conv2DForward3x3TensorCores.cuh:715-716); these materialise the converted tensor with one kernel.
#pragma once
#include <ATen/ATen.h>
#include <ATen/cuda/CUDAContext.h>
#include <stdexcept>
#include <string>

#include "../../../include/lowbit_cnn.h"

namespace lowbit {

constexpr uint32_t VECT_C = 16;

inline void lbc_throw(lbc_status st, const char* what)
{
    if (st != LBC_OK) throw std::runtime_error(std::string(what) + ": " + lbc_last_error_string());
}

inline lbc_stream current_stream() { return (lbc_stream)at::cuda::getCurrentCUDAStream().stream(); }

// [N,C,H,W] -> [N,C/V,H,W,V]
inline at::Tensor to_vect_c(const at::Tensor& tensor, const uint32_t V = VECT_C)
{
    auto t = tensor.contiguous();
    TORCH_CHECK(t.is_cuda() && t.dim() == 4 && t.size(1) % V == 0, "to_vect_c: need a CUDA NCHW tensor with C % V == 0");
    TORCH_CHECK(t.element_size() == 1 || t.element_size() == 4, "to_vect_c: 1- or 4-byte elements");
    auto out = at::empty({t.size(0), t.size(1) / V, t.size(2), t.size(3), (int64_t)V}, t.options());
    lbc_throw(lbc_to_vect_c(t.data_ptr(), out.data_ptr(), (int32_t)t.size(0), (int32_t)t.size(1), (int32_t)t.size(2),
                            (int32_t)t.size(3), (int32_t)V, (int32_t)t.element_size(), current_stream()),
              "lbc_to_vect_c");
    return out;
}

// [N,C/V,H,W,V] -> [N,C,H,W]
inline at::Tensor from_vect_c(const at::Tensor& tensor)
{
    auto t = tensor.contiguous();
    TORCH_CHECK(t.is_cuda() && t.dim() == 5, "from_vect_c: need a CUDA [N,C/V,H,W,V] tensor");
    const int64_t V = t.size(4), C = t.size(1) * V;
    auto out = at::empty({t.size(0), C, t.size(2), t.size(3)}, t.options());
    lbc_throw(lbc_from_vect_c(t.data_ptr(), out.data_ptr(), (int32_t)t.size(0), (int32_t)C, (int32_t)t.size(2),
                              (int32_t)t.size(3), (int32_t)V, (int32_t)t.element_size(), current_stream()),
              "lbc_from_vect_c");
    return out;
}

}  // namespace lowbit
