// igemm_tc.cu — tcgen05 implicit-GEMM int8 convolution for sm_100a.
//
//   GEMM view:  D[M = N*P*Q][K_out] = A[M][R*S*C] * B[K_out][R*S*C]^T     (int8 x int8 -> int32 in TMEM)
//   A tiles  :  TMA im2col-mode loads of the NHWC activation tensor (one filter tap x <=128 channels per
//               pipeline stage), or plain 2-D tiled loads when the conv is a pure GEMM (1x1, stride 1, no pad);
//               hardware zero-fill implements the padding halo and the M tail.
//   B tiles  :  TMA 2-D loads of the pre-packed [K_out][R][S][C_pad] filter matrix.
//   MMA      :  tcgen05.mma.cta_group::1.kind::i8, M=128 x N=bn x K=32 per instruction, issued by one thread,
//               operands straight from 128B/64B/32B-swizzled shared memory.
//   Epilogue :  4 warps drain the TMEM accumulator (tcgen05.ld 32x32b), fuse bias + per-channel fp32 scale +
//               round-to-nearest-even + ReLU/saturate, pack to int8 and write 16-byte vectors (NHWC).
//   Schedule :  persistent CTAs (one per SM), static round-robin over (m,n) tiles, a `stages`-deep smem ring
//               between the TMA warp and the MMA warp, and two TMEM accumulator stages so the epilogue of tile
//               i overlaps the main loop of tile i+1.
//
// Replaces CUDAConv2DForward3x3TensorCoures (cpp/int8conv/conv2DForward3x3TensorCores.cuh:537-693: wmma
// m32n8k16, single-buffered smem, int32 stores, 3x3/stride-1/VALID only).
#include "common.cuh"
#include "ptx.cuh"

#include <algorithm>
#include <mutex>

namespace lbc {

namespace {

constexpr int kBlockM = 128;
constexpr int kNumThreads = 192;          // warp 0: TMA, warp 1: MMA + TMEM alloc, warps 2..5: epilogue
constexpr int kEpilogueThreads = 128;
constexpr int kMaxStages = 8;

struct IgemmParams {
    int64_t m_total;
    int32_t k_out;
    int32_t bn, bkc, stages, a_im2col;
    int32_t tiles_m, tiles_n, k_blocks, cblocks;   // cblocks = c_pad / bkc
    int32_t p, q, s_taps;                          // output rows/cols, filter width
    int32_t stride_h, stride_w, pad_h, pad_w, dil_h, dil_w;
    int32_t relu, out_mode;
    uint32_t tmem_cols;
    uint32_t a_stage_bytes, b_stage_bytes;
};

__device__ int g_timeout_flag = 0;

struct SmemLayout {
    // dynamic smem: [stages x A tile][stages x B tile] (1024-aligned), then this control block.
    uint64_t full[kMaxStages];
    uint64_t empty[kMaxStages];
    uint64_t tmem_full[2];
    uint64_t tmem_empty[2];
    uint32_t tmem_base;
    uint32_t pad_;
};

// Requantise / bias NCOLS consecutive accumulator columns of one output pixel and store them with 16-byte
// vectors.  bias/scale are read straight from global memory: every lane of the warp reads the same address
// (one broadcast transaction, L1-resident), which needs no cross-warp staging barrier.
template <int NCOLS>
__device__ __forceinline__ void epilogue_store_chunk(const uint32_t* v, const float* __restrict__ sc,
                                                     const int32_t* __restrict__ bi, float lo, int32_t out_mode,
                                                     void* y, int64_t out_off)
{
    if (out_mode == LBC_OUT_INT32) {
        int32_t* yo = reinterpret_cast<int32_t*>(y) + out_off;
#pragma unroll
        for (int j = 0; j < NCOLS; j += 4) {
            const int4 b = bi ? __ldg(reinterpret_cast<const int4*>(bi + j)) : make_int4(0, 0, 0, 0);
            ptx::st_global_v4(yo + j, v[j] + (uint32_t)b.x, v[j + 1] + (uint32_t)b.y, v[j + 2] + (uint32_t)b.z,
                              v[j + 3] + (uint32_t)b.w);
        }
    } else {
        int8_t* yo = reinterpret_cast<int8_t*>(y) + out_off;
#pragma unroll
        for (int j = 0; j < NCOLS; j += 16) {
            uint32_t w[4];
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const int c = j + 4 * t;
                const int4 b = bi ? __ldg(reinterpret_cast<const int4*>(bi + c)) : make_int4(0, 0, 0, 0);
                const float4 f = __ldg(reinterpret_cast<const float4*>(sc + c));
                const uint32_t b0 = requant_u8bits((int32_t)v[c + 0], b.x, f.x, lo);
                const uint32_t b1 = requant_u8bits((int32_t)v[c + 1], b.y, f.y, lo);
                const uint32_t b2 = requant_u8bits((int32_t)v[c + 2], b.z, f.z, lo);
                const uint32_t b3 = requant_u8bits((int32_t)v[c + 3], b.w, f.w, lo);
                w[t] = pack4_u8(b0, b1, b2, b3);
            }
            ptx::st_global_v4(yo + j, w[0], w[1], w[2], w[3]);
        }
    }
}

__global__ void __launch_bounds__(kNumThreads, 1)
igemm_i8_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                const IgemmParams prm, const int32_t* __restrict__ bias, const float* __restrict__ scale,
                void* __restrict__ y)
{
    extern __shared__ uint8_t smem_raw[];
    // 1024-byte alignment for the 128B-swizzle atoms (the dynamic smem base is only 16B-aligned by contract).
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + (size_t)prm.stages * prm.a_stage_bytes;
    SmemLayout* ctl = reinterpret_cast<SmemLayout*>(smem_b + (size_t)prm.stages * prm.b_stage_bytes);

    const uint32_t warp = threadIdx.x >> 5;
    const uint32_t lane = threadIdx.x & 31;
    volatile int* tflag = &g_timeout_flag;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tensormap(&tm_a);
        ptx::prefetch_tensormap(&tm_b);
        for (int i = 0; i < prm.stages; ++i) {
            ptx::mbar_init(&ctl->full[i], 1);
            ptx::mbar_init(&ctl->empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            ptx::mbar_init(&ctl->tmem_full[i], 1);
            ptx::mbar_init(&ctl->tmem_empty[i], kEpilogueThreads / 32);
        }
        ptx::fence_barrier_init();
    }
    if (warp == 1) {
        ptx::tmem_alloc(&ctl->tmem_base, prm.tmem_cols);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = ctl->tmem_base;

    const int32_t num_tiles = prm.tiles_m * prm.tiles_n;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            bool ok = true;
            const uint32_t tx_bytes = prm.a_stage_bytes + prm.b_stage_bytes;
            for (int32_t tile = blockIdx.x; tile < num_tiles && ok; tile += gridDim.x) {
                const int32_t n_blk = tile % prm.tiles_n;
                const int32_t m_blk = tile / prm.tiles_n;
                const int64_t m0 = (int64_t)m_blk * kBlockM;
                // base pixel of the tile in input coordinates (im2col mode)
                const int32_t q0 = (int32_t)(m0 % prm.q);
                const int32_t p0 = (int32_t)((m0 / prm.q) % prm.p);
                const int32_t n0 = (int32_t)(m0 / ((int64_t)prm.q * prm.p));
                const int32_t w_base = q0 * prm.stride_w - prm.pad_w;
                const int32_t h_base = p0 * prm.stride_h - prm.pad_h;
                for (int32_t kb = 0; kb < prm.k_blocks; ++kb) {
                    ok = ptx::mbar_wait(&ctl->empty[stage], phase ^ 1, tflag);
                    if (!ok) break;
                    ptx::mbar_expect_tx(&ctl->full[stage], tx_bytes);
                    const int32_t tap = kb / prm.cblocks;
                    const int32_t c0 = (kb - tap * prm.cblocks) * prm.bkc;
                    uint8_t* dst_a = smem_a + (size_t)stage * prm.a_stage_bytes;
                    uint8_t* dst_b = smem_b + (size_t)stage * prm.b_stage_bytes;
                    if (prm.a_im2col) {
                        const int32_t fr = tap / prm.s_taps;
                        const int32_t fs = tap - fr * prm.s_taps;
                        ptx::tma_load_im2col_4d(dst_a, &tm_a, &ctl->full[stage], c0, w_base, h_base, n0,
                                                (uint16_t)(fs * prm.dil_w), (uint16_t)(fr * prm.dil_h));
                    } else {
                        ptx::tma_load_2d(dst_a, &tm_a, &ctl->full[stage], c0, (int32_t)m0);
                    }
                    ptx::tma_load_2d(dst_b, &tm_b, &ctl->full[stage], kb * prm.bkc, n_blk * prm.bn);
                    if (++stage == (uint32_t)prm.stages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            uint32_t acc_stage = 0, acc_phase = 0;
            bool ok = true;
            const uint32_t idesc = ptx::make_idesc_i8(kBlockM, (uint32_t)prm.bn);
            const uint32_t k_steps = (uint32_t)prm.bkc / 32;
            for (int32_t tile = blockIdx.x; tile < num_tiles && ok; tile += gridDim.x) {
                ok = ptx::mbar_wait(&ctl->tmem_empty[acc_stage], acc_phase ^ 1, tflag);
                if (!ok) break;
                ptx::tc_fence_after();
                const uint32_t tmem_d = tmem_base + acc_stage * (uint32_t)prm.bn;
                for (int32_t kb = 0; kb < prm.k_blocks; ++kb) {
                    ok = ptx::mbar_wait(&ctl->full[stage], phase, tflag);
                    if (!ok) break;
                    ptx::tc_fence_after();
                    const uint32_t a_addr = ptx::smem_u32(smem_a + (size_t)stage * prm.a_stage_bytes);
                    const uint32_t b_addr = ptx::smem_u32(smem_b + (size_t)stage * prm.b_stage_bytes);
                    const uint64_t da = ptx::make_kmajor_desc(a_addr, (uint32_t)prm.bkc);
                    const uint64_t db = ptx::make_kmajor_desc(b_addr, (uint32_t)prm.bkc);
                    for (uint32_t k = 0; k < k_steps; ++k) {
                        // advance 32 bytes along K inside the swizzle span: +2 in the (addr >> 4) field
                        ptx::mma_i8_ss(tmem_d, da + 2ull * k, db + 2ull * k, idesc, (kb > 0 || k > 0) ? 1u : 0u);
                    }
                    ptx::mma_commit(&ctl->empty[stage]);      // smem slot reusable once these MMAs retire
                    if (++stage == (uint32_t)prm.stages) { stage = 0; phase ^= 1; }
                }
                if (!ok) break;
                ptx::mma_commit(&ctl->tmem_full[acc_stage]);  // accumulator complete -> epilogue
                acc_stage ^= 1;
                if (acc_stage == 0) acc_phase ^= 1;
            }
        }
    } else {
        // ===================== epilogue (warps 2..5) =====================
        const uint32_t quarter = warp & 3;                    // TMEM lanes [32*quarter, 32*quarter+32)
        const float lo = prm.relu ? 0.0f : -128.0f;
        uint32_t acc_stage = 0, acc_phase = 0;
        bool ok = true;
        for (int32_t tile = blockIdx.x; tile < num_tiles && ok; tile += gridDim.x) {
            const int32_t n_blk = tile % prm.tiles_n;
            const int32_t m_blk = tile / prm.tiles_n;
            const int32_t col0 = n_blk * prm.bn;
            ok = ptx::mbar_wait(&ctl->tmem_full[acc_stage], acc_phase, tflag);
            if (!ok) break;
            ptx::tc_fence_after();

            const int64_t row = (int64_t)m_blk * kBlockM + quarter * 32 + lane;
            const bool row_ok = row < prm.m_total;
            const uint32_t taddr = tmem_base + ((quarter * 32u) << 16) + acc_stage * (uint32_t)prm.bn;
            const float* sc = scale ? scale + col0 : nullptr;
            const int32_t* bi = bias ? bias + col0 : nullptr;
            int32_t c = 0;
            for (; c + 32 <= prm.bn; c += 32) {
                uint32_t v[32];
                ptx::tmem_ld_32x32b_x32(taddr + (uint32_t)c, v);
                ptx::tmem_ld_wait();
                if (row_ok) {
                    const int64_t off = row * prm.k_out + col0 + c;
                    if (col0 + c + 32 <= prm.k_out)
                        epilogue_store_chunk<32>(v, sc + c, bi ? bi + c : nullptr, lo, prm.out_mode, y, off);
                    else if (col0 + c < prm.k_out)      // N tail: K_out % 16 == 0, so exactly 16 valid columns
                        epilogue_store_chunk<16>(v, sc + c, bi ? bi + c : nullptr, lo, prm.out_mode, y, off);
                }
            }
            if (c < prm.bn) {   // bn % 32 == 16
                uint32_t v[16];
                ptx::tmem_ld_32x32b_x16(taddr + (uint32_t)c, v);
                ptx::tmem_ld_wait();
                if (row_ok && col0 + c < prm.k_out)
                    epilogue_store_chunk<16>(v, sc + c, bi ? bi + c : nullptr, lo, prm.out_mode, y,
                                             row * prm.k_out + col0 + c);
            }
            // accumulator drained: hand the TMEM stage back to the MMA warp
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&ctl->tmem_empty[acc_stage]);
            acc_stage ^= 1;
            if (acc_stage == 0) acc_phase ^= 1;
        }
    }

    // ---- teardown ----
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_base, prm.tmem_cols);
    }
}

// ---- host side ---------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
typedef CUresult (*PFN_encodeIm2col)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                     const cuuint64_t*, const int*, const int*, cuuint32_t, cuuint32_t,
                                     const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                     CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

PFN_encodeTiled g_encode_tiled = nullptr;
PFN_encodeIm2col g_encode_im2col = nullptr;
std::once_flag g_entry_once;

// The driver entry points are resolved at run time so the library links (and loads on a CPU-only box)
// without libcuda.so.
lbc_status resolve_driver_entry_points()
{
    std::call_once(g_entry_once, [] {
        cudaDriverEntryPointQueryResult qres;
        void* fn = nullptr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            g_encode_tiled = reinterpret_cast<PFN_encodeTiled>(fn);
        fn = nullptr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &fn, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            g_encode_im2col = reinterpret_cast<PFN_encodeIm2col>(fn);
    });
    LBC_REQUIRE(g_encode_tiled && g_encode_im2col, LBC_ERR_CUDA, "cuTensorMapEncode* driver entry points unavailable");
    return LBC_OK;
}

CUtensorMapSwizzle swizzle_for(int bkc)
{
    return bkc == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : bkc == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
}

// Driver quirk handled the way CUTLASS does it (cute/atom/copy_traits_sm90_im2col.hpp, "driver_version <=
// 13010"): for tensors smaller than 128 KiB a descriptor bit must be cleared or the copy misbehaves.
void small_tensor_fixup(CUtensorMap* tm, size_t tensor_bytes, int driver_version)
{
    if (driver_version <= 13010 && tensor_bytes < 131072)
        reinterpret_cast<uint64_t*>(tm)[1] &= ~(1ull << 21);
}

bool g_attr_set = false;
std::mutex g_attr_mu;

}  // namespace

bool igemm_supported(const ConvGeom& g, std::string* why)
{
    const lbc_conv_desc& d = g.d;
    auto no = [&](const char* m) { if (why) *why = m; return false; };
    if (d.groups != 1) return no("groups != 1");
    if (d.c % 16 != 0) return no("C % 16 != 0 (TMA global strides must be multiples of 16 bytes)");
    if (d.k % 16 != 0) return no("K % 16 != 0 (16-byte output vectors)");
    if (d.stride_h > 8 || d.stride_w > 8) return no("stride > 8 (TMA traversal stride)");
    if (d.pad_h > 127 || d.pad_w > 127) return no("padding beyond the im2col corner range");
    if ((d.r - 1) * d.dil_h > 127 + d.pad_h || (d.s - 1) * d.dil_w > 127 + d.pad_w) return no("filter extent beyond the im2col corner range");
    if (g.m_total >= (1ll << 31)) return no("N*P*Q >= 2^31");
    return true;
}

lbc_status igemm_make_config(const ConvGeom& g, const DeviceInfo& dev, IgemmConfig* cfg)
{
    const lbc_conv_desc& d = g.d;
    IgemmConfig c{};
    c.bkc = (d.c % 128 == 0) ? 128 : (d.c % 64 == 0) ? 64 : 32;
    c.c_pad = (d.c + c.bkc - 1) / c.bkc * c.bkc;
    // N tile: the whole K_out when it fits one 256-wide tile, else the largest of 256/128/... dividing it.
    if (d.k <= 256) c.bn = d.k;
    else if (d.k % 256 == 0) c.bn = 256;
    else if (d.k % 128 == 0) c.bn = 128;
    else c.bn = 256;   // tail tile handled by TMA zero-fill + masked stores
    c.tiles_n = (d.k + c.bn - 1) / c.bn;
    c.tiles_m = (int32_t)((g.m_total + kBlockM - 1) / kBlockM);
    c.k_blocks = d.r * d.s * (c.c_pad / c.bkc);
    c.a_im2col = !(d.r == 1 && d.s == 1 && d.stride_h == 1 && d.stride_w == 1 && d.pad_h == 0 && d.pad_w == 0);
    if (getenv("LBC_FORCE_IM2COL")) c.a_im2col = 1;   // debugging aid: exercise the im2col path on 1x1 layers
    const size_t stage_bytes = (size_t)(kBlockM + c.bn) * c.bkc;
    const size_t budget = 227 * 1024 - 1024 /*alignment slack*/ - sizeof(SmemLayout);
    int stages = (int)std::min<size_t>(kMaxStages, budget / stage_bytes);
    stages = std::max(2, std::min(stages, std::max(2, c.k_blocks * 2)));
    c.stages = stages;
    c.smem_bytes = 1024 + (size_t)stages * stage_bytes + sizeof(SmemLayout);
    uint32_t cols = 32;
    while (cols < 2u * (uint32_t)c.bn) cols <<= 1;
    c.tmem_cols = cols;
    c.grid = std::min(dev.sm_count > 0 ? dev.sm_count : 148, c.tiles_m * c.tiles_n);
    LBC_REQUIRE(c.smem_bytes <= 227 * 1024, LBC_ERR_UNSUPPORTED, "igemm: smem %zu too large", c.smem_bytes);
    *cfg = c;
    return LBC_OK;
}

lbc_status igemm_encode(const ConvGeom& g, const IgemmConfig& cfg, const DeviceInfo& dev, const int8_t* x,
                        const int8_t* w_packed, IgemmLaunch* out)
{
    lbc_status st = resolve_driver_entry_points();
    if (st != LBC_OK) return st;
    const lbc_conv_desc& d = g.d;
    LBC_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(w_packed) & 15) == 0,
                LBC_ERR_INVALID_ARG, "igemm: x and packed weights must be 16-byte aligned");
    out->cfg = cfg;
    const CUtensorMapSwizzle swz = swizzle_for(cfg.bkc);
    const cuuint32_t ones[5] = {1, 1, 1, 1, 1};

    // ---- B: [K_out rows][R*S*c_pad bytes], box {bkc, bn}
    {
        const cuuint64_t dims[2] = {(cuuint64_t)d.r * d.s * cfg.c_pad, (cuuint64_t)d.k};
        const cuuint64_t strides[1] = {(cuuint64_t)d.r * d.s * cfg.c_pad};
        const cuuint32_t box[2] = {(cuuint32_t)cfg.bkc, (cuuint32_t)cfg.bn};
        CUresult r = g_encode_tiled(&out->tm_b, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, (void*)w_packed, dims, strides, box,
                                    ones, CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        LBC_REQUIRE(r == CUDA_SUCCESS, LBC_ERR_CUDA, "cuTensorMapEncodeTiled(B) failed: %d", (int)r);
    }
    // ---- A
    if (!cfg.a_im2col) {
        const cuuint64_t dims[2] = {(cuuint64_t)d.c, (cuuint64_t)g.m_total};
        const cuuint64_t strides[1] = {(cuuint64_t)d.c};
        const cuuint32_t box[2] = {(cuuint32_t)cfg.bkc, (cuuint32_t)kBlockM};
        CUresult r = g_encode_tiled(&out->tm_a, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, (void*)x, dims, strides, box, ones,
                                    CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        LBC_REQUIRE(r == CUDA_SUCCESS, LBC_ERR_CUDA, "cuTensorMapEncodeTiled(A) failed: %d", (int)r);
    } else {
        // rank-4 (C, W, H, N); corners in (W, H) order.  lower = -pad ; upper = pad - (filter-1)*dilation.
        const cuuint64_t dims[4] = {(cuuint64_t)d.c, (cuuint64_t)d.w, (cuuint64_t)d.h, (cuuint64_t)d.n};
        const cuuint64_t strides[3] = {(cuuint64_t)d.c, (cuuint64_t)d.c * d.w, (cuuint64_t)d.c * d.w * d.h};
        const int lower[2] = {-d.pad_w, -d.pad_h};
        const int upper[2] = {d.pad_w - (d.s - 1) * d.dil_w, d.pad_h - (d.r - 1) * d.dil_h};
        const cuuint32_t trav[4] = {1, (cuuint32_t)d.stride_w, (cuuint32_t)d.stride_h, 1};
        CUresult r = g_encode_im2col(&out->tm_a, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, (void*)x, dims, strides, lower,
                                     upper, (cuuint32_t)cfg.bkc, (cuuint32_t)kBlockM, trav,
                                     CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        LBC_REQUIRE(r == CUDA_SUCCESS, LBC_ERR_CUDA, "cuTensorMapEncodeIm2col(A) failed: %d", (int)r);
        small_tensor_fixup(&out->tm_a, (size_t)d.n * d.h * d.w * d.c, dev.driver_version);
    }
    return LBC_OK;
}

lbc_status igemm_launch(const ConvGeom& g, const IgemmLaunch& l, const EpilogueParams& ep, void* y,
                        cudaStream_t stream)
{
    const IgemmConfig& c = l.cfg;
    const lbc_conv_desc& d = g.d;
    IgemmParams prm{};
    prm.m_total = g.m_total;
    prm.k_out = d.k;
    prm.bn = c.bn; prm.bkc = c.bkc; prm.stages = c.stages; prm.a_im2col = c.a_im2col;
    prm.tiles_m = c.tiles_m; prm.tiles_n = c.tiles_n; prm.k_blocks = c.k_blocks; prm.cblocks = c.c_pad / c.bkc;
    prm.p = g.p; prm.q = g.q; prm.s_taps = d.s;
    prm.stride_h = d.stride_h; prm.stride_w = d.stride_w; prm.pad_h = d.pad_h; prm.pad_w = d.pad_w;
    prm.dil_h = d.dil_h; prm.dil_w = d.dil_w;
    prm.relu = ep.relu; prm.out_mode = ep.out_mode;
    prm.tmem_cols = c.tmem_cols;
    prm.a_stage_bytes = (uint32_t)(kBlockM * c.bkc);
    prm.b_stage_bytes = (uint32_t)(c.bn * c.bkc);
    LBC_REQUIRE((reinterpret_cast<uintptr_t>(y) & 15) == 0, LBC_ERR_INVALID_ARG, "igemm: y must be 16-byte aligned");
    {
        std::lock_guard<std::mutex> lk(g_attr_mu);
        if (!g_attr_set) {
            LBC_CUDA_TRY(cudaFuncSetAttribute(igemm_i8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
            g_attr_set = true;
        }
    }
    igemm_i8_kernel<<<c.grid, kNumThreads, c.smem_bytes, stream>>>(l.tm_a, l.tm_b, prm, ep.bias, ep.scale, y);
    LBC_CUDA_TRY(cudaGetLastError());
    return LBC_OK;
}

lbc_status igemm_check_timeout()
{
    int flag = 0;
    LBC_CUDA_TRY(cudaMemcpyFromSymbol(&flag, g_timeout_flag, sizeof(int)));
    if (flag) {
        int zero = 0;
        cudaMemcpyToSymbol(g_timeout_flag, &zero, sizeof(int));
        set_error("igemm: device pipeline watchdog fired (mbarrier wait exceeded 2 s)");
        return LBC_ERR_KERNEL_TIMEOUT;
    }
    return LBC_OK;
}

}  // namespace lbc
