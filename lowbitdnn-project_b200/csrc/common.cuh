// common.cuh — internal declarations shared by the liblowbit-cnn translation units.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "../../include/lowbit_cnn.h"

namespace lbc {

// ---- error plumbing (never throws across the C boundary) ---------------------------------------
void set_error(const char* fmt, ...);
const char* last_error();

#define LBC_CUDA_TRY(expr)                                                                      \
    do {                                                                                        \
        cudaError_t _e = (expr);                                                                \
        if (_e != cudaSuccess) {                                                                \
            ::lbc::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return LBC_ERR_CUDA;                                                                \
        }                                                                                       \
    } while (0)

#define LBC_REQUIRE(cond, status, ...)        \
    do {                                      \
        if (!(cond)) {                        \
            ::lbc::set_error(__VA_ARGS__);    \
            return (status);                  \
        }                                     \
    } while (0)

// ---- geometry ----------------------------------------------------------------------------------
struct ConvGeom {
    lbc_conv_desc d;
    int32_t p, q;        // output spatial size
    int32_t cg, kg;      // channels / filters per group
    int64_t m_total;     // N*P*Q
};

lbc_status make_geom(const lbc_conv_desc* d, ConvGeom* g);

// ---- device info (cached) ----------------------------------------------------------------------
struct DeviceInfo {
    int device = -1;
    int sm_count = 0;
    int cc_major = 0, cc_minor = 0;
    size_t hbm_bytes = 0;
    int driver_version = 0;
};
lbc_status current_device(DeviceInfo* info);  // LBC_ERR_NO_DEVICE unless cc 10.x

// ---- kernels (one TU each) ---------------------------------------------------------------------
struct EpilogueParams {
    const int32_t* bias;   // may be null
    const float* scale;    // null only for int32 output
    int32_t relu;
    int32_t out_mode;      // lbc_out_mode
};

// direct_conv.cu — CUDA-core direct convolution, any shape; weights [K][R][S][C/g].
lbc_status launch_direct_conv(const ConvGeom& g, const int8_t* x, const int8_t* w_krsc, const EpilogueParams& ep,
                              void* y, cudaStream_t stream);

// depthwise.cu — groups == C == K; weights packed [R][S][C].
// The 3x3 kernel stages input tiles in shared memory with TMA when the shape allows it (DwLaunch::tiled); the launch
// descriptor (tensor map over x) is built once per (plan, x) by depthwise_encode.
struct DwLaunch {
    int32_t tiled = 0;           // 1: TMA-staged tile kernel, 0: direct global-memory kernel
    CUtensorMap tm_x;
    int32_t cc, th, tq, nb, tw, in_h, in_w, tiles_c, tiles_q, tiles_p, tiles_n;
    uint32_t tile_bytes;
    int32_t reverse = 0;         // tiled kernel: walk the tiles last-to-first
};
lbc_status depthwise_encode(const ConvGeom& g, const lbc_plan_options& opt, const int8_t* x, DwLaunch* out);
// `flag`: the plan's device watchdog word (see IgemmRuntime)
lbc_status launch_depthwise(const ConvGeom& g, const int8_t* x, const int8_t* w_rsc, const EpilogueParams& ep,
                            void* y, const DwLaunch* dw, int* flag, cudaStream_t stream);
lbc_status encode_tiled_u8_4d(CUtensorMap* tm, const void* base, const uint64_t dims[4], const uint32_t box[4]);

// igemm_tc.cu — tcgen05 implicit GEMM.
struct IgemmConfig {
    int32_t mode;        // 0 tiled (pure GEMM), 1 TMA im2col, 2 shifted window
    int32_t bn;          // N tile (multiple of 16, <= 256)
    int32_t bkc;         // bytes of K per pixel row of an A block: 16 (window, C == 16) / 32 / 64 / 128
    int32_t bkb;         // bytes of K per row of a B block
    int32_t c_pad;       // C rounded up to a multiple of bkc
    int32_t s_pad;       // filter width rounded up to even when taps are consumed in pairs (C == 16)
    int32_t cblocks, inner, k_blocks;
    size_t packed_row_bytes;   // bytes per output channel in the packed filter matrix
    int32_t stages, win_stages, tps;
    uint32_t a_block_bytes, b_block_bytes;
    uint32_t a_stage_bytes, b_stage_bytes, win_stage_bytes, win_tx_bytes;
    int32_t wt, rows_per_tile, cols_per_tile, row_tiles, col_tiles;
    int32_t tiles_m, tiles_n;
    int32_t panel_bytes, panel_swz_bits, n_panels;
    int32_t stage_bufs;  // staging panels per epilogue team
    int32_t warp_store;  // ring modes: every epilogue warp stages and TMA-stores its own 32 rows (no team barrier per panel)
    int32_t team_warps;  // 8 (two epilogue teams) or 4 (four teams, N tile <= 64)
    int32_t epi_split;   // both epilogue teams drain every tile (panels / column halves) instead of alternate tiles
    int32_t k_mod;       // bias/scale index modulo (pixel-group rewrite), 0 = none
    int32_t n_tab;       // MMA issue table: A-descriptor offsets (16-byte units) of one channel chunk
    uint16_t a_tab[192];
    uint16_t b_tab[192]; // resident-B window mode: matching B-descriptor offsets
    int32_t res_b;       // filter matrix resident in shared memory (loaded once per CTA)
    int32_t res_one;     // ... only the CTA's own N tile of it (N-stationary schedule, grid % tiles_n == 0)
    int32_t n_mma;       // MMA-issuing warps (1 or 2)
    int32_t pair;        // two M tiles per CTA step share every B block (window A, streaming B)
    int32_t it_imgs;     // pair modes: image (ring modes: M tile) radix of the tile numbering, padded so tiles come in pairs
    int32_t fold;        // bias folded into the first MMA of every tile (resident filter matrix); off_fold = its operand blocks
    uint32_t off_fold;
    int32_t tpi;         // tiles per epilogue-team iteration (2: narrow N tiles with 8 accumulator stages)
    int32_t cta2;        // CTA-pair mode: clusters of two CTAs, cta_group::2 MMAs, half of the B rows per CTA
    uint32_t win_sub_bytes;   // bytes of one window (a pair-mode stage holds two)
    uint32_t b_total_bytes;
    uint32_t off_b, off_stage, off_ctl;
    int32_t grid;        // persistent CTAs
    size_t smem_bytes;
    uint32_t tmem_cols;  // power of two >= n_acc*bn
    int32_t n_acc;       // TMEM accumulator stages (2, 4 or 8)
    int32_t tail_first, tail_count, tail_m0;   // split last round of a CTA-pair launch (IgemmParams::tail_first), -1 = off
    int32_t reverse;     // lbc_plan_options::reverse: walk the tiles last-to-first
    int32_t pdl;         // programmatic dependent launch allowed
};
// per-launch state that belongs to the plan, not to the tiling
struct IgemmRuntime {
    int* flag = nullptr;            // device watchdog word of the plan (never null on a launch)
    long long* trace = nullptr;     // optional pipeline trace buffer (development aid)
    int32_t trace_tiles = 0;
};
struct IgemmLaunch {
    CUtensorMap tm_a;
    CUtensorMap tm_b;
    CUtensorMap tm_out;
    CUtensorMap tm_out2;   // window tiles with per-warp stores: the ragged last run of a tile row (else a copy of tm_out)
    IgemmConfig cfg;
    int32_t reverse = 0;   // walk the M tiles / images last-to-first (L2 reuse of the producer's most recent output)
    int32_t early_b = 0;   // the packed weights are not written by anything earlier in the stream: a resident filter matrix
                           // may be fetched before the programmatic-dependency wait (set by the network runner)
};
bool igemm_supported(const ConvGeom& g, std::string* why);
bool igemm_trace_compiled();      // the library was built with -DLBC_TRACE=1 (lib/liblowbit_cnn_trace.so)
lbc_status igemm_make_config(const ConvGeom& g, const DeviceInfo& dev, const lbc_plan_options& opt, IgemmConfig* cfg);
lbc_status igemm_encode(const ConvGeom& g, const IgemmConfig& cfg, const DeviceInfo& dev, const int8_t* x,
                        const int8_t* w_packed, void* y, IgemmLaunch* out);
lbc_status igemm_launch(const ConvGeom& g, const IgemmLaunch& l, const EpilogueParams& ep, void* y,
                        const IgemmRuntime& rt, cudaStream_t stream);
// Fused bottleneck tail (igemm_fused_tail_kernel): conv A (R x S, stride 1, 64 output channels, window mode with a resident
// filter) -> int8 -> conv B (1x1, 64 -> 256) in one launch.  `cfg_a` is conv A's own plan config; the fused launch re-carves
// shared memory around it.
struct FusedLaunch {
    CUtensorMap tm_a, tm_b1, tm_b2, tm_out;
    IgemmConfig cfg_a;           // window geometry, issue tables, K chunking of conv A (ring depths recomputed for the fused carve-up)
    int32_t k2, relu2, stage_bufs2, win_stages;
    uint32_t off_b1, off_b2, off_a2, off_stage2, off_ctl2, b1_bytes, b2_bytes;
    size_t smem_bytes;
    int32_t grid;
    int32_t reverse = 0;
};
bool fused_tail_supported(const ConvGeom& ga, const IgemmConfig& cfg_a, const ConvGeom& gb, const IgemmConfig& cfg_b, std::string* why);
lbc_status fused_tail_encode(const ConvGeom& ga, const IgemmConfig& cfg_a, const ConvGeom& gb, const DeviceInfo& dev, const int8_t* x,
                             const int8_t* wa_packed, const int8_t* wb_packed, void* y, FusedLaunch* out);
lbc_status fused_tail_launch(const ConvGeom& ga, const ConvGeom& gb, const FusedLaunch& l, const EpilogueParams& epa,
                             const EpilogueParams& epb, const IgemmRuntime& rt, cudaStream_t stream);

// per-device one-time kernel attributes: true the first time it is called for `device` under `mask`
bool first_use_on_device(uint64_t (&mask)[4], int device);

// layout.cu
lbc_status launch_prepack_krsc(const int8_t* src, int32_t src_layout, int8_t* dst, int32_t k, int32_t r, int32_t s,
                               int32_t cg, int32_t c_pad, cudaStream_t stream);              // -> [K][R][S][c_pad]
// -> the tcgen05 kernel's filter matrix: per output channel `row_bytes`, K ordered [tap][chunk][bkc] (ring modes)
// or [chunk][tap][bkc] (window mode); taps = R x s_pad with zero phantom taps / zero channel padding.
lbc_status launch_prepack_igemm(const int8_t* src, int32_t src_layout, int8_t* dst, int32_t k, int32_t r, int32_t s,
                                int32_t cg, int32_t s_pad, int32_t bkc, int32_t cblocks, int32_t chunk_outer,
                                cudaStream_t stream);
// small-C ("stem") rewrite: zero-pad + space-to-depth into 16-channel pixels, and the matching filter matrix
lbc_status launch_stem_xform(const int8_t* x, void* out, int32_t n, int32_t h, int32_t w, int32_t c, int32_t hs,
                             int32_t ws, int32_t sh, int32_t sw, int32_t pad_h, int32_t pad_w, cudaStream_t stream);
lbc_status launch_prepack_stem(const int8_t* src, int32_t src_layout, int8_t* dst, int32_t k, int32_t r, int32_t s,
                               int32_t c, int32_t sh, int32_t sw, int32_t r2, int32_t s_pad, cudaStream_t stream);
// pixel-group rewrite of pointwise layers: [K][C] -> block-diagonal [f*K][f*C]
lbc_status launch_blockdiag(const int8_t* src, int8_t* dst, int32_t k, int32_t c, int32_t f, cudaStream_t stream);
lbc_status launch_prepack_depthwise(const int8_t* src, int32_t src_layout, int8_t* dst, int32_t c, int32_t r,
                                    int32_t s, cudaStream_t stream);                         // -> [R][S][C]
lbc_status launch_dgrad_weights(const int8_t* src, int8_t* dst, int32_t k, int32_t r, int32_t s, int32_t c, cudaStream_t stream);
lbc_status launch_permute5(const void* src, void* dst, const int32_t dims[5], const int32_t perm[5], int32_t elt,
                           cudaStream_t stream);

// pool_add.cu — int8 ops between convolutions
lbc_status launch_maxpool(const lbc_pool_desc& d, int32_t p, int32_t q, const int8_t* x, int8_t* y, int sm_count, cudaStream_t stream);
lbc_status launch_add_relu(const int8_t* a, const int8_t* b, int8_t* y, size_t n, int32_t relu, int sm_count, cudaStream_t stream);
lbc_status launch_global_avgpool(const int8_t* x, int8_t* y, int32_t n, int32_t hw, int32_t c, float scale, cudaStream_t stream);

// probes.cu
lbc_status probe_int8_mma_peak(int32_t iters, double* tops, cudaStream_t stream);
lbc_status probe_hbm_copy(size_t bytes, int32_t iters, double* gbs, cudaStream_t stream);
lbc_status flush_l2(cudaStream_t stream);

// ---- the fused epilogue, shared by every kernel (bit-exact with oracle_requant) ----------------
// The reference's quantize() (cpp/int8conv/conv2DForward3x3WinogradFused.cuh:39-46): round FIRST, then clamp.
// t = acc + bias (int32 wraparound) ; f = float(t) RNE ; f *= scale (single multiply) ; q = cvt.rni.s32.f32(f)
// (round-half-to-even, saturating, NaN -> 0) ; q = max(q, lo) with lo = 0 (ReLU) or -128 ; the upper clamp to 127 is
// done by the saturating int32->int8 pack.  Rounding then clamping at an integer bound equals clamping then rounding
// for every finite value; the order only matters for NaN (-> 0, as in the reference).
// Measured on B200 (tools/exp/epi_bench.cu): this F2I + cvt.pack.sat form is ~25% faster than a float clamp +
// magic-number rounding + PRMT packing.
__device__ __forceinline__ int32_t requant_s32(int32_t acc, int32_t bias, float scale, int32_t lo)
{
    const int32_t t = acc + bias;
    return max(__float2int_rn(__fmul_rn(__int2float_rn(t), scale)), lo);
}

// {a, b, c, d} (int32, already >= -128) -> four saturated int8 packed little-endian.
__device__ __forceinline__ uint32_t pack4_sat_s8(int32_t a, int32_t b, int32_t c, int32_t d)
{
    uint32_t hi, r;
    asm("cvt.pack.sat.s8.s32.b32 %0, %1, %2, 0;" : "=r"(hi) : "r"(d), "r"(c));
    asm("cvt.pack.sat.s8.s32.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(b), "r"(a), "r"(hi));
    return r;
}

// Four accumulators (bias already added) -> four requantised int8, packed: the same rule as requant_s32 + pack4_sat_s8 with
// the clamp at zero applied to the PACKED bytes (sign-replicating PRMT + AND), so that nothing sits between the float ->
// int conversions and the saturating pack and ptxas fuses them into two F2IP; the scale multiplies are two FMUL2
// (mul.rn.f32x2: each lane the same IEEE product as __fmul_rn).  14 instructions per 4 outputs instead of 26.
__device__ __forceinline__ uint32_t requant4_pack(const int32_t (&acc)[4], const float (&sc)[4], bool relu)
{
    float x0 = __int2float_rn(acc[0]), x1 = __int2float_rn(acc[1]), x2 = __int2float_rn(acc[2]), x3 = __int2float_rn(acc[3]);
    uint64_t a, b, m, n;
    asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "f"(x0), "f"(x1));
    asm("mov.b64 %0, {%1, %2};" : "=l"(m) : "f"(sc[0]), "f"(sc[1]));
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(a) : "l"(a), "l"(m));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(x0), "=f"(x1) : "l"(a));
    asm("mov.b64 %0, {%1, %2};" : "=l"(b) : "f"(x2), "f"(x3));
    asm("mov.b64 %0, {%1, %2};" : "=l"(n) : "f"(sc[2]), "f"(sc[3]));
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(b) : "l"(b), "l"(n));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(x2), "=f"(x3) : "l"(b));
    uint32_t r = pack4_sat_s8(__float2int_rn(x0), __float2int_rn(x1), __float2int_rn(x2), __float2int_rn(x3));
    if (relu) {
        uint32_t neg;
        asm("prmt.b32 %0, %1, %1, 0xba98;" : "=r"(neg) : "r"(r));     // 0xff in every byte whose sign bit is set
        r &= ~neg;
    }
    return r;
}

__device__ __forceinline__ int8_t requant_s8(int32_t acc, int32_t bias, float scale, int32_t lo)
{
    return (int8_t)min(requant_s32(acc, bias, scale, lo), 127);
}

}  // namespace lbc
