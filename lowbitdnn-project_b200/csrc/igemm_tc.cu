// igemm_tc.cu — tcgen05 implicit-GEMM int8 convolution for sm_100a.
//
//   GEMM view:  D[M = N*P*Q][K_out] = A[M][R*S*C] * B[K_out][R*S*C]^T     (int8 x int8 -> int32 in TMEM)
//
//   A operand, three TMA modes (planner picks one per layer):
//     TILED   1x1 / stride 1 / no pad: A is the plain [M][C] matrix, 2-D tiled loads.
//     IM2COL  any R,S,stride,dilation: one TMA im2col-mode load per (filter tap, channel chunk) and stage;
//             hardware zero-fill implements the padding halo and the M tail.
//     WINDOW  stride-1 RxS: the (rows+R-1) x (cols+S-1) input halo window of a tile is loaded ONCE per channel
//             chunk (4-D tiled load, OOB zero-fill = padding) and every filter tap is issued as an MMA whose
//             shared-memory descriptor starts (r*Wt + s) pixel rows further into the same window.  Swizzled
//             K-major descriptors may start on any row (measured: tools/exp/desc_shift.cu), so the 9 taps of a
//             3x3 cost one L2->smem fill instead of nine.  Output positions that fall in the halo columns are
//             computed and dropped (87.5% of the MMA rows are useful for 56/28/112/224-wide layers).
//             For C == 16 (space-to-depth'd stems) pixel rows are 16 bytes, unswizzled, and one K=32 MMA covers
//             two horizontally adjacent taps through the descriptor's leading-dimension offset.
//   B operand:  the pre-packed filter matrix, either streamed through the mbarrier ring ([bn][<=128 B] blocks, TMA
//               2-D loads) or - all N tiles together <= 80 KB - RESIDENT in shared memory: loaded once per CTA, no per-block
//               handshake and no L2 re-fetch per tile (stems, ResNet stage 1, MobileNetV2 pointwise layers).
//   MMA      :  tcgen05.mma.cta_group::1.kind::i8, M=128 x N=bn x K=32 per instruction.  Issue loop: descriptor
//               offsets from a table in the kernel-parameter bank, 32-bit arithmetic on the descriptor low word,
//               batches of one filter row; two issuing warps on alternate tiles (each with its own half of the A-side
//               ring) where one warp's ~75 cycles per issue would be the bound.
//   Epilogue :  16 warps (2 teams x 8, or 4 x 4 for N tiles <= 64) drain the TMEM accumulators (tcgen05.ld 32x32b.x32),
//               fuse bias + per-channel fp32 scale + round-to-nearest-even + ReLU/saturate, pack to int8 (F2IP), stage
//               the tile in swizzled shared memory (ring of up to 3 panels, one named barrier per panel) and write it
//               with TMA stores (full 128-byte lines).  int32 output mode writes 16-byte vectors directly.
//               Ring-mode 256-wide tiles with short K loops: every warp stages and TMA-stores its own 32 rows (no team
//               barrier).  Resident-filter layers with >= 128-column tiles: the bias enters through the first MMA of each
//               tile (constant A block x bias digits), so the epilogue skips its add and its bias fetch.
//   CTA pairs:  layers that stream their filter matrix through a long K loop run as two-CTA clusters: each CTA loads its
//               own A tile and half of the B rows, the leader issues tcgen05.mma.cta_group::2 (M = 256) and multicast
//               commits; see IgemmParams::cta2.
//   Schedule :  persistent CTAs (one per SM), division-free strided walk over tiles (TileIter), mbarrier rings between
//               the TMA warps and the MMA warps, 2 or 4 TMEM accumulator stages so the epilogue of tile i overlaps
//               the main loops of the following tiles; programmatic dependent launch hides the prologue behind the
//               previous layer's tail.  A launch may walk its tiles last-to-first (IgemmLaunch::reverse) so that a
//               consumer starts on what its producer wrote last (L2 reuse; the network runner alternates directions).
//
// Replaces CUDAConv2DForward3x3TensorCoures (cpp/int8conv/conv2DForward3x3TensorCores.cuh:537-693: wmma
// m32n8k16, single-buffered 34x34x16 halo tile in smem, int32 stores, 3x3/stride-1/VALID only).
#include "common.cuh"
#include "ptx.cuh"

#include <algorithm>
#include <type_traits>
#include <mutex>

namespace lbc {

namespace {

// Epilogue formulation switches (bit-exact either way; compile-time so that A/B builds can be compared):
//   LBC_EPI_PIPE   two 16-column register buffers, the tcgen05.ld of chunk i+1 in flight while chunk i is converted
//   LBC_EPI_F32X2  the per-channel scale multiply as packed mul.rn.f32x2 (FMUL2: one issue slot per two outputs)
// (LBC_EPI_PIPE is off: holding a second 16-register buffer across the conversion makes every variant spill at the
//  96-register cap of a 640-thread CTA - 0.5 KB stack frames - and four epilogue warps per scheduler already cover the
//  TMEM latency; the epilogue is bound by pipe throughput, ~2 cycles per ALU instruction, not by latency.)
#ifndef LBC_EPI_PIPE
#define LBC_EPI_PIPE 0
#endif
#ifndef LBC_EPI_F32X2
#define LBC_EPI_F32X2 1
#endif

// Pipeline tracing (clock64 stamps of CTA 0, see lbc_conv_plan_set_trace) is compiled in only with -DLBC_TRACE=1
// (lib/liblowbit_cnn_trace.so, used by tools/trace_layer.py): the three stamps per epilogue tile were ~30 of the ~220
// instructions a warp spends per tile outside its conversion loop (ncu source counters, r02).
#ifndef LBC_TRACE
#define LBC_TRACE 0
#endif

constexpr int kBlockM = 128;
constexpr int kEpiWarps = 16;                          // epilogue warps: 2 teams of 8 or 4 teams of 4
constexpr int kFirstEpiWarp = 4;                       // warp 0: ring TMA, 1: MMA (even tiles) + TMEM alloc, 2: window TMA, 3: MMA (odd tiles)
constexpr int kNumThreads = (kFirstEpiWarp + kEpiWarps) * 32;   // 640
constexpr int kMaxStages = 8;
constexpr int kMaxWinStages = 12;
constexpr int kMaxTab = 192;                           // A-descriptor offsets per channel chunk (taps x K-steps)

enum : int32_t { A_TILED = 0, A_IM2COL = 1, A_WINDOW = 2 };

struct IgemmParams {
    int64_t m_total;
    int32_t k_out, n_img, p, q;
    int32_t mode;                 // A_TILED / A_IM2COL / A_WINDOW
    int32_t bn;                   // N tile
    int32_t bkc;                  // bytes of K per pixel row of an A block: 16 (window only) / 32 / 64 / 128
    int32_t bkb;                  // bytes of K per row of a B block (== bkc, or S_pad*16 when bkc == 16)
    int32_t tiles_m, tiles_n;
    int32_t cblocks;              // channel chunks
    int32_t inner;                // B blocks per channel chunk: taps (bkc >= 32) or filter rows (bkc == 16); 1 for TILED
    int32_t mma_outer, mma_inner; // MMA loop nest: (cblocks, inner) in WINDOW mode, (1, cblocks*inner) otherwise
    int32_t s_taps;               // filter width S
    int32_t stride_h, stride_w, pad_h, pad_w, dil_h, dil_w;
    int32_t stages;               // ring depth (B, and A in TILED/IM2COL)
    int32_t tps;                  // blocks per ring stage (one mbarrier round trip covers tps blocks)
    uint32_t a_block_bytes, b_block_bytes;   // one block; a ring stage holds tps of each
    uint32_t a_stage_bytes, b_stage_bytes;
    // window mode
    int32_t win_stages;
    uint32_t win_stage_bytes, win_tx_bytes;
    int32_t wt;                   // window pitch in pixels = cols_per_tile + (S_eff-1)*dil_w
    int32_t rows_per_tile, cols_per_tile, row_tiles, col_tiles;
    // epilogue
    int32_t relu, out_mode;
    uint32_t tmem_cols;
    int32_t n_acc;                // TMEM accumulator stages (2, 4 or 8): how far the MMA warp may run ahead of the epilogue
    int32_t tpi;                  // tiles an epilogue team takes per iteration (2 for N tiles <= 64 columns with 8 stages)
    int32_t panel_bytes, panel_swz_bits, n_panels;
    int32_t stage_bufs;           // staging panels per epilogue team (2 or 3: TMA stores drain while later panels fill)
    int32_t team_warps;           // 8: two epilogue teams (wide N tiles); 4: four teams (N tile <= 64 columns)
    int32_t warp_store;           // 1: each epilogue warp owns a 32-row staging buffer and issues its own TMA stores
    // Column-split epilogue (N tiles > 128 columns, i.e. two TMEM accumulator stages): both 8-warp teams drain EVERY tile,
    // each its own panels (int8) / column half (int32), instead of alternate tiles.  With alternate tiles a team holds an
    // accumulator stage for a whole ~2900-cycle drain (two drains share the issue slots), and with only two stages the
    // MMA of tile k+2 cannot start before the drain of tile k has ended: a latency chain MMA -> drain -> MMA that ran the
    // 1x1 channel expansions at ~2100-2600 cycles per tile against ~1500 of epilogue work (traces, r02).  Sharing the
    // tile halves the time a stage is held, so the MMA of the next tile hides completely behind the drain.
    int32_t epi_split;
    // Bias folded into the MMA (resident-filter kernels): the first MMA of every tile multiplies a constant A block (every
    // row = 31 x 127, 1) with a B block of per-channel bias digits (bias = 127 * sum(d_0..d_30) + d_31), so the
    // accumulator starts at the bias and the epilogue skips its add and the bias fetch (1.25 of its 4.75 instructions
    // per output).  Both blocks live at off_fold, written once per CTA by the epilogue warps; biases beyond +-500000 turn
    // the fold off for the launch (ctl->fold_ok).
    int32_t fold;
    uint32_t off_fold;
    int32_t rev_m;                // ring modes: > 0 = number of M tiles, visited in reverse order (see TileIter::m0)
    int32_t early_b;              // resident filter matrix: fetch it before griddepcontrol.wait (nothing in the stream writes it)
    // Tail split (CTA pairs, ring modes, one 256-wide N tile, int8 out): when the pair-steps do not divide by the pairs
    // (stage 3 of ResNet-50: 392 steps for 74 pairs), the R leftover steps of the last round are run as 2R half-steps of
    // 128 columns on 2R pairs, so the round lasts half a tile.  The half-tiles are virtual tiles behind the real ones:
    // tile index tail_first + c belongs to CTA c (c < tail_count = 4R) and means M tile tail_m0 + 2*(c/4) + (c&1),
    // columns [128*((c/2)&1), +128).  tail_first < 0: off.
    int32_t tail_first, tail_count, tail_m0;
    int32_t k_mod;                // bias/scale index = channel % k_mod (pixel-group rewrite replicates them), 0 = plain
    // smem carve-up (byte offsets from the 1024-aligned base)
    uint32_t off_b, off_stage, off_ctl;
    // MMA issue table: A-descriptor offset (16-byte units, relative to the A stage / window base) of every MMA of one
    // channel chunk, in issue order.  Lives in the kernel-parameter bank so the issue loop reads it with uniform loads.
    int32_t n_tab;
    uint16_t a_tab[kMaxTab];
    uint16_t b_tab[kMaxTab];      // resident-B window mode: B-descriptor offsets of the same MMAs (relative to the chunk)
    int32_t res_b;                // 1: the filter matrix stays in shared memory (loaded once per CTA)
    int32_t res_one;              // with res_b: only THIS CTA's N tile is resident.  The persistent grid is a multiple of
                                  // tiles_n and tiles are numbered N-tile-fastest, so CTA b only ever works on N tile
                                  // b % tiles_n ("N-stationary"): the 256->1024 / 512->2048 expansions keep their 64 / 128 KB
                                  // tile instead of re-streaming it from L2 for every M tile (96 -> 32 KB of fill per tile)
    int32_t n_mma;                // MMA-issuing warps (1 or 2); CTA-local tile L belongs to warp L % n_mma and to the
                                  // L % n_mma-th sub-ring of the A ring / window ring (stages / n_mma stages each)
    uint32_t b_total_bytes;       // resident B: bytes of the filter matrix
    // division-free tile iteration: a tile index is the mixed-radix number (img | rt | ct | n_blk) with radices
    // (it_rows, it_cols, tiles_n); step_* are the digits of gridDim.x in that system
    int32_t it_cols, it_rows;
    int32_t step_nb, step_ct, step_rt, step_img;
    int32_t tile_stride;          // tile indices one CTA advances per step: gridDim.x, or 2 * gridDim.x in pair mode
    // the same for an epilogue team, which takes every n_teams-th tile of its CTA: digits of n_teams * tile_stride
    int32_t team_stride, tstep_nb, tstep_ct, tstep_rt, tstep_img;
    // Pair mode (window A, streaming B): a CTA works on TWO consecutive M tiles at once - two windows per window stage,
    // two TMEM accumulators - and every B block fetched from L2 feeds the MMAs of both, halving the bytes per MMA that
    // bound wide 3x3 layers.  Tiles are then numbered N-tile-major (n_blk | img | rt | ct) so the two tiles of a pair
    // are the consecutive indices 2q, 2q + 1; it_imgs is the (padded, so the count is even) image radix.
    int32_t pair, it_imgs;
    // CTA-pair mode (cta_group::2, streaming B): CTAs 2c and 2c+1 form a cluster and work on the consecutive M tiles
    // 2q, 2q+1 of one N tile (same N-tile-major numbering, tile space padded to an even count).  Each CTA loads its own
    // A tile / window and HALF of the B rows; the leader (even) CTA issues M=256 MMAs that read both halves, so the
    // shared-memory fill and operand-read traffic per MMA of each SM drops from A + B to A + B/2.
    int32_t cta2, n_major;
    uint32_t win_sub_bytes;       // pair mode: bytes of one of the two windows of a stage
    // optional pipeline trace (development aid): CTA 0 writes clock64 stamps, 16 slots per local tile
    long long* trace;
    int32_t trace_tiles;
    int* flag;                    // the plan's device watchdog word (1: a bounded wait gave up, 2: misaligned smem base)
};

enum : int { EV_P_ISSUE = 0, EV_W_ISSUE, EV_M_START, EV_M_WIN, EV_M_FULL, EV_M_DONE, EV_E_START, EV_E_DRAINED, EV_E_STORED,
             EV_P_DONE };

// `local` = index of the tile in this CTA's own sequence (0, 1, 2, ...)
// `tracing` is the CTA-uniform "trace buffer attached and this is CTA 0", evaluated once per role: the per-call cost
// in a normal run is one predicated branch
__device__ __forceinline__ void trace_ev(const IgemmParams& prm, bool tracing, int32_t local, int ev)
{
    if (tracing && local < prm.trace_tiles) prm.trace[local * 16 + ev] = clock64();
}

struct Ctl {
    uint64_t full[kMaxStages];
    uint64_t empty[kMaxStages];
    uint64_t wfull[kMaxWinStages];
    uint64_t wempty[kMaxWinStages];
    uint64_t tmem_full[8];
    uint64_t tmem_empty[8];
    uint64_t bfull;               // resident filter matrix has landed
    uint64_t bias_ready;          // bias-digit block written (one arrival per epilogue warp)
    uint32_t tmem_base;
    uint32_t fold_ok;             // 0: some |bias| is out of the digit range - the epilogue adds the bias as usual
    alignas(16) float scale[4][256];     // [team][column of the N tile] (4 teams only exist for N tiles <= 64)
    alignas(16) int32_t bias[4][256];
};

// Persistent tile walk without divisions in the loop: the digits are decoded once (init) and then advanced by the
// digits of gridDim.x with carries (next).  In the ring modes it_cols == it_rows == 1 and `img` is the M-tile index.
// W: window modes (tiles are (image, row tile, column tile)); ring modes have it_cols == it_rows == 1, so their digits
// and carries vanish at compile time.  `nm` (N-tile-major numbering: CTA pairs / paired tiles) is a run-time flag only
// where the kernel variant can have both numberings.
template <bool W>
struct TileIter {
    int32_t tile, n_blk, ct, rt, img, local;
    // loop constants, copied into registers once: the producer roles are single warps whose per-tile instruction count
    // (not the TMA engine) paces the rings - every constant-bank reload in next() showed up in the traces
    int32_t stride, tiles_n, cols, rows, imgs, s_nb, s_ct, s_rt, s_img, rpt, cpt, rev_m;
    bool pair;
    // team_steps: advance by an epilogue team's stride (n_teams tiles of the CTA's sequence) per next()
    __device__ __forceinline__ void init(const IgemmParams& prm, int32_t t0, bool nm, bool team_steps = false)
    {
        stride = prm.tile_stride; tiles_n = prm.tiles_n; imgs = prm.it_imgs;
        cols = W ? prm.it_cols : 1; rows = W ? prm.it_rows : 1;
        s_nb = prm.step_nb; s_ct = prm.step_ct; s_rt = prm.step_rt; s_img = prm.step_img;
        if (team_steps) {
            stride = prm.team_stride;
            s_nb = prm.tstep_nb; s_ct = prm.tstep_ct; s_rt = prm.tstep_rt; s_img = prm.tstep_img;
        }
        rpt = prm.rows_per_tile; cpt = prm.cols_per_tile;
        rev_m = prm.rev_m;
        pair = nm;
        tile = t0;
        local = 0;
        int32_t mt;
        if (pair) {   // N-tile-major numbering
            const int32_t per_n = cols * rows * imgs;
            n_blk = t0 / per_n;
            mt = t0 - n_blk * per_n;
        } else {
            n_blk = t0 % tiles_n;
            mt = t0 / tiles_n;
        }
        if (W) {
            ct = mt % cols;
            rt = (mt / cols) % rows;
            img = mt / (cols * rows);
        } else {
            ct = rt = 0;
            img = mt;
        }
    }
    __device__ __forceinline__ void next(const IgemmParams&)
    {
        tile += stride;
        ++local;
        if (pair) {
            if (W) {
                ct += s_ct;
                if (ct >= cols) { ct -= cols; ++rt; }
                rt += s_rt;
                if (rt >= rows) { rt -= rows; ++img; }
            }
            img += s_img;
            if (img >= imgs) { img -= imgs; ++n_blk; }
            n_blk += s_nb;
        } else {
            n_blk += s_nb;
            if (W) {
                if (n_blk >= tiles_n) { n_blk -= tiles_n; ++ct; }
                ct += s_ct;
                if (ct >= cols) { ct -= cols; ++rt; }
                rt += s_rt;
                if (rt >= rows) { rt -= rows; ++img; }
            } else if (n_blk >= tiles_n) { n_blk -= tiles_n; ++img; }
            img += s_img;
        }
    }
    // the M tile right after this one (pair mode: the second tile of the pair; same n_blk because the count is even)
    __device__ __forceinline__ TileIter succ(const IgemmParams&) const
    {
        TileIter t = *this;
        ++t.tile;
        if (++t.ct >= cols) { t.ct = 0; if (++t.rt >= rows) { t.rt = 0; ++t.img; } }
        return t;
    }
    // WINDOW: first output row / column of the tile; ring modes: first GEMM row
    __device__ __forceinline__ int32_t p0(const IgemmParams&) const { return rt * rpt; }
    __device__ __forceinline__ int32_t q0(const IgemmParams&) const { return ct * cpt; }
    // ring modes: first GEMM row.  With `rev_m` (> 0: the number of real M tiles) the M tiles are visited last-to-first, so
    // a layer starts on the rows its producer wrote last - the part of its input most likely still in L2 (the network
    // runner alternates the direction along every producer -> consumer edge: +1.3% on ResNet-50).
    __device__ __forceinline__ int32_t m0() const { return image() * kBlockM; }
    // window modes: the image of the tile (same reversal, rev_m = number of images; padding tiles keep their index >= N)
    __device__ __forceinline__ int32_t image() const { return (rev_m > 0 && img < rev_m) ? rev_m - 1 - img : img; }
};

// bounded wait for the single-thread roles: false => give up (the watchdog flag is set)
// (whole warps call it convergently: the vote-based form keeps the producers' loop state on the uniform datapath)
__device__ __forceinline__ bool wait_or_quit(uint64_t* bar, uint32_t parity, volatile int* flag)
{
    return ptx::mbar_wait_u(bar, parity, flag);
}

// One 16-column group of one output pixel: requantise (bias/scale from smem) and return 16 packed int8.
// The rule is requant_s32's (round half-to-even with NaN -> 0, then clamp): cvt.rni.s32.f32 rounds and saturates to
// int32, the saturating int32 -> int8 pack clamps to [-128, 127], and with RELU the clamp at zero is applied to the
// PACKED bytes (sign-replicating PRMT + AND: 2 instructions per 4 outputs instead of 4 IMNMX).
// FOLD: the accumulator already contains the bias (see IgemmParams::fold)
// two independent round-to-nearest fp32 products in one instruction (FMUL2); each lane is the same IEEE multiply as
// __fmul_rn, so the results are bit-identical to the scalar form
__device__ __forceinline__ void fmul2_rn(float& a, float& b, float sa, float sb)
{
#if LBC_EPI_F32X2
    uint64_t x, s, r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(x) : "f"(a), "f"(b));
    asm("mov.b64 %0, {%1, %2};" : "=l"(s) : "f"(sa), "f"(sb));
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(x), "l"(s));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(r));
#else
    a = __fmul_rn(a, sa);
    b = __fmul_rn(b, sb);
#endif
}

template <bool RELU, bool FOLD>
__device__ __forceinline__ uint4 requant16(const uint32_t* v, const float* sc, const int32_t* bi)
{
    uint32_t w[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        const float4 f = *reinterpret_cast<const float4*>(sc + 4 * t);
        const int4 b = FOLD ? make_int4(0, 0, 0, 0) : *reinterpret_cast<const int4*>(bi + 4 * t);
        float x0 = __int2float_rn((int32_t)v[4 * t + 0] + b.x), x1 = __int2float_rn((int32_t)v[4 * t + 1] + b.y);
        float x2 = __int2float_rn((int32_t)v[4 * t + 2] + b.z), x3 = __int2float_rn((int32_t)v[4 * t + 3] + b.w);
        fmul2_rn(x0, x1, f.x, f.y);
        fmul2_rn(x2, x3, f.z, f.w);
        const int32_t q0 = __float2int_rn(x0), q1 = __float2int_rn(x1), q2 = __float2int_rn(x2), q3 = __float2int_rn(x3);
        uint32_t r = pack4_sat_s8(q0, q1, q2, q3);
        if (RELU) {
            uint32_t neg;
            asm("prmt.b32 %0, %1, %1, 0xba98;" : "=r"(neg) : "r"(r));     // 0xff in every byte whose sign bit is set
            r &= ~neg;
        }
        w[t] = r;
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
}

struct EpiThread {
    bool valid;          // this TMEM lane holds a real output pixel of the tile (window mode may drop halo lanes)
    uint32_t srow;       // row of the tile's staging buffer
    int32_t wrow, wcol;  // window mode: output row / column inside the tile
};

// Process NG 16-column groups held in v[]: tile columns [c, c + 16*NG), which are staging-panel columns
// [pc, pc + 16*NG).  OUT8 selects the fused int8 path (requantise -> swizzled staging panel) or raw int32 stores.
template <int NG, bool OUT8, bool RELU, bool FOLD>
__device__ __forceinline__ void epi_consume(const IgemmParams& prm, const float* sc, const int32_t* bi,
                                            const uint32_t* v, int32_t c, int32_t pc, const EpiThread& et,
                                            uint32_t staging, uint32_t row_off, uint32_t swz_mask, int32_t* y32,
                                            int64_t out_row, int32_t col0, int32_t& wait_mode)
{
#pragma unroll
    for (int g = 0; g < NG; ++g) {
        const int32_t cc = c + 16 * g;
        if (OUT8) {
            const uint4 r = requant16<RELU, FOLD>(v + 16 * g, sc + cc, bi + cc);
            // per-warp stores: the staging buffer may still be being read by this warp's earlier TMA store.  The wait sits
            // HERE, behind the first group's conversion, so that the store's smem-read latency runs under ~75 instructions
            // instead of in front of them (1: wait for every store, 2: all but the most recent one - two buffers)
            if (wait_mode) {
                if (ptx::lane_id() == 0) {
                    if (wait_mode == 2) ptx::tma_store_wait_read<1>();
                    else ptx::tma_store_wait_read<0>();
                }
                __syncwarp();
                wait_mode = 0;
            }
            // swizzle: the XOR term depends only on the staging row (a panel row never crosses a 128-byte line), so
            // `swz_mask` arrives here already as this thread's ((row_off >> 7) & mask) << 4
            // `staging` arrives as a 32-bit shared-window address (one conversion per panel, not one per store)
            if (et.valid) ptx::st_shared_v4(staging + ((row_off + (uint32_t)(pc + 16 * g)) ^ swz_mask), r.x, r.y, r.z, r.w);
            if (FOLD) asm volatile("" ::: "memory");   // keep the next group's parameter loads behind this store (register cap)
        } else if (out_row >= 0 && col0 + cc < prm.k_out) {
            int32_t* yo = y32 + out_row * prm.k_out + col0 + cc;
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
                const int4 b = FOLD ? make_int4(0, 0, 0, 0) : *reinterpret_cast<const int4*>(bi + cc + j);
                ptx::st_global_v4(yo + j, v[16 * g + j] + (uint32_t)b.x, v[16 * g + j + 1] + (uint32_t)b.y,
                                  v[16 * g + j + 2] + (uint32_t)b.z, v[16 * g + j + 3] + (uint32_t)b.w);
            }
        }
    }
}

// Drain this warp's share [c0, c1) of one panel out of TMEM.
template <bool OUT8, bool RELU, bool FOLD>
__device__ __forceinline__ void epi_drain(const IgemmParams& prm, const float* sc, const int32_t* bi, uint32_t taddr,
                                          int32_t pbase, int32_t c0, int32_t c1, const EpiThread& et, uint32_t staging,
                                          uint32_t row_off, uint32_t swz_mask, int32_t* y32, int64_t out_row,
                                          int32_t col0, int32_t wait_mode)
{
    int32_t c = c0;
#if LBC_EPI_PIPE
    // Software pipeline over 16-column chunks with two register buffers: right after the wait that completes chunk i the
    // load of chunk i+1 is issued, so its TMEM latency runs under the ~75 instructions that convert chunk i.
    // (tcgen05.wait::ld waits for every outstanding load of the thread, hence exactly one load in flight at a wait.)
    uint32_t va[16], vb[16];
    if (c + 16 <= c1) ptx::tmem_ld_32x32b_x16(taddr + (uint32_t)c, va);
    while (c + 16 <= c1) {
        ptx::tmem_ld_wait_dep16(va);
        if (c + 32 <= c1) ptx::tmem_ld_32x32b_x16(taddr + (uint32_t)(c + 16), vb);
        epi_consume<1, OUT8, RELU, FOLD>(prm, sc, bi, va, c, c - pbase, et, staging, row_off, swz_mask, y32, out_row, col0, wait_mode);
        c += 16;
        if (c + 16 > c1) break;
        ptx::tmem_ld_wait_dep16(vb);
        if (c + 32 <= c1) ptx::tmem_ld_32x32b_x16(taddr + (uint32_t)(c + 16), va);
        epi_consume<1, OUT8, RELU, FOLD>(prm, sc, bi, vb, c, c - pbase, et, staging, row_off, swz_mask, y32, out_row, col0, wait_mode);
        c += 16;
    }
#else
    for (; c + 32 <= c1; c += 32) {
        uint32_t v[32];
        ptx::tmem_ld_32x32b_x32(taddr + (uint32_t)c, v);
        ptx::tmem_ld_wait_dep(v);
        epi_consume<2, OUT8, RELU, FOLD>(prm, sc, bi, v, c, c - pbase, et, staging, row_off, swz_mask, y32, out_row, col0, wait_mode);
    }
    if (c + 16 <= c1) {
        uint32_t v16[16];
        ptx::tmem_ld_32x32b_x16(taddr + (uint32_t)c, v16);
        ptx::tmem_ld_wait_dep16(v16);
        epi_consume<1, OUT8, RELU, FOLD>(prm, sc, bi, v16, c, c - pbase, et, staging, row_off, swz_mask, y32, out_row, col0, wait_mode);
    }
#endif
}

// One-time set-up of the bias-fold operand blocks (see IgemmParams::fold), by the 512 epilogue threads.
// Called before the role computes its per-thread constants, so its temporaries do not overlap their live ranges.
__device__ __forceinline__ void write_fold_blocks(uint8_t* fold_base, uint32_t etid, int32_t bn, int32_t c_first, int32_t k_out,
                                               int32_t k_mod, const int32_t* __restrict__ bias, uint32_t* fold_ok)
{
    // [A': 4 KB][B': bn x 32 B], both as 8-row x 16-byte core matrices (K chunks 128 B apart, 8-row groups 256 B apart)
    // A': every row = 31 x 127, then 1.  16-byte piece i sits at i * 16 and belongs to K chunk (i >> 3) & 1.
    for (uint32_t i = etid; i < 256u; i += kEpiWarps * 32u) {
        const uint32_t last = ((i >> 3) & 1u) ? 0x017f7f7fu : 0x7f7f7f7fu;
        *reinterpret_cast<uint4*>(fold_base + i * 16u) = make_uint4(0x7f7f7f7fu, 0x7f7f7f7fu, 0x7f7f7f7fu, last);
    }
    // B': one thread per output channel (local row c holds global channel c_first + c)
    bool ok = true;
    for (uint32_t c = etid; c < (uint32_t)bn; c += kEpiWarps * 32u) {
        const int32_t kc = c_first + (int32_t)c;
        const bool in = kc < k_out;
        const int32_t kp = k_mod ? kc % k_mod : kc;
        int32_t b = (in && bias) ? __ldg(bias + kp) : 0;
        if (b > 500000 || b < -500000) { ok = false; b = 0; }
        int32_t q = (b + (b >= 0 ? 63 : -63)) / 127;          // bias = 127 * q + r0, |r0| <= 63
        const int32_t r0 = b - 127 * q;
        uint32_t wds[8];
#pragma unroll
        for (int wi = 0; wi < 8; ++wi) {
            uint32_t wv = 0;
#pragma unroll
            for (int bj = 0; bj < 4; ++bj) {
                int32_t dgt;
                if (wi == 7 && bj == 3) dgt = r0;
                else { dgt = max(-127, min(127, q)); q -= dgt; }
                wv |= ((uint32_t)dgt & 0xffu) << (8 * bj);
            }
            wds[wi] = wv;
        }
        uint8_t* dst = fold_base + 4096u + (c >> 3) * 256u + (c & 7u) * 16u;
        *reinterpret_cast<uint4*>(dst) = make_uint4(wds[0], wds[1], wds[2], wds[3]);
        *reinterpret_cast<uint4*>(dst + 128) = make_uint4(wds[4], wds[5], wds[6], wds[7]);
    }
    if (!ok) *reinterpret_cast<volatile uint32_t*>(fold_ok) = 0u;
}

// Selects the epilogue instantiation for the (warp-uniform) run-time switches.  MAY_FOLD is false in kernels that never
// fold the bias (streaming filter matrix), so they carry no second copy of the conversion code.
template <bool MAY_FOLD>
__device__ __forceinline__ void epi_run(bool int8_out, bool relu, bool fold, const IgemmParams& prm, const float* sc,
                                        const int32_t* bi, uint32_t taddr, int32_t pbase, int32_t c0, int32_t c1, const EpiThread& et,
                                        uint32_t staging, uint32_t row_off, uint32_t swz_mask, int32_t* y32, int64_t out_row,
                                        int32_t col0, int32_t wait_mode = 0)
{
#define LBC_EPI(O8, RL, FD) epi_drain<O8, RL, FD>(prm, sc, bi, taddr, pbase, c0, c1, et, staging, row_off, swz_mask, y32, out_row, col0, wait_mode)
    if (MAY_FOLD && fold) {
        if (int8_out && relu) LBC_EPI(true, true, true);
        else if (int8_out) LBC_EPI(true, false, true);
        else LBC_EPI(false, false, true);
    } else {
        if (int8_out && relu) LBC_EPI(true, true, false);
        else if (int8_out) LBC_EPI(true, false, false);
        else LBC_EPI(false, false, false);
    }
#undef LBC_EPI
}

// KM: 0 tiled A, 1 im2col A, 2 window A (>= 32-byte pixels), 3 window A with 16-byte pixels (paired taps)
// KS: MMA K-steps (32 bytes each) per B block = bkb / 32
// RESB: the filter matrix is resident in shared memory (one N tile, loaded once per CTA); the ring then carries
//       only A blocks (tiled / im2col) or does not exist at all (window modes)
template <bool CTA2>
__device__ __forceinline__ void mma_issue(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                          uint32_t accumulate, uint32_t leader)
{
    if (CTA2) ptx::mma_i8_ss_pred32_2cta(tmem_d, a_lo, a_hi, b_lo, b_hi, idesc, accumulate, leader);
    else ptx::mma_i8_ss_pred32(tmem_d, a_lo, a_hi, b_lo, b_hi, idesc, accumulate, leader);
}

template <bool CTA2>
__device__ __forceinline__ void mma_commit(uint64_t* bar, uint32_t leader)
{
    if (CTA2) ptx::mma_commit_2cta_pred(bar, leader);     // arrives on the barrier at this offset in BOTH CTAs
    else ptx::mma_commit_pred(bar, leader);
}

// CTA2: CTA-pair mode (see IgemmParams::cta2); only instantiated with RESB == false
// MAYFOLD: the launch may fold the bias into the MMA (see IgemmParams::fold); only with RESB.  A separate instantiation,
// because carrying the folded copy of the tile loops costs the narrow-tile kernels registers (spills: +4-7% time on
// 64-column tiles, measured), so kernels that never fold are compiled without it.
template <int KM, int KS, bool RESB, bool CTA2, bool MAYFOLD>
__global__ void __launch_bounds__(kNumThreads, 1)
igemm_i8_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                const __grid_constant__ CUtensorMap tm_out, const __grid_constant__ CUtensorMap tm_out2, const IgemmParams prm,
                const int32_t* __restrict__ bias, const float* __restrict__ scale, void* __restrict__ y)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* smem_a = smem;                         // A ring (TILED/IM2COL) or window ring (WINDOW)
    uint8_t* smem_b = smem + prm.off_b;
    uint8_t* staging = smem + prm.off_stage;
    Ctl* ctl = reinterpret_cast<Ctl*>(smem + prm.off_ctl);

    // the warp index through a shuffle: the compiler then KNOWS it is warp-uniform, so the role dispatch is a uniform branch
    // and everything a role derives from it (ring halves, stage cursors, descriptors) can live on the uniform datapath
    // the warp index through a shuffle: the compiler then KNOWS it is warp-uniform, so the role dispatch is a uniform branch
    // and everything a role derives from it (ring halves, stage cursors, descriptors) can live on the uniform datapath.
    // (Putting the four issue roles on the four highest hardware warps - the scheduler is said to prefer the highest ready
    //  warp id - changed nothing measurable: +-1% per layer, r02.)
    const uint32_t warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    const uint32_t lane = threadIdx.x & 31;
    volatile int* tflag = prm.flag;
    const uint32_t cta_rank = CTA2 ? ptx::cluster_ctarank() : 0u;   // 0 = leader of the pair
    // what the template parameters already decide (the run-time flags only matter where a variant can have both)
    constexpr bool kWindow = (KM >= 2);
    const bool pair_mode = kWindow && !RESB && !CTA2 && prm.pair != 0;   // paired tiles: window A, streaming B, no CTA pairs
    const bool n_major = CTA2 || pair_mode;                               // N-tile-major tile numbering
    using Iter = TileIter<kWindow>;
    const bool tracing = LBC_TRACE && prm.trace != nullptr && blockIdx.x == 0;
    // trace build: four wall-clock stamps per CTA (%globaltimer, ns: comparable across SMs and launches) behind the
    // 2 x grid clock64 stamps - kernel entry, set-up done, past griddepcontrol.wait, last store issued
    long long* const gt = (LBC_TRACE && prm.trace != nullptr && threadIdx.x == 0)
                              ? prm.trace + (size_t)prm.trace_tiles * 16 + 2 * (size_t)gridDim.x + 4 * (size_t)blockIdx.x : nullptr;
    if (LBC_TRACE && gt) gt[0] = (long long)ptx::globaltimer_ns();

    if ((ptx::smem_u32(smem) & 1023u) != 0) {       // swizzle atoms need a 1024-byte aligned base
        if (threadIdx.x == 0) *tflag = 2;
        return;
    }

    ptx::griddep_launch_dependents();   // the next layer's CTAs may take this SM as soon as this CTA has left it
    if (warp == 0 && lane == 0) {
        ptx::prefetch_tensormap(&tm_a);
        ptx::prefetch_tensormap(&tm_b);
        ptx::prefetch_tensormap(&tm_out);
        ptx::prefetch_tensormap(&tm_out2);
        const uint32_t consumers = pair_mode ? 2u : 1u;   // pair mode: both MMA warps read every stage
        for (int i = 0; i < prm.stages; ++i) {
            ptx::mbar_init(&ctl->full[i], 1);
            ptx::mbar_init(&ctl->empty[i], consumers);
        }
        for (int i = 0; i < prm.win_stages; ++i) {
            ptx::mbar_init(&ctl->wfull[i], 1);
            ptx::mbar_init(&ctl->wempty[i], consumers);
        }
        for (int i = 0; i < prm.n_acc; ++i) {
            ptx::mbar_init(&ctl->tmem_full[i], 1);
            // one arrival per warp that drains the stage: a team, or (column-split epilogue) all 16 warps; pair: both CTAs
            ptx::mbar_init(&ctl->tmem_empty[i], (uint32_t)(prm.epi_split ? kEpiWarps : prm.team_warps) * (CTA2 ? 2u : 1u));
        }
        ptx::mbar_init(&ctl->bfull, 1);
        ptx::mbar_init(&ctl->bias_ready, kEpiWarps);
        ctl->fold_ok = 1;
        ptx::fence_barrier_init();
    }
    if (warp == 1) {
        if (CTA2) {
            ptx::tmem_alloc_2cta(&ctl->tmem_base, prm.tmem_cols);
            ptx::tmem_relinquish_2cta();
        } else {
            ptx::tmem_alloc(&ctl->tmem_base, prm.tmem_cols);
            ptx::tmem_relinquish();
        }
    }
    ptx::tc_fence_before();
    if (CTA2) ptx::cluster_sync();     // the peer's barriers must be initialised before anything signals them
    else __syncthreads();
    ptx::tc_fence_after();
    // Programmatic dependent launch: everything above is on-chip set-up (barriers, TMEM, descriptor prefetch) and may
    // overlap the tail of the previous kernel in the stream; from here on global memory is read and written.
    const bool tail_on = CTA2 && !kWindow && prm.tail_first >= 0;       // see IgemmParams::tail_first
    const int32_t num_tiles = tail_on ? prm.tail_first + prm.tail_count
                              : n_major ? prm.it_cols * prm.it_rows * prm.it_imgs * prm.tiles_n : prm.tiles_m * prm.tiles_n;
    const int32_t first_tile = pair_mode ? 2 * (int32_t)blockIdx.x : (int32_t)blockIdx.x;
    // this CTA's half-tile of the split last round: first GEMM row (with the launch's traversal direction) and column half
    int32_t tail_row0 = 0, tail_half = 0;
    if (tail_on) {
        const int32_t mt = prm.tail_m0 + 2 * (int32_t)(blockIdx.x >> 2) + (int32_t)(blockIdx.x & 1u);
        tail_row0 = ((prm.rev_m > 0 && mt < prm.rev_m) ? prm.rev_m - 1 - mt : mt) * kBlockM;
        tail_half = (int32_t)((blockIdx.x >> 1) & 1u);
    }

    // ... with one exception: a resident filter matrix that no kernel in the stream writes (IgemmLaunch::early_b, set by
    // the network runner, whose weights were uploaded long before) is fetched BEFORE the wait, so its 32-128 KB arrive
    // while the previous layer drains.
    auto load_resident_b = [&]() {
        if (ptx::elect_one()) {
            // CTA pairs: each CTA keeps ITS half of the filter rows (cta_group::2 MMAs read B from both CTAs); the leader's
            // barrier collects the bytes of both halves, so the one MMA-issuing thread waits on a single barrier
            if (!CTA2 || cta_rank == 0) ptx::mbar_expect_tx(&ctl->bfull, prm.b_total_bytes * (CTA2 ? 2u : 1u));
            const int32_t nblk = prm.cblocks * prm.inner;
            uint8_t* dst = smem_b;
            // every N tile, or (N-stationary) only the one this CTA works on
            const int32_t nt0 = prm.res_one ? (int32_t)(blockIdx.x % (uint32_t)prm.tiles_n) : 0;
            const int32_t nt1 = prm.res_one ? nt0 + 1 : prm.tiles_n;
            const uint32_t bfull0 = CTA2 ? ptx::mapa(ptx::smem_u32(&ctl->bfull), 0) : 0u;
            const int32_t half_off = CTA2 ? (int32_t)cta_rank * (prm.bn >> 1) : 0;
            for (int32_t nt = nt0; nt < nt1; ++nt) {
                int32_t bcol = 0;
                for (int32_t i = 0; i < nblk; ++i, dst += prm.b_block_bytes, bcol += prm.bkb) {
                    if (CTA2) ptx::tma_load_2d_2sm(dst, &tm_b, bfull0, bcol, nt * prm.bn + half_off);
                    else ptx::tma_load_2d(dst, &tm_b, &ctl->bfull, bcol, nt * prm.bn);
                }
            }
        }
        __syncwarp();
    };
    // The same for a STREAMED filter matrix: the B blocks of this CTA's first ring stages go out before the wait (the A
    // blocks of those stages follow after it, onto the same transaction count), so the burst of first loads that all
    // CTAs send to L2 at the moment of the release is only the A third of it.
    constexpr bool kRingEarly = !(kWindow && RESB);
    const int32_t ring_pre = (!RESB && kRingEarly && prm.early_b == 1)
                                 ? min(prm.cblocks * prm.inner / prm.tps, prm.stages / prm.n_mma) : 0;
    if (ring_pre > 0 && warp == 0 && first_tile < num_tiles) {
        if (ptx::elect_one()) {
            Iter it0;
            it0.init(prm, first_tile, n_major);
            const bool tail0 = tail_on && it0.tile >= prm.tail_first;
            const int32_t brow = tail0 ? tail_half * (prm.bn >> 1) + (int32_t)cta_rank * (prm.bn >> 2)
                                       : it0.n_blk * prm.bn + (CTA2 ? (int32_t)cta_rank * (prm.bn >> 1) : 0);
            const uint32_t tx_all = (prm.b_stage_bytes + (kWindow ? 0u : prm.a_stage_bytes)) * (CTA2 ? 2u : 1u);
            const uint32_t full0 = CTA2 ? ptx::mapa(ptx::smem_u32(&ctl->full[0]), 0) : 0u;
            int32_t bcol = 0;
            for (int32_t st = 0; st < ring_pre; ++st) {
                if (cta_rank == 0) ptx::mbar_expect_tx(&ctl->full[st], tx_all);
                uint8_t* dst_b = smem_b + (uint32_t)st * prm.b_stage_bytes;
                for (int32_t t = 0; t < prm.tps; ++t, dst_b += prm.b_block_bytes, bcol += prm.bkb) {
                    if (CTA2) ptx::tma_load_2d_2sm(dst_b, &tm_b, full0 + (uint32_t)st * 8u, bcol, brow);
                    else ptx::tma_load_2d(dst_b, &tm_b, &ctl->full[st], bcol, brow);
                }
            }
        }
        __syncwarp();
    }
    if (RESB && prm.early_b && warp == 0) load_resident_b();
    if (LBC_TRACE && gt) gt[1] = (long long)ptx::globaltimer_ns();
    ptx::griddep_wait();
    if (LBC_TRACE && gt) gt[2] = (long long)ptx::globaltimer_ns();
    if (LBC_TRACE && prm.trace != nullptr && threadIdx.x == 0) prm.trace[(size_t)prm.trace_tiles * 16 + 2 * blockIdx.x] = clock64();
    const uint32_t tmem_base = ctl->tmem_base;
    // pair mode counts the padded tile space (dummy tiles of the padding image load zeros and store nothing)
    // The three issue roles below run with ALL 32 lanes of their warp executing the (warp-uniform) loops; only
    // the TMA / MMA / commit instructions themselves are predicated on one elected lane.  Keeping the loops
    // convergent lets the compiler hold addresses and descriptors in uniform registers; a `lane == 0` branch
    // around the whole loop made every tcgen05.mma cost ~150 scalar instructions (ncu, r01 v2).
    constexpr bool kRing = !(kWindow && RESB);      // resident B + window A: no ring at all
    if (warp == 0) {
        // ===================== ring producer: B blocks (+ A blocks in TILED / IM2COL) =====================
        const bool leader = ptx::elect_one();
        if (RESB) {
            // the whole filter matrix, once: per N tile, k_blocks boxes of [bn][bkb] side by side
            if (!prm.early_b) load_resident_b();
        }
        if (kRing) {
            // one (stage, phase) cursor per sub-ring: with two MMA warps, even tiles flow through the first half of
            // the ring and odd tiles through the second, so each consumer sees its stages strictly in phase order
            const uint32_t sub_len = (uint32_t)prm.stages / (uint32_t)prm.n_mma;
            uint32_t stage_e = 0, phase_e = 0, stage_o = 0, phase_o = 0;   // cursors of the even / odd sub-ring
            // pair mode: the leader's full barrier collects the bytes of both CTAs' loads
            const uint32_t tx_bytes = ((RESB ? 0u : prm.b_stage_bytes) + (kWindow ? 0u : prm.a_stage_bytes)) * (CTA2 ? 2u : 1u);
            const uint32_t full0 = CTA2 ? ptx::mapa(ptx::smem_u32(&ctl->full[0]), 0) : 0u;
            const int32_t brow_off = CTA2 ? (int32_t)cta_rank * (prm.bn >> 1) : 0;
            const int32_t stages_per_tile = prm.cblocks * prm.inner / prm.tps;
            const int32_t tps = prm.tps, cblocks = prm.cblocks;
            const int32_t bkb = prm.bkb, bkc = prm.bkc, s_taps = prm.s_taps;
            const int32_t dil_w = prm.dil_w, dil_h = prm.dil_h;
            const uint32_t a_block = prm.a_block_bytes, b_block = prm.b_block_bytes;
            const uint32_t a_stage = prm.a_stage_bytes, b_stage = prm.b_stage_bytes;
            bool ok = true;
            Iter it;
            for (it.init(prm, first_tile, n_major); it.tile < num_tiles && ok; it.next(prm)) {
                const bool tail = tail_on && it.tile >= prm.tail_first;
                const int32_t m0 = tail ? tail_row0 : it.m0();
                int32_t w_base = 0, h_base = 0, n0 = 0;
                if (KM == A_IM2COL) {
                    const uint32_t um = (uint32_t)m0, uq = (uint32_t)prm.q, up = (uint32_t)prm.p;
                    const uint32_t row = um / uq;
                    const int32_t q0 = (int32_t)(um - row * uq);
                    n0 = (int32_t)(row / up);
                    const int32_t p0 = (int32_t)(row - (uint32_t)n0 * up);
                    w_base = q0 * prm.stride_w - prm.pad_w;
                    h_base = p0 * prm.stride_h - prm.pad_h;
                }
                // (a half-tile's MMA reads the first bn/4 rows of each CTA's B block: the box still brings bn/2, the rest is unused)
                const int32_t brow = tail ? tail_half * (prm.bn >> 1) + (int32_t)cta_rank * (prm.bn >> 2) : it.n_blk * prm.bn + brow_off;
                const uint32_t sub = (uint32_t)it.local & (uint32_t)(prm.n_mma - 1);
                const uint32_t sub_base = sub * sub_len;
                uint32_t stage = sub_base + (sub ? stage_o : stage_e), phase = sub ? phase_o : phase_e;
                // ring modes walk K as [tap][channel chunk]: (off_h, off_w) filter tap offset, c0 channel offset
                int32_t c0 = 0, cbi = 0, off_w = 0, off_h = 0, fs = 0, bcol = 0;
                for (int32_t st = 0; st < stages_per_tile; ++st) {
                    ok = wait_or_quit(&ctl->empty[stage], phase ^ 1, tflag);
                    if (!ok) break;
                    if (st == 0 && leader) trace_ev(prm, tracing, it.local, EV_P_ISSUE);
                    const bool pre = it.local == 0 && st < ring_pre;      // B block and transaction count already out (above)
                    if (!pre && leader && cta_rank == 0) ptx::mbar_expect_tx(&ctl->full[stage], tx_bytes);
                    uint8_t* dst_a = smem_a + stage * a_stage;
                    uint8_t* dst_b = smem_b + stage * b_stage;
                    for (int32_t t = 0; t < tps; ++t) {
                        if (CTA2) {
                            if (leader) {
                                const uint32_t fbar = full0 + stage * 8u;
                                if (KM == A_IM2COL)
                                    ptx::tma_load_im2col_4d_2sm(dst_a, &tm_a, fbar, c0, w_base, h_base, n0, (uint16_t)off_w, (uint16_t)off_h);
                                else if (KM == A_TILED)
                                    ptx::tma_load_2d_2sm(dst_a, &tm_a, fbar, c0, m0);
                                if (!RESB && !pre) ptx::tma_load_2d_2sm(dst_b, &tm_b, fbar, bcol, brow);
                            }
                        } else if (leader) {
                            if (KM == A_IM2COL)
                                ptx::tma_load_im2col_4d(dst_a, &tm_a, &ctl->full[stage], c0, w_base, h_base, n0, (uint16_t)off_w,
                                                        (uint16_t)off_h);
                            else if (KM == A_TILED)
                                ptx::tma_load_2d(dst_a, &tm_a, &ctl->full[stage], c0, m0);
                            if (!RESB && !pre) ptx::tma_load_2d(dst_b, &tm_b, &ctl->full[stage], bcol, brow);
                        }
                        dst_a += a_block;
                        dst_b += b_block;
                        bcol += bkb;
                        if (!kWindow) {
                            c0 += bkc;
                            if (++cbi == cblocks) {
                                cbi = 0; c0 = 0;
                                off_w += dil_w;
                                if (++fs == s_taps) { fs = 0; off_w = 0; off_h += dil_h; }
                            }
                        }
                    }
                    if (++stage == sub_base + sub_len) { stage = sub_base; phase ^= 1; }
                }
                if (sub) { stage_o = stage - sub_base; phase_o = phase; }
                else { stage_e = stage - sub_base; phase_e = phase; }
                if (leader) trace_ev(prm, tracing, it.local, EV_P_DONE);
            }
        }
    } else if (warp == 2) {
        // ===================== window producer (WINDOW modes only) =====================
        if (kWindow) {
            const uint32_t sub_len = (uint32_t)prm.win_stages / (uint32_t)prm.n_mma;
            uint32_t ws_e = 0, wphase_e = 0, ws_o = 0, wphase_o = 0;
            const bool leader = ptx::elect_one();
            bool ok = true;
            const int32_t pad_w = prm.pad_w, pad_h = prm.pad_h, cblocks = prm.cblocks, bkc = prm.bkc;
            const uint32_t n_mma_mask = (uint32_t)(prm.n_mma - 1), win_stage_bytes = prm.win_stage_bytes;
            const uint32_t win_sub_bytes = prm.win_sub_bytes;
            const bool pair = pair_mode;
            const uint32_t win_tx = (pair || CTA2) ? 2u * prm.win_tx_bytes : prm.win_tx_bytes;
            const uint32_t wfull0 = CTA2 ? ptx::mapa(ptx::smem_u32(&ctl->wfull[0]), 0) : 0u;
            Iter it;
            for (it.init(prm, first_tile, n_major); it.tile < num_tiles && ok; it.next(prm)) {
                const int32_t wq = it.q0(prm) - pad_w, wp = it.p0(prm) - pad_h;
                int32_t wq1 = 0, wp1 = 0, img1 = 0;
                if (pair) {                                             // the second tile of the pair
                    const Iter it1 = it.succ(prm);
                    wq1 = it1.q0(prm) - pad_w; wp1 = it1.p0(prm) - pad_h; img1 = it1.image();
                }
                const uint32_t sub = (uint32_t)it.local & n_mma_mask;
                const uint32_t sub_base = sub * sub_len;
                uint32_t ws = sub_base + (sub ? ws_o : ws_e), wphase = sub ? wphase_o : wphase_e;
                int32_t c0 = 0;
                for (int32_t cb = 0; cb < cblocks; ++cb, c0 += bkc) {
                    ok = wait_or_quit(&ctl->wempty[ws], wphase ^ 1, tflag);
                    if (!ok) break;
                    if (leader) {
                        if (cb == 0) trace_ev(prm, tracing, it.local, EV_W_ISSUE);
                        if (cta_rank == 0) ptx::mbar_expect_tx(&ctl->wfull[ws], win_tx);
                        if (CTA2) ptx::tma_load_4d_2sm(smem_a + ws * win_stage_bytes, &tm_a, wfull0 + ws * 8u, c0, wq, wp, it.image());
                        else ptx::tma_load_4d(smem_a + ws * win_stage_bytes, &tm_a, &ctl->wfull[ws], c0, wq, wp, it.image());
                        if (pair)
                            ptx::tma_load_4d(smem_a + ws * win_stage_bytes + win_sub_bytes, &tm_a, &ctl->wfull[ws], c0, wq1, wp1, img1);
                    }
                    if (++ws == sub_base + sub_len) { ws = sub_base; wphase ^= 1; }
                }
                if (sub) { ws_o = ws - sub_base; wphase_o = wphase; }
                else { ws_e = ws - sub_base; wphase_e = wphase; }
            }
        }
    } else if (warp == 1 || warp == 3) {
        // ===================== MMA issuers (two warps, alternating tiles) =====================
        // Convergent, exit-free loops (a watchdog trip only makes the waits return early).  One tcgen05.mma of
        // M=128 x N<=128 takes 48-64 cycles (tools/exp/mma_rates.cu) but a single warp needs ~75 cycles of scalar /
        // uniform-datapath work to issue one (traces, r01), so layers with narrow N tiles were issue-bound.  Two
        // warps may therefore issue for alternate tiles of the CTA (n_mma == 2): tile L (CTA-local) belongs to warp
        // L % n_mma, uses TMEM accumulator stage L % n_acc and flows through that warp's half of the A-side ring (the
        // producers fill the two sub-rings alternately, so every consumer sees its stages strictly in phase order - a
        // shared ring would alias mbarrier parities between the two consumers).  Different tiles accumulate into
        // different TMEM columns, so the interleaving of the two instruction streams in the tensor pipe is irrelevant.
        // Descriptor offsets of a channel chunk come from tables in the kernel-parameter bank, only the low 32 bits
        // of the descriptors are ever touched (the smem address field cannot carry into the LBO field), and with a
        // resident filter matrix there is no per-block handshake at all.
        const uint32_t which = warp >> 1;                              // 0 (warp 1) or 1 (warp 3)
        const uint32_t n_mma = (uint32_t)prm.n_mma;
        const uint32_t leader = ptx::elect_one() ? 1u : 0u;
        const uint32_t idesc = ptx::make_idesc_i8(CTA2 ? 2 * kBlockM : kBlockM, (uint32_t)prm.bn);
        const uint32_t idesc_half = ptx::make_idesc_i8(CTA2 ? 2 * kBlockM : kBlockM, (uint32_t)prm.bn >> 1);
        const uint64_t db_base = ptx::make_kmajor_desc(ptx::smem_u32(smem_b), (uint32_t)prm.bkb);
        const uint64_t da_base = (KM != 3) ? ptx::make_kmajor_desc(ptx::smem_u32(smem_a), (uint32_t)prm.bkc)
                                           : ptx::make_kmajor_desc_nosw(ptx::smem_u32(smem_a), (uint32_t)prm.dil_w * 16u, 128u);
        const uint32_t da_lo = (uint32_t)da_base, da_hi = (uint32_t)(da_base >> 32);
        const uint32_t db_lo = (uint32_t)db_base, db_hi = (uint32_t)(db_base >> 32);
        const uint32_t a_stage16 = (kWindow ? prm.win_stage_bytes : prm.a_stage_bytes) >> 4;
        const uint32_t b_stage16 = prm.b_stage_bytes >> 4;
        const uint32_t b_block16 = prm.b_block_bytes >> 4;
        const uint32_t b_chunk16 = b_block16 * (uint32_t)prm.inner;          // resident B: one channel chunk of blocks
        const int32_t inner_stages = prm.mma_inner / prm.tps;
        const int32_t tps = prm.tps, mma_outer = prm.mma_outer, n_tab = prm.n_tab;
        // this warp's sub-ring: stages [ring_lo, ring_hi) of the A/B ring and [win_lo, win_hi) of the window ring
        const uint32_t ring_len = (uint32_t)prm.stages / n_mma, win_len = (uint32_t)prm.win_stages / n_mma;
        const uint32_t sub_ring = pair_mode ? 0u : which;   // pair mode: both warps walk the whole (single) ring
        const uint32_t ring_lo = sub_ring * ring_len, ring_hi = ring_lo + ring_len;
        const uint32_t win_lo = sub_ring * win_len, win_hi = win_lo + win_len;
        const uint32_t bn = (uint32_t)prm.bn;
        const uint32_t acc_mask = (uint32_t)prm.n_acc - 1u, acc_shift = prm.n_acc == 8 ? 3u : prm.n_acc == 4 ? 2u : 1u;
        uint32_t stage = ring_lo, phase = 0, ws = win_lo, wphase = 0;
        const bool active = (which < n_mma || pair_mode) && cta_rank == 0;   // CTA pairs: only the leader issues
        bool ready = (kRing && active) ? ptx::mbar_test_u(&ctl->full[stage], phase) : true;
        bool wready = (kWindow && active) ? ptx::mbar_test_u(&ctl->wfull[ws], wphase) : true;
        if (RESB && active) ptx::mbar_wait_soft_u(&ctl->bfull, 0, tflag);
        // bias folded into the MMA: the epilogue warps write the constant A block and the bias-digit B block first
        bool fold = false;
        uint32_t fa_lo = 0, fa_hi = 0, fb_lo = 0, fb_hi = 0;
        if (MAYFOLD && prm.fold && active) {
            ptx::mbar_wait_soft_u(&ctl->bias_ready, 0, tflag);
            fold = *reinterpret_cast<volatile uint32_t*>(&ctl->fold_ok) != 0;
            const uint64_t dfa = ptx::make_kmajor_desc_nosw(ptx::smem_u32(smem + prm.off_fold), 128u, 256u);
            const uint64_t dfb = ptx::make_kmajor_desc_nosw(ptx::smem_u32(smem + prm.off_fold) + 4096u, 128u, 256u);
            fa_lo = (uint32_t)dfa; fa_hi = (uint32_t)(dfa >> 32); fb_lo = (uint32_t)dfb; fb_hi = (uint32_t)(dfb >> 32);
        }
        if (pair_mode) {
            // ---- pair mode: two M tiles per step share every B block (see IgemmParams::pair).  Warp `which` issues
            // the MMAs of the which-th tile of the pair: both read the same B stage and the same window stage (one
            // window each), so those stages are released by TWO commits (their empty barriers count 2).
            const uint32_t win_sub16 = prm.win_sub_bytes >> 4;
            int32_t lp = 0;   // CTA-local pair index; its tiles are CTA-local tiles 2*lp and 2*lp + 1
            for (int32_t tile = first_tile; tile < num_tiles; tile += prm.tile_stride, ++lp) {
                const uint32_t acc = ((uint32_t)(2 * lp) & acc_mask) + which;
                const uint32_t acc_phase = ((uint32_t)(2 * lp) >> acc_shift) & 1u;
                ptx::mbar_wait_soft_u(&ctl->tmem_empty[acc], acc_phase ^ 1, tflag);
                ptx::tc_fence_after();
                if (leader && which == 0) trace_ev(prm, tracing, lp, EV_M_START);
                const uint32_t tmem_d = tmem_base + acc * bn;
                uint32_t accumulate = 0;
                for (int32_t cb = 0; cb < mma_outer; ++cb) {
                    if (!wready) ptx::mbar_wait_soft_u(&ctl->wfull[ws], wphase, tflag);
                    const uint32_t a_base = da_lo + ws * a_stage16 + which * win_sub16;
                    if (cb == 0 && leader && which == 0) trace_ev(prm, tracing, lp, EV_M_WIN);
                    int32_t j = 0;
                    for (int32_t st = 0; st < inner_stages; ++st) {
                        if (!ready) ptx::mbar_wait_soft_u(&ctl->full[stage], phase, tflag);
                        ptx::tc_fence_after();
                        if (st == 0 && cb == 0 && leader && which == 0) trace_ev(prm, tracing, lp, EV_M_FULL);
                        uint32_t nstage = stage + 1, nphase = phase;
                        if (nstage == ring_hi) { nstage = ring_lo; nphase ^= 1; }
                        const bool ready_next = ptx::mbar_test_u(&ctl->full[nstage], nphase);
                        uint32_t b_lo = db_lo + stage * b_stage16;
                        for (int32_t t = 0; t < tps; ++t) {
#pragma unroll
                            for (int k = 0; k < KS; ++k) {
                                ptx::mma_i8_ss_pred32(tmem_d, a_base + (uint32_t)prm.a_tab[j + k], da_hi, b_lo + 2u * k, db_hi, idesc,
                                                      accumulate, leader);
                                accumulate = 1;
                            }
                            j += KS;
                            b_lo += b_block16;
                        }
                        ptx::mma_commit_pred(&ctl->empty[stage], leader);
                        stage = nstage; phase = nphase; ready = ready_next;
                    }
                    ptx::mma_commit_pred(&ctl->wempty[ws], leader);
                    if (++ws == win_hi) { ws = win_lo; wphase ^= 1; }
                    wready = ptx::mbar_test_u(&ctl->wfull[ws], wphase);
                }
                ptx::mma_commit_pred(&ctl->tmem_full[acc], leader);
                if (leader && which == 0) trace_ev(prm, tracing, lp, EV_M_DONE);
            }
        } else {
        int32_t local = (int32_t)which;
        const int32_t tile_step = (int32_t)(n_mma * gridDim.x);
        // resident B with several N tiles: the tile's N index (fastest digit of the tile number), advanced without division
        const uint32_t tiles_n_u = (uint32_t)prm.tiles_n;
        const uint32_t nb_step = RESB ? (uint32_t)tile_step % tiles_n_u : 0u;
        uint32_t n_blk = RESB ? (blockIdx.x + which * gridDim.x) % tiles_n_u : 0u;
        // one N tile of the resident matrix (N-stationary: the only resident tile sits at offset 0)
        const uint32_t b_tile16 = prm.res_one ? 0u : b_block16 * (uint32_t)(prm.cblocks * prm.inner);
        const uint32_t fold_tile16 = prm.res_one ? 0u : (bn * 32u) >> 4;
        // (a pair leader counts its own tiles; the peer's tile of each step shares the MMAs)
        for (int32_t tile = active ? (int32_t)(blockIdx.x + which * gridDim.x) : num_tiles; tile < num_tiles;
             tile += tile_step, local += (int32_t)n_mma) {
            const uint32_t acc_stage = (uint32_t)local & acc_mask;
            const uint32_t acc_phase = ((uint32_t)local >> acc_shift) & 1u;
            ptx::mbar_wait_soft_u(&ctl->tmem_empty[acc_stage], acc_phase ^ 1, tflag);
            ptx::tc_fence_after();
            if (leader) trace_ev(prm, tracing, local, EV_M_START);
            const uint32_t tmem_d = tmem_base + acc_stage * bn;
            uint32_t accumulate = 0;
            const uint32_t idesc_t = (tail_on && tile >= prm.tail_first) ? idesc_half : idesc;     // split last round: N = bn / 2
            const uint32_t db_n = RESB ? db_lo + n_blk * b_tile16 : db_lo;
            if (MAYFOLD && fold) {      // D = ones-block x bias digits: the accumulator starts at the bias
                ptx::mma_i8_ss_pred32(tmem_d, fa_lo, fa_hi, fb_lo + n_blk * fold_tile16, fb_hi, idesc, 0u, leader);
                accumulate = 1;
            }
            if (RESB) { n_blk += nb_step; if (n_blk >= tiles_n_u) n_blk -= tiles_n_u; }
            if (kWindow && RESB) {
                // ---- window A, resident B: per channel chunk one wait, then a flat run of table-driven MMAs
                uint32_t b_base = db_n;
                for (int32_t cb = 0; cb < mma_outer; ++cb, b_base += b_chunk16) {
                    if (!wready) ptx::mbar_wait_soft_u(&ctl->wfull[ws], wphase, tflag);
                    ptx::tc_fence_after();
                    const uint32_t a_base = da_lo + ws * a_stage16;
                    if (cb == 0 && leader) trace_ev(prm, tracing, local, EV_M_WIN);
                    // unrolled by the length of a filter row's worth of MMAs so the table loads and descriptor adds of one
                    // batch overlap (the uniform datapath has long latencies and the loop's fixed cost is ~25 instructions)
                    constexpr int kBatch = (KM == 3) ? 8 : 9;
                    int32_t j = 0;
                    for (; j + kBatch <= n_tab; j += kBatch) {
#pragma unroll
                        for (int u = 0; u < kBatch; ++u) {
                            mma_issue<CTA2>(tmem_d, a_base + (uint32_t)prm.a_tab[j + u], da_hi, b_base + (uint32_t)prm.b_tab[j + u],
                                            db_hi, idesc, accumulate, leader);
                            accumulate = 1;
                        }
                    }
                    for (; j < n_tab; ++j) {
                        mma_issue<CTA2>(tmem_d, a_base + (uint32_t)prm.a_tab[j], da_hi, b_base + (uint32_t)prm.b_tab[j], db_hi,
                                        idesc, accumulate, leader);
                        accumulate = 1;
                    }
                    mma_commit<CTA2>(&ctl->wempty[ws], leader);
                    if (++ws == win_hi) { ws = win_lo; wphase ^= 1; }
                    wready = ptx::mbar_test_u(&ctl->wfull[ws], wphase);
                }
            } else {
                uint32_t b_res = db_n;    // resident B (ring modes): walks the blocks of the tile in order
                for (int32_t cb = 0; cb < mma_outer; ++cb) {
                    uint32_t a_base = da_lo;
                    if (kWindow) {
                        if (!wready) ptx::mbar_wait_soft_u(&ctl->wfull[ws], wphase, tflag);
                        a_base = da_lo + ws * a_stage16;
                        if (cb == 0 && leader) trace_ev(prm, tracing, local, EV_M_WIN);
                    }
                    int32_t j = 0;   // index into the chunk's A-offset table
                    for (int32_t st = 0; st < inner_stages; ++st) {
                        if (!ready) ptx::mbar_wait_soft_u(&ctl->full[stage], phase, tflag);
                        ptx::tc_fence_after();
                        if (st == 0 && cb == 0 && leader) trace_ev(prm, tracing, local, EV_M_FULL);
                        // probe the NEXT stage now: the (non-blocking) test's latency overlaps this stage's MMA issue
                        uint32_t nstage = stage + 1, nphase = phase;
                        if (nstage == ring_hi) { nstage = ring_lo; nphase ^= 1; }
                        const bool ready_next = ptx::mbar_test_u(&ctl->full[nstage], nphase);
                        if (!kWindow) a_base = da_lo + stage * a_stage16;
                        uint32_t b_lo = RESB ? b_res : db_lo + stage * b_stage16;
                        for (int32_t t = 0; t < tps; ++t) {
#pragma unroll
                            for (int k = 0; k < KS; ++k) {
                                const uint32_t a_lo = a_base + (uint32_t)prm.a_tab[kWindow ? j + k : k];
                                mma_issue<CTA2>(tmem_d, a_lo, da_hi, b_lo + 2u * k, db_hi, idesc_t, accumulate, leader);
                                accumulate = 1;
                            }
                            j += KS;
                            b_lo += b_block16;
                        }
                        if (RESB) b_res = b_lo;
                        mma_commit<CTA2>(&ctl->empty[stage], leader);     // slot reusable once these MMAs retire
                        stage = nstage; phase = nphase; ready = ready_next;
                    }
                    if (kWindow) {
                        mma_commit<CTA2>(&ctl->wempty[ws], leader);
                        if (++ws == win_hi) { ws = win_lo; wphase ^= 1; }
                        wready = ptx::mbar_test_u(&ctl->wfull[ws], wphase);
                    }
                }
            }
            mma_commit<CTA2>(&ctl->tmem_full[acc_stage], leader);     // accumulator complete -> epilogue (of both CTAs)
            if (leader) trace_ev(prm, tracing, local, EV_M_DONE);
        }
        }
    } else if (warp >= kFirstEpiWarp) {
        // ===================== epilogue: 16 warps in 2 teams of 8 or 4 teams of 4 =====================
        // A team owns every n_teams-th tile of this CTA (CTA-local tile L sits in TMEM accumulator stage L % n_acc) and a
        // ring of staging panels.  Wide N tiles use 2 x 8 warps (the two warp sets of a team split a panel's columns);
        // N tiles of <= 64 columns use 4 x 4 warps, so a warp converts a whole row of the tile and the per-tile fixed
        // cost (barrier, waits, address set-up) is paid half as often per output.
        // Per panel (<= 128 columns): drain TMEM + requantise into the swizzled staging panel, ONE team barrier, then
        // one thread issues the TMA store.  With 3 staging buffers no second barrier is needed: before the barrier of
        // panel p the issuer has waited until the store of panel p-2 has read its buffer, which is the buffer panel
        // p+1 will be written to.
        // Control flow is uniform across a team (named barriers): a watchdog trip only stops the waiting.
        const uint32_t e = warp - kFirstEpiWarp;                  // 0..15
        // (first thing in the role, while nothing else is live: its temporaries must not cost the tile loops registers)
        bool fold = false;
        if (MAYFOLD && prm.fold) {
            write_fold_blocks(smem + prm.off_fold, e * 32u + lane, prm.res_one ? prm.bn : prm.bn * prm.tiles_n,
                              prm.res_one ? (int32_t)(blockIdx.x % (uint32_t)prm.tiles_n) * prm.bn : 0, prm.k_out, prm.k_mod, bias,
                              &ctl->fold_ok);
            ptx::fence_proxy_async();          // generic-proxy writes -> visible to the tensor core's async-proxy reads
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&ctl->bias_ready);
            ptx::mbar_wait(&ctl->bias_ready, 0, tflag);
            fold = *reinterpret_cast<volatile uint32_t*>(&ctl->fold_ok) != 0;
        }
        const bool small_teams = prm.team_warps == 4;
        const uint32_t team = small_teams ? (e >> 2) : (e >> 3);
        const uint32_t tw = small_teams ? (e & 3) : (e & 7);      // warp inside the team
        const uint32_t n_teams = small_teams ? 4 : 2;
        const uint32_t team_threads = small_teams ? 128 : 256;
        const uint32_t quarter = warp & 3;                        // TMEM lanes [32*quarter, +32) for this warp
        const uint32_t half = tw >> 2;                            // which half of a panel's columns (8-warp teams)
        const uint32_t tt_id = (tw << 5) | lane;                  // thread inside the team
        const bool issuer = (tt_id == 0);
        const uint32_t bar_id = 1 + team;                         // named barrier of this team
        // the per-panel barrier, with immediate operands where the team shape is the common one (BAR.SYNC imm, imm)
        auto team_sync = [&]() {
            if (!small_teams) {
                if (team == 0) asm volatile("bar.sync 1, 256;" ::: "memory");
                else asm volatile("bar.sync 2, 256;" ::: "memory");
            } else {
                ptx::named_bar_sync(bar_id, team_threads);
            }
        };
        const bool int8_out = (prm.out_mode == LBC_OUT_INT8);
        const uint32_t panel_smem = (uint32_t)(kBlockM * prm.panel_bytes);
        const uint32_t nbufs = (uint32_t)prm.stage_bufs;
        const uint32_t team_staging_s = ptx::smem_u32(staging) + team * nbufs * panel_smem;   // 32-bit shared-window address
        const uint32_t tmem_lane_base = tmem_base + ((quarter * 32u) << 16);
        const uint32_t tmem_empty_s = ptx::smem_u32(&ctl->tmem_empty[0]);
        const uint32_t tmem_full_s = ptx::smem_u32(&ctl->tmem_full[0]);
        const uint32_t bn_u = (uint32_t)prm.bn;
        uint32_t sbuf = 0;
        float* sc = ctl->scale[team];
        int32_t* bi = ctl->bias[team];
        // Column-split epilogue (IgemmParams::epi_split): both teams walk every tile of the CTA; int8 mode deals the panels
        // of a tile to the teams (panel index = team, team + 2, ...), int32 mode gives each team one half of the columns.
        const bool split = prm.epi_split != 0;
        const int32_t i32_first = ((prm.bn / 16 + 1) / 2) * 16;                 // int32 + split: team 0 owns [0, first)
        const int32_t pbase0 = (split && !int8_out && team) ? i32_first : 0;    // first column of this team's range
        // columns of a panel handled by this warp: [pc_begin, pc_end), multiples of 16
        // (int32 mode: the team's whole column range is one "panel")
        const int32_t pcols = int8_out ? prm.panel_bytes : !split ? prm.bn : team ? prm.bn - i32_first : i32_first;
        const int32_t psplit = small_teams ? pcols : ((pcols / 16 + 1) / 2) * 16;
        const int32_t pc_begin = half ? psplit : 0;
        const int32_t pc_end = half ? pcols : psplit;
        const int32_t n_panels = int8_out ? prm.n_panels : 1;
        const int32_t p_first = (split && int8_out) ? (int32_t)team : 0;        // team path: this team's panels
        const int32_t p_step = (split && int8_out) ? (int32_t)n_teams : 1;

        const uint32_t lane_row = quarter * 32 + lane;            // TMEM lane == row of the MMA tile
        EpiThread et;
        if (kWindow) {
            et.wrow = (int32_t)lane_row / prm.wt;
            et.wcol = (int32_t)lane_row - et.wrow * prm.wt;
            et.valid = et.wcol < prm.cols_per_tile && et.wrow < prm.rows_per_tile;
            et.srow = (uint32_t)(et.wrow * prm.cols_per_tile + et.wcol);
        } else {
            et.wrow = et.wcol = 0;
            et.valid = true;
            et.srow = lane_row;
        }

        const uint32_t row_off = et.srow * (uint32_t)prm.panel_bytes;     // byte offset of this lane's staging row
        // per-thread XOR term of the staging swizzle (see epi_consume)
        const uint32_t swz_mask = ((row_off >> 7) & ((1u << prm.panel_swz_bits) - 1u)) << 4;
        const uint32_t acc_mask = (uint32_t)prm.n_acc - 1u, acc_shift = prm.n_acc == 8 ? 3u : prm.n_acc == 4 ? 2u : 1u;
        // The tile loops exist once per value of the (launch-wide) fold switch: inside a copy the switch is a compile-time
        // constant, so it costs no register in loops that already run at the 96-register cap.
        auto epilogue_tiles = [&](auto fold_tag) __attribute__((always_inline)) {
        constexpr bool kFold = decltype(fold_tag)::value;
        int32_t cur_nblk = -1;
        const uint32_t tmem_empty0 = CTA2 ? ptx::mapa(ptx::smem_u32(&ctl->tmem_empty[0]), 0) : 0u;   // the leader's barriers
        if (!LBC_TRACE && !kWindow && !CTA2 && prm.warp_store && !split && !small_teams && n_panels == 2 && prm.n_acc == 2) {
            // ---- lean loop for the hottest epilogue-bound shape: ring modes, 256-wide tiles, per-warp stores (the 1x1
            // channel expansions).  Everything a tile needs is a running register: two teams and two accumulator stages
            // mean team t ALWAYS drains stage t (only the phase bit flips), the warp's panel, TMEM address, staging buffers
            // and swizzle term never change, and the tile index advances by (n_blk, m) digit steps.  The general loop below
            // spent ~190 instructions per warp and tile on re-deriving these (ncu source counters, r02: as many samples as
            // the conversion loop itself); this one spends ~40.
            const uint32_t full_bar = tmem_full_s + team * 8u, empty_bar = tmem_empty_s + team * 8u;
            const uint32_t taddr = tmem_lane_base + team * bn_u;
            const int32_t pbase = (int32_t)half * pcols;                       // this warp's panel: columns [pbase, pbase + 128)
            const uint32_t wbuf0 = ptx::smem_u32(staging) + e * nbufs * (32u * 128u);
            const uint32_t wrow_off = lane * 128u, wswz = (lane & 7u) << 4;
            const int32_t tiles_n = prm.tiles_n, tiles_m = prm.tiles_m, s_nb = prm.tstep_nb, s_m = prm.tstep_img, rev = prm.rev_m;
            const int32_t bn = prm.bn, k_out = prm.k_out, row_in_tile = (int32_t)(quarter * 32u);
            const int32_t wait_mode = nbufs >= 2 ? 2 : 1;
            const bool relu = prm.relu != 0;
            EpiThread wt = et;
            wt.valid = true;
            uint32_t phase = 0, wbuf = wbuf0;
            int32_t n_blk, m;
            {
                const int32_t t0 = (int32_t)(blockIdx.x + team * gridDim.x);
                n_blk = t0 % tiles_n;
                m = t0 / tiles_n;
            }
            for (; m < tiles_m;) {
                const int32_t col0 = n_blk * bn;
                if (n_blk != cur_nblk) {       // per-channel parameters of this N tile -> smem (only when the N tile changes)
                    ptx::named_bar_sync(bar_id, team_threads);
                    for (int32_t c = (int32_t)tt_id; c < bn; c += (int32_t)team_threads) {
                        const int32_t kc = col0 + c;
                        const bool in = kc < k_out;
                        const int32_t kp = prm.k_mod ? kc % prm.k_mod : kc;
                        sc[c] = (in && scale) ? __ldg(scale + kp) : 0.0f;
                        bi[c] = (in && bias) ? __ldg(bias + kp) : 0;
                    }
                    cur_nblk = n_blk;
                    ptx::named_bar_sync(bar_id, team_threads);
                }
                ptx::mbar_wait_s(full_bar, phase, tflag);
                ptx::tc_fence_after();
                epi_run<kFold>(true, relu, kFold, prm, sc, bi, taddr, pbase, pbase, pbase + pcols, wt, wbuf, wrow_off, wswz, nullptr, -1,
                              col0, wait_mode);
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive_s(empty_bar);
                ptx::fence_proxy_async();
                __syncwarp();
                if (lane == 0) {
                    const int32_t cbyte = col0 + pbase;
                    const int32_t mm = (rev > 0) ? rev - 1 - m : m;
                    if (cbyte < k_out) ptx::tma_store_2d_s(&tm_out, wbuf, cbyte, mm * kBlockM + row_in_tile);
                    ptx::tma_store_commit();
                }
                if (nbufs >= 2) wbuf ^= (wbuf0 ^ (wbuf0 + 32u * 128u));      // toggle between the warp's two buffers
                phase ^= 1u;
                n_blk += s_nb;
                if (n_blk >= tiles_n) { n_blk -= tiles_n; ++m; }
                m += s_m;
            }
            // (a CTA only has to outlive the READS of its staging buffers: the writes of a bulk store are memory operations
            // of this grid like any other, complete before the grid does and visible after the next kernel's
            // griddepcontrol.wait - waiting for them here kept every CTA ~1 us longer on its SM at the end of every launch)
            if (lane == 0) ptx::tma_store_wait_read<0>();
        } else if (!CTA2 && prm.warp_store && small_teams && !split && prm.n_acc == 8) {
            // ---- narrow N tiles (one tile of 32 / 64 columns, 8 accumulator stages) with per-warp stores, all A modes.
            // A 128 x 64 tile is 8 KB of output against ~1000 cycles of per-team bookkeeping in the team paths below (two
            // named barriers, the issuer's store wait, the iterator: traces r02 show 2450 cycles per team and tile of which
            // 1300 convert) - the stems and every 64-channel layer ran at that pace, not at HBM's or the tensor pipe's.
            // Here nothing is shared inside a team: the warp converts its 32 TMEM lanes x bn columns into its OWN staging
            // buffer and stores them with its own TMA store; the team only shares the accumulator stage's barriers.
            // Four teams and eight stages: team t drains stage t and t + 4 alternately, the phase flips every second tile.
            // WINDOW tiles: the window pitch is a power of two >= 32 (the planner pads it), so a warp's 32 lanes are ONE
            // run of <= 32 consecutive output pixels of one tile row: box {bn, 32, 1, 1} (tm_out), or the ragged last run
            // of a row with box {bn, cols % 32, 1, 1} (tm_out2); warps whose lanes are all halo convert nothing.
            for (int32_t c = (int32_t)tt_id; c < prm.bn; c += (int32_t)team_threads) {
                const bool in = c < prm.k_out;
                const int32_t kp = prm.k_mod ? c % prm.k_mod : c;
                sc[c] = (in && scale) ? __ldg(scale + kp) : 0.0f;
                bi[c] = (in && bias) ? __ldg(bias + kp) : 0;
            }
            ptx::named_bar_sync(bar_id, team_threads);
            const uint32_t wbytes = 32u * (uint32_t)prm.panel_bytes;
            const uint32_t wbuf0 = ptx::smem_u32(staging) + e * nbufs * wbytes;
            const uint32_t wrow_off = lane * (uint32_t)prm.panel_bytes;
            const uint32_t wswz = ((wrow_off >> 7) & ((1u << prm.panel_swz_bits) - 1u)) << 4;
            const int32_t wait_mode = nbufs >= 2 ? 2 : 1;
            const bool relu = prm.relu != 0;
            const int32_t lane0_row = (int32_t)(quarter * 32u);
            // this warp's run of output pixels inside a window tile
            int32_t wr = 0, wc0 = 0, vc = 32;
            if (kWindow) {
                wr = lane0_row / prm.wt;
                wc0 = lane0_row - wr * prm.wt;
                vc = min(32, prm.cols_per_tile - wc0);
            }
            const bool has_rows = !kWindow || (vc > 0 && wr < prm.rows_per_tile);
            const bool ragged = vc < 32;
            EpiThread wt_ = et;
            wt_.valid = true;
            uint32_t acc = team, phase = 0, wbuf = wbuf0;
            Iter it;
            it.init(prm, (int32_t)(blockIdx.x + team * gridDim.x), n_major, true);
            for (; it.tile < num_tiles; it.next(prm)) {
                ptx::mbar_wait_s(tmem_full_s + acc * 8u, phase, tflag);
                ptx::tc_fence_after();
                if (issuer) trace_ev(prm, tracing, it.local * 4 + (int32_t)team, EV_E_START);
                if (has_rows)
                    epi_run<kFold>(true, relu, kFold, prm, sc, bi, tmem_lane_base + acc * bn_u, 0, 0, pcols, wt_, wbuf, wrow_off, wswz,
                                  nullptr, -1, 0, wait_mode);
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive_s(tmem_empty_s + acc * 8u);
                if (issuer) trace_ev(prm, tracing, it.local * 4 + (int32_t)team, EV_E_DRAINED);
                if (has_rows) {
                    ptx::fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) {
                        if (kWindow) {
                            const int32_t pp = it.p0(prm) + wr, qq = it.q0(prm) + wc0, img = it.image();
                            if (img < prm.n_img && pp < prm.p) {
                                if (ragged) ptx::tma_store_4d_s(&tm_out2, wbuf, 0, qq, pp, img);
                                else ptx::tma_store_4d_s(&tm_out, wbuf, 0, qq, pp, img);
                            }
                        } else {
                            ptx::tma_store_2d_s(&tm_out, wbuf, 0, it.m0() + lane0_row);
                        }
                        ptx::tma_store_commit();
                    }
                    if (nbufs >= 2) wbuf ^= (wbuf0 ^ (wbuf0 + wbytes));      // toggle between the warp's two buffers
                }
                if (issuer) trace_ev(prm, tracing, it.local * 4 + (int32_t)team, EV_E_STORED);
                acc ^= 4u;
                if (acc == team) phase ^= 1u;
            }
            if (lane == 0) ptx::tma_store_wait_read<0>();
        } else if (!CTA2 && prm.tpi == 2) {
            // ---- narrow N tiles (<= 64 columns, one N tile, 8 accumulator stages): a 4-warp team takes TWO consecutive
            // CTA-local tiles per iteration - adjacent TMEM stages, one staging panel each - so the per-iteration
            // bookkeeping (iterator, barriers, waits, store issue), which is a third of this role's instructions on
            // 64-column tiles, is paid once per two tiles.
            Iter ia, ib;
            ia.init(prm, (int32_t)(blockIdx.x + (2u * team) * gridDim.x), n_major, true);
            ib.init(prm, (int32_t)(blockIdx.x + (2u * team + 1u) * gridDim.x), n_major, true);
            // per-channel parameters: a single N tile, loaded once
            for (int32_t c = (int32_t)tt_id; c < prm.bn; c += (int32_t)team_threads) {
                const bool in = c < prm.k_out;
                const int32_t kp = prm.k_mod ? c % prm.k_mod : c;
                sc[c] = (in && scale) ? __ldg(scale + kp) : 0.0f;
                bi[c] = (in && bias) ? __ldg(bias + kp) : 0;
            }
            ptx::named_bar_sync(bar_id, team_threads);
            int32_t* y32 = reinterpret_cast<int32_t*>(y);
            for (; ia.tile < num_tiles; ia.next(prm), ib.next(prm)) {
                const int32_t tile0 = ia.local * 8 + 2 * (int32_t)team;        // CTA-local index of the first tile
                const uint32_t acc0 = (uint32_t)tile0 & 7u, acc_phase = ((uint32_t)tile0 >> 3) & 1u;
                const bool have_b = ib.tile < num_tiles;
#pragma unroll 1
                for (uint32_t sub = 0; sub < 2; ++sub) {
                    if (sub == 1 && !have_b) break;
                    const Iter& it2 = sub ? ib : ia;
                    const uint32_t acc = acc0 + sub;
                    ptx::mbar_wait_s(tmem_full_s + acc * 8u, acc_phase, tflag);
                    ptx::tc_fence_after();
                    if (issuer) trace_ev(prm, tracing, tile0 + (int32_t)sub, EV_E_START);
                    int64_t out_row = -1;
                    if (!int8_out) {
                        if (kWindow) {
                            const int32_t pp = it2.p0(prm) + et.wrow, qq = it2.q0(prm) + et.wcol;
                            if (et.valid && pp < prm.p && qq < prm.q && it2.image() < prm.n_img) out_row = ((int64_t)it2.image() * prm.p + pp) * prm.q + qq;
                        } else {
                            const int64_t r = (int64_t)it2.m0() + lane_row;
                            if (r < prm.m_total) out_row = r;
                        }
                    }
                    const uint32_t taddr = tmem_lane_base + acc * bn_u;
                    const uint32_t staging_s = team_staging_s + sbuf * panel_smem;
                    if (nbufs == 1 && int8_out) {
                        if (issuer) ptx::tma_store_wait_read<0>();
                        ptx::named_bar_sync(bar_id, team_threads);
                    }
                    epi_run<kFold>(int8_out, prm.relu != 0, kFold, prm, sc, bi, taddr, 0, 0, pcols, et, staging_s, row_off, swz_mask, y32,
                                  out_row, 0);
                    ptx::tc_fence_before();
                    __syncwarp();
                    if (lane == 0) ptx::mbar_arrive_s(tmem_empty_s + acc * 8u);
                    if (issuer) trace_ev(prm, tracing, tile0 + (int32_t)sub, EV_E_DRAINED);
                    if (int8_out) {   // same staging-ring protocol as the general path below: one barrier per panel
                        ptx::fence_proxy_async();
                        if (issuer) {
                            if (nbufs >= 3) ptx::tma_store_wait_read<1>();
                            else if (nbufs == 2) ptx::tma_store_wait_read<0>();
                        }
                        ptx::named_bar_sync(bar_id, team_threads);
                        if (issuer) {
                            if (kWindow) {
                                if (it2.image() < prm.n_img) ptx::tma_store_4d_s(&tm_out, staging_s, 0, it2.q0(prm), it2.p0(prm), it2.image());
                            } else {
                                ptx::tma_store_2d_s(&tm_out, staging_s, 0, it2.m0());
                            }
                            ptx::tma_store_commit();
                            trace_ev(prm, tracing, tile0 + (int32_t)sub, EV_E_STORED);
                        }
                        if (++sbuf == nbufs) sbuf = 0;
                    }
                }
            }
            if (issuer && int8_out) ptx::tma_store_wait_read<0>();
        } else {
        Iter it;
        // pair mode: team t takes the t-th tile of every pair (two teams); otherwise every n_teams-th tile
        // column-split: every team takes every tile of the CTA
        it.init(prm, pair_mode ? first_tile + (int32_t)team : split ? (int32_t)blockIdx.x : (int32_t)(blockIdx.x + team * gridDim.x),
                n_major, !pair_mode && !split);
        for (; it.tile < num_tiles;) {
            // CTA-local tile index (it.local counts pairs / team steps / plain CTA steps)
            const int32_t tile = pair_mode ? 2 * it.local + (int32_t)team : split ? it.local : it.local * (int32_t)n_teams + (int32_t)team;
            const bool tail = tail_on && it.tile >= prm.tail_first;        // this CTA's half-tile of the split last round
            struct { int32_t n_blk, img, p0, q0, m0; } tc = {tail ? 0x4000 + tail_half : it.n_blk, it.image(), it.p0(prm), it.q0(prm),
                                                             tail ? tail_row0 : it.m0()};
            const int32_t col0 = tail ? tail_half * (prm.bn >> 1) : tc.n_blk * prm.bn;
            const int32_t n_panels_t = tail ? 1 : n_panels;                // a half-tile is one 128-byte panel
            // per-channel parameters of this N tile -> smem (only when the N tile changes)
            if (tc.n_blk != cur_nblk) {
                ptx::named_bar_sync(bar_id, team_threads);       // everyone done with the previous parameters
                for (int32_t c = (int32_t)tt_id; c < prm.bn; c += (int32_t)team_threads) {
                    const int32_t kc = col0 + c;
                    const bool in = kc < prm.k_out;
                    const int32_t kp = prm.k_mod ? kc % prm.k_mod : kc;
                    sc[c] = (in && scale) ? __ldg(scale + kp) : 0.0f;
                    bi[c] = (in && bias) ? __ldg(bias + kp) : 0;
                }
                cur_nblk = tc.n_blk;
                ptx::named_bar_sync(bar_id, team_threads);       // parameters visible to the whole team
            }
            // accumulator ready?  CTA-local tile L lives in TMEM stage L % n_acc, on that stage's (L / n_acc)-th use
            const uint32_t acc = (uint32_t)tile & acc_mask;
            const uint32_t acc_phase = ((uint32_t)tile >> acc_shift) & 1u;
            ptx::mbar_wait_s(tmem_full_s + acc * 8u, acc_phase, tflag);
            ptx::tc_fence_after();
            if (issuer) trace_ev(prm, tracing, tile, EV_E_START);

            int64_t out_row = -1;   // int32 mode: global output row of this lane
            if (!int8_out) {
                if (kWindow) {
                    const int32_t pp = tc.p0 + et.wrow, qq = tc.q0 + et.wcol;
                    if (et.valid && pp < prm.p && qq < prm.q && tc.img < prm.n_img) out_row = ((int64_t)tc.img * prm.p + pp) * prm.q + qq;
                } else {
                    const int64_t r = (int64_t)tc.m0 + lane_row;
                    if (r < prm.m_total) out_row = r;
                }
            }
            const uint32_t taddr = tmem_lane_base + acc * bn_u;
            int32_t* y32 = reinterpret_cast<int32_t*>(y);

            if (prm.warp_store) {
                // ---- ring modes, int8 output: no team barrier at all.  The warp converts whole panels (its 32 rows x
                // <= 128 columns, panels dealt round-robin to the two warp sets of an 8-warp team) into its OWN swizzled
                // staging buffer and stores them with its own TMA store; lane 0 tracks the buffer through its bulk group.
                // stage_bufs staging buffers per warp (2 where shared memory allows): with one, the wait below exposes the
                // smem-read latency of the store issued a moment ago on every panel; with two it waits for the store before
                const uint32_t wbytes = 32u * (uint32_t)prm.panel_bytes;
                const uint32_t wbuf0 = ptx::smem_u32(staging) + e * nbufs * wbytes;
                const uint32_t wrow_off = lane * (uint32_t)prm.panel_bytes;
                const uint32_t wswz = ((wrow_off >> 7) & ((1u << prm.panel_swz_bits) - 1u)) << 4;
                const int32_t n_halves = small_teams ? 1 : 2;
                // panels dealt to the warp sets of the team, or (column-split) to the warp sets of both teams
                const int32_t w_first = split ? (int32_t)team * n_halves + (int32_t)half : (int32_t)half;
                const int32_t w_step = split ? (int32_t)n_teams * n_halves : n_halves;
                EpiThread wt = et;
                wt.valid = true;
                for (int32_t pnl = w_first; pnl < n_panels_t; pnl += w_step) {
                    const int32_t pbase = pnl * pcols;
                    const uint32_t wbuf = wbuf0 + sbuf * wbytes;
                    // (the wait for the buffer's last store to have read it out happens inside, before the first smem store)
                    epi_run<kFold>(true, prm.relu != 0, kFold, prm, sc, bi, taddr, pbase, pbase, pbase + pcols, wt, wbuf, wrow_off, wswz,
                                  y32, -1, col0, nbufs >= 2 ? 2 : 1);
                    if (pnl + w_step >= n_panels_t) {
                        ptx::tc_fence_before();
                        __syncwarp();
                        if (lane == 0) {
                            if (CTA2) ptx::mbar_arrive_cluster(tmem_empty0 + acc * 8u);
                            else ptx::mbar_arrive_s(tmem_empty_s + acc * 8u);
                        }
                        if (issuer) trace_ev(prm, tracing, tile, EV_E_DRAINED);
                    }
                    ptx::fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) {
                        const int32_t cbyte = col0 + pbase;
                        if (cbyte < prm.k_out) ptx::tma_store_2d_s(&tm_out, wbuf, cbyte, tc.m0 + (int32_t)(quarter * 32u));
                        ptx::tma_store_commit();
                    }
                    if (++sbuf == nbufs) sbuf = 0;
                }
                if (issuer) trace_ev(prm, tracing, tile, EV_E_STORED);
            } else
            for (int32_t pnl = p_first; pnl < n_panels_t; pnl += p_step) {
                const int32_t pbase = pnl * pcols + pbase0;
                if (nbufs == 1 && int8_out) {
                    // a single staging panel (shared memory is tight): its previous store must have read it out
                    if (issuer) ptx::tma_store_wait_read<0>();
                    team_sync();
                }
                const uint32_t staging_s = team_staging_s + sbuf * panel_smem;
                epi_run<kFold>(int8_out, prm.relu != 0, kFold, prm, sc, bi, taddr, pbase, pbase + pc_begin, pbase + pc_end, et, staging_s,
                              row_off, swz_mask, y32, out_row, col0);
                if (pnl + p_step >= n_panels_t) {
                    // accumulator drained: hand the TMEM stage back to the MMA warp
                    ptx::tc_fence_before();
                    __syncwarp();
                    if (lane == 0) {
                        if (CTA2) ptx::mbar_arrive_cluster(tmem_empty0 + acc * 8u);
                        else ptx::mbar_arrive_s(tmem_empty_s + acc * 8u);
                    }
                    if (issuer) trace_ev(prm, tracing, tile, EV_E_DRAINED);
                }
                if (int8_out) {
                    ptx::fence_proxy_async();
                    // the buffer the NEXT panel goes to must have been read out by its last store before anyone
                    // passes the barrier below: with nbufs buffers that is the store issued nbufs-1 panels ago
                    if (issuer) {
                        if (nbufs >= 3) ptx::tma_store_wait_read<1>();
                        else if (nbufs == 2) ptx::tma_store_wait_read<0>();
                    }
                    team_sync();
                    if (issuer) {
                        const int32_t cbyte = col0 + pbase;
                        if (cbyte < prm.k_out) {
                            if (kWindow) {
                                if (tc.img < prm.n_img)   // the pair modes pad the tile space with dummy tiles
                                    ptx::tma_store_4d_s(&tm_out, staging_s, cbyte, tc.q0, tc.p0, tc.img);
                            }
                            else
                                ptx::tma_store_2d_s(&tm_out, staging_s, cbyte, tc.m0);
                        }
                        ptx::tma_store_commit();
                        if (pnl + p_step >= n_panels_t) trace_ev(prm, tracing, tile, EV_E_STORED);
                    }
                    if (++sbuf == nbufs) sbuf = 0;
                }
            }
            it.next(prm);
        }
        if ((issuer || (prm.warp_store && lane == 0)) && int8_out) ptx::tma_store_wait_read<0>();
        }
        };   // epilogue_tiles
        if (MAYFOLD && fold) epilogue_tiles(std::true_type{});
        else epilogue_tiles(std::false_type{});
    }

    // ---- teardown ----
    ptx::tc_fence_before();
    if (CTA2) ptx::cluster_sync();     // neither CTA may leave while its peer still reads its operands / signals its barriers
    else __syncthreads();
    // trace mode: every CTA also leaves its own start / end time (its SM's cycle counter) behind the per-tile stamps
    if (LBC_TRACE && prm.trace != nullptr && threadIdx.x == 0) {
        long long* cta_times = prm.trace + (size_t)prm.trace_tiles * 16;
        cta_times[2 * blockIdx.x + 1] = clock64();
        gt[3] = (long long)ptx::globaltimer_ns();
    }
    if (warp == 1) {
        ptx::tc_fence_after();
        if (CTA2) ptx::tmem_dealloc_2cta(tmem_base, prm.tmem_cols);
        else ptx::tmem_dealloc(tmem_base, prm.tmem_cols);
    }
}

// =====================================================================================================
// Fused bottleneck tail (first slice): an R x S stride-1 convolution with 64 output channels (window A,
// resident filter) whose ReLU'd int8 result never leaves the SM - the epilogue writes it as a K-major, 64-byte-swizzled
// operand tile into shared memory and a second MMA warp multiplies it by the resident [256][64] filter of the following
// 1x1 convolution.  Only the 4x wider result goes to HBM: one write and one read of the middle tensor per bottleneck less
// (ResNet-50 stage 1: 2 x 103 MB per block), one launch less, and the second convolution's MMAs (256 tensor cycles per
// tile) run on a pipe that the first one leaves half idle.  The chain being replaced: conv -> relu -> conv of
// python/tmp.py:43-56, run as two launches.
//
//   TMEM     D1: 4 stages x 64 columns [0, 256)      D2: one stage of 256 columns [256, 512)
//   warp 0   loads both filter matrices once          warp 2   window producer (one halo window per tile and channel chunk)
//   warp 1   MMA1: tile L -> D1[L & 3]                warp 3   MMA2: A2[L & 1] (smem) x B2 -> D2
//   warps 4-19 (one group of 16):  per step  epi1(L + 1): D1 -> bias1/scale1/ReLU -> int8 -> A2[(L + 1) & 1]
//                                            epi2(L):     D2 -> bias2/scale2/[ReLU] -> int8 -> staging -> TMA store
//   so MMA2(L + 1) (which needs A2 written and D2 drained) runs under epi1(L + 2): its latency is hidden.
// The halo rows of a window tile flow through both GEMMs as garbage rows and are dropped by the epilogue's row compaction,
// exactly as in the unfused window kernel.
struct FusedParams {
    IgemmParams a;                 // the first convolution (window mode, resident filter, one N tile of 64 columns)
    int32_t k2, relu2;             // second convolution: output channels (256), ReLU
    int32_t c_mid;                 // channels between the two (== a.bn == 64)
    int32_t stage_bufs2;           // staging panels per team for the final output
    uint32_t off_b2, off_a2, off_stage2, off_ctl2, b1_bytes, b2_bytes;
};

struct FusedCtl {
    uint64_t wfull[kMaxWinStages];
    uint64_t wempty[kMaxWinStages];
    uint64_t d1_full[4], d1_empty[4];
    uint64_t a2_full[2], a2_empty[2];
    uint64_t d2_full, d2_empty;
    uint64_t bfull;
    uint32_t tmem_base;
    alignas(16) float scale1[64];
    alignas(16) int32_t bias1[64];
    alignas(16) float scale2[256];
    alignas(16) int32_t bias2[256];
};

__global__ void __launch_bounds__(kNumThreads, 1)
igemm_fused_tail_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b1,
                        const __grid_constant__ CUtensorMap tm_b2, const __grid_constant__ CUtensorMap tm_out, const FusedParams fp,
                        const int32_t* __restrict__ bias1, const float* __restrict__ scale1, const int32_t* __restrict__ bias2,
                        const float* __restrict__ scale2)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    const IgemmParams& prm = fp.a;
    uint8_t* smem_a = smem;                                   // window ring
    uint8_t* smem_b1 = smem + prm.off_b;
    uint8_t* smem_b2 = smem + fp.off_b2;
    uint8_t* smem_a2 = smem + fp.off_a2;                      // two 128 x 64 B operand tiles for the second GEMM
    uint8_t* staging = smem + fp.off_stage2;
    FusedCtl* ctl = reinterpret_cast<FusedCtl*>(smem + fp.off_ctl2);
    const uint32_t warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    const uint32_t lane = threadIdx.x & 31;
    volatile int* tflag = prm.flag;
    using Iter = TileIter<true>;
    constexpr uint32_t kA2Bytes = kBlockM * 64u;
    constexpr uint32_t kD2Col = 256u;

    if ((ptx::smem_u32(smem) & 1023u) != 0) {
        if (threadIdx.x == 0) *tflag = 2;
        return;
    }
    ptx::griddep_launch_dependents();
    if (warp == 0 && lane == 0) {
        ptx::prefetch_tensormap(&tm_a);
        ptx::prefetch_tensormap(&tm_b1);
        ptx::prefetch_tensormap(&tm_b2);
        ptx::prefetch_tensormap(&tm_out);
        for (int i = 0; i < prm.win_stages; ++i) {
            ptx::mbar_init(&ctl->wfull[i], 1);
            ptx::mbar_init(&ctl->wempty[i], 1);
        }
        for (int i = 0; i < 4; ++i) {
            ptx::mbar_init(&ctl->d1_full[i], 1);
            ptx::mbar_init(&ctl->d1_empty[i], kEpiWarps);
        }
        for (int i = 0; i < 2; ++i) {
            ptx::mbar_init(&ctl->a2_full[i], kEpiWarps);
            ptx::mbar_init(&ctl->a2_empty[i], 1);
        }
        ptx::mbar_init(&ctl->d2_full, 1);
        ptx::mbar_init(&ctl->d2_empty, kEpiWarps);
        ptx::mbar_init(&ctl->bfull, 1);
        ptx::fence_barrier_init();
    }
    if (warp == 1) {
        ptx::tmem_alloc(&ctl->tmem_base, 512);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    ptx::griddep_wait();
    const uint32_t tmem_base = ctl->tmem_base;
    const int32_t num_tiles = prm.tiles_m;                    // one N tile

    if (warp == 0) {
        // ---- both filter matrices, once
        if (ptx::elect_one()) {
            ptx::mbar_expect_tx(&ctl->bfull, fp.b1_bytes + fp.b2_bytes);
            const int32_t nblk = prm.cblocks * prm.inner;
            uint8_t* dst = smem_b1;
            int32_t bcol = 0;
            for (int32_t i = 0; i < nblk; ++i, dst += prm.b_block_bytes, bcol += prm.bkb) ptx::tma_load_2d(dst, &tm_b1, &ctl->bfull, bcol, 0);
            ptx::tma_load_2d(smem_b2, &tm_b2, &ctl->bfull, 0, 0);
        }
        __syncwarp();
    } else if (warp == 2) {
        // ---- window producer
        const bool leader = ptx::elect_one();
        const int32_t pad_w = prm.pad_w, pad_h = prm.pad_h, cblocks = prm.cblocks, bkc = prm.bkc;
        const uint32_t win_stages = (uint32_t)prm.win_stages, win_stage_bytes = prm.win_stage_bytes, win_tx = prm.win_tx_bytes;
        uint32_t ws = 0, wphase = 0;
        bool ok = true;
        Iter it;
        for (it.init(prm, (int32_t)blockIdx.x, false); it.tile < num_tiles && ok; it.next(prm)) {
            const int32_t wq = it.q0(prm) - pad_w, wp = it.p0(prm) - pad_h;
            int32_t c0 = 0;
            for (int32_t cb = 0; cb < cblocks; ++cb, c0 += bkc) {
                ok = wait_or_quit(&ctl->wempty[ws], wphase ^ 1, tflag);
                if (!ok) break;
                if (leader) {
                    ptx::mbar_expect_tx(&ctl->wfull[ws], win_tx);
                    ptx::tma_load_4d(smem_a + ws * win_stage_bytes, &tm_a, &ctl->wfull[ws], c0, wq, wp, it.image());
                }
                if (++ws == win_stages) { ws = 0; wphase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ---- MMA1: the R x S convolution, tile L -> D1[L & 3]
        const uint32_t leader = ptx::elect_one() ? 1u : 0u;
        const uint32_t idesc = ptx::make_idesc_i8(kBlockM, (uint32_t)prm.bn);
        const uint64_t db_base = ptx::make_kmajor_desc(ptx::smem_u32(smem_b1), (uint32_t)prm.bkb);
        const uint64_t da_base = ptx::make_kmajor_desc(ptx::smem_u32(smem_a), (uint32_t)prm.bkc);
        const uint32_t da_lo = (uint32_t)da_base, da_hi = (uint32_t)(da_base >> 32);
        const uint32_t db_lo = (uint32_t)db_base, db_hi = (uint32_t)(db_base >> 32);
        const uint32_t a_stage16 = prm.win_stage_bytes >> 4;
        const uint32_t b_chunk16 = (prm.b_block_bytes >> 4) * (uint32_t)prm.inner;
        const int32_t cblocks = prm.cblocks, n_tab = prm.n_tab;
        const uint32_t win_stages = (uint32_t)prm.win_stages, bn = (uint32_t)prm.bn;
        uint32_t ws = 0, wphase = 0;
        ptx::mbar_wait_soft_u(&ctl->bfull, 0, tflag);
        int32_t local = 0;
        for (int32_t tile = (int32_t)blockIdx.x; tile < num_tiles; tile += (int32_t)gridDim.x, ++local) {
            const uint32_t st = (uint32_t)local & 3u, ph = ((uint32_t)local >> 2) & 1u;
            ptx::mbar_wait_soft_u(&ctl->d1_empty[st], ph ^ 1, tflag);
            ptx::tc_fence_after();
            const uint32_t tmem_d = tmem_base + st * bn;
            uint32_t accumulate = 0, b_base = db_lo;
            for (int32_t cb = 0; cb < cblocks; ++cb, b_base += b_chunk16) {
                ptx::mbar_wait_soft_u(&ctl->wfull[ws], wphase, tflag);
                ptx::tc_fence_after();
                const uint32_t a_base = da_lo + ws * a_stage16;
                for (int32_t j = 0; j < n_tab; ++j) {
                    ptx::mma_i8_ss_pred32(tmem_d, a_base + (uint32_t)prm.a_tab[j], da_hi, b_base + (uint32_t)prm.b_tab[j], db_hi, idesc,
                                          accumulate, leader);
                    accumulate = 1;
                }
                ptx::mma_commit_pred(&ctl->wempty[ws], leader);
                if (++ws == win_stages) { ws = 0; wphase ^= 1; }
            }
            ptx::mma_commit_pred(&ctl->d1_full[st], leader);
        }
    } else if (warp == 3) {
        // ---- MMA2: the 1x1 convolution on the tile the epilogue just requantised, A2[L & 1] x B2 -> D2
        const uint32_t leader = ptx::elect_one() ? 1u : 0u;
        const uint32_t idesc = ptx::make_idesc_i8(kBlockM, (uint32_t)fp.k2);
        const uint64_t da = ptx::make_kmajor_desc(ptx::smem_u32(smem_a2), 64u);
        const uint64_t db = ptx::make_kmajor_desc(ptx::smem_u32(smem_b2), 64u);
        const uint32_t da_lo = (uint32_t)da, da_hi = (uint32_t)(da >> 32), db_lo = (uint32_t)db, db_hi = (uint32_t)(db >> 32);
        const uint32_t ksteps = (uint32_t)fp.c_mid / 32u;
        ptx::mbar_wait_soft_u(&ctl->bfull, 0, tflag);
        int32_t local = 0;
        for (int32_t tile = (int32_t)blockIdx.x; tile < num_tiles; tile += (int32_t)gridDim.x, ++local) {
            const uint32_t buf = (uint32_t)local & 1u;
            ptx::mbar_wait_soft_u(&ctl->a2_full[buf], ((uint32_t)local >> 1) & 1u, tflag);
            ptx::mbar_wait_soft_u(&ctl->d2_empty, ((uint32_t)local & 1u) ^ 1u, tflag);
            ptx::tc_fence_after();
            const uint32_t a_lo = da_lo + buf * (kA2Bytes >> 4);
            for (uint32_t k = 0; k < ksteps; ++k)
                ptx::mma_i8_ss_pred32(tmem_base + kD2Col, a_lo + 2u * k, da_hi, db_lo + 2u * k, db_hi, idesc, k ? 1u : 0u, leader);
            ptx::mma_commit_pred(&ctl->d2_full, leader);
            ptx::mma_commit_pred(&ctl->a2_empty[buf], leader);
        }
    } else if (warp >= kFirstEpiWarp) {
        // ---- epilogue: one group of 16 warps
        const uint32_t e = warp - kFirstEpiWarp;
        const uint32_t quarter = warp & 3;                        // TMEM lanes [32*quarter, +32)
        const uint32_t cg = e >> 2;                               // epi1: 16 columns each; epi2: team = cg >> 1, half = cg & 1
        const uint32_t team = cg >> 1, half = cg & 1u;
        const uint32_t tid = e * 32u + lane;
        const bool issuer = ((e & 7u) == 0u) && lane == 0;        // first thread of each team
        for (uint32_t c = tid; c < 64u; c += kEpiWarps * 32u) {
            ctl->scale1[c] = (int32_t)c < prm.k_out ? __ldg(scale1 + c) : 0.0f;
            ctl->bias1[c] = ((int32_t)c < prm.k_out && bias1) ? __ldg(bias1 + c) : 0;
        }
        for (uint32_t c = tid; c < 256u; c += kEpiWarps * 32u) {
            ctl->scale2[c] = (int32_t)c < fp.k2 ? __ldg(scale2 + c) : 0.0f;
            ctl->bias2[c] = ((int32_t)c < fp.k2 && bias2) ? __ldg(bias2 + c) : 0;
        }
        ptx::named_bar_sync(1, kEpiWarps * 32);
        const uint32_t lane_row = quarter * 32 + lane;
        EpiThread et;                                             // window compaction of the final output (as the unfused kernel)
        et.wrow = (int32_t)lane_row / prm.wt;
        et.wcol = (int32_t)lane_row - et.wrow * prm.wt;
        et.valid = et.wcol < prm.cols_per_tile && et.wrow < prm.rows_per_tile;
        et.srow = (uint32_t)(et.wrow * prm.cols_per_tile + et.wcol);
        EpiThread e1;                                             // the operand tile keeps EVERY MMA row, halo rows included
        e1.wrow = e1.wcol = 0; e1.valid = true; e1.srow = lane_row;
        const uint32_t tmem_lane_base = tmem_base + ((quarter * 32u) << 16);
        const uint32_t a2_row_off = lane_row * 64u, a2_swz = ((a2_row_off >> 7) & 3u) << 4;
        const uint32_t a2_s = ptx::smem_u32(smem_a2);
        const uint32_t row_off = et.srow * 128u, swz = ((row_off >> 7) & 7u) << 4;
        const uint32_t nbufs = (uint32_t)fp.stage_bufs2, panel_smem = kBlockM * 128u;
        const uint32_t team_staging_s = ptx::smem_u32(staging) + team * nbufs * panel_smem;
        const bool relu1 = prm.relu != 0, relu2 = fp.relu2 != 0;
        const uint32_t bn1 = (uint32_t)prm.bn;
        uint32_t sbuf = 0;

        // D1[T & 3] -> int8 operand tile A2[T & 1]
        auto epi1 = [&](int32_t T) {
            const uint32_t st = (uint32_t)T & 3u, buf = (uint32_t)T & 1u;
            ptx::mbar_wait(&ctl->d1_full[st], ((uint32_t)T >> 2) & 1u, tflag);
            ptx::mbar_wait(&ctl->a2_empty[buf], (((uint32_t)T >> 1) & 1u) ^ 1u, tflag);     // MMA2(T - 2) has read this buffer
            ptx::tc_fence_after();
            epi_run<false>(true, relu1, false, prm, ctl->scale1, ctl->bias1, tmem_lane_base + st * bn1, 0, (int32_t)(16u * cg),
                           (int32_t)(16u * cg + 16u), e1, a2_s + buf * kA2Bytes, a2_row_off, a2_swz, nullptr, -1, 0);
            ptx::tc_fence_before();
            ptx::fence_proxy_async();          // generic-proxy stores -> visible to the tensor core's async-proxy reads
            __syncwarp();
            if (lane == 0) {
                ptx::mbar_arrive(&ctl->a2_full[buf]);
                ptx::mbar_arrive(&ctl->d1_empty[st]);
            }
        };
        // D2 -> int8 NHWC output: team t stages and stores panel t (128 columns), its two warp sets 64 columns each
        auto epi2 = [&](int32_t T, int32_t img, int32_t p0, int32_t q0) {
            ptx::mbar_wait(&ctl->d2_full, (uint32_t)T & 1u, tflag);
            ptx::tc_fence_after();
            const uint32_t staging_s = team_staging_s + sbuf * panel_smem;
            const int32_t pbase = (int32_t)(team * 128u);
            if (nbufs == 1) {
                if (issuer) ptx::tma_store_wait_read<0>();
                if (team == 0) asm volatile("bar.sync 2, 256;" ::: "memory");
                else asm volatile("bar.sync 3, 256;" ::: "memory");
            }
            epi_run<false>(true, relu2, false, prm, ctl->scale2, ctl->bias2, tmem_lane_base + kD2Col, pbase, pbase + (int32_t)(half * 64u),
                           pbase + (int32_t)(half * 64u) + 64, et, staging_s, row_off, swz, nullptr, -1, 0);
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&ctl->d2_empty);
            ptx::fence_proxy_async();
            if (issuer && nbufs >= 2) ptx::tma_store_wait_read<0>();      // the buffer the next tile goes to has been read out
            if (team == 0) asm volatile("bar.sync 2, 256;" ::: "memory");
            else asm volatile("bar.sync 3, 256;" ::: "memory");
            if (issuer) {
                if (img < prm.n_img && pbase < fp.k2) ptx::tma_store_4d_s(&tm_out, staging_s, pbase, q0, p0, img);
                ptx::tma_store_commit();
            }
            if (++sbuf == nbufs) sbuf = 0;
        };
        Iter it;
        it.init(prm, (int32_t)blockIdx.x, false);
        int32_t L = 0;
        if (it.tile < num_tiles) epi1(0);
        while (it.tile < num_tiles) {
            const int32_t img = it.image(), p0 = it.p0(prm), q0 = it.q0(prm);
            it.next(prm);
            if (it.tile < num_tiles) epi1(L + 1);
            epi2(L, img, p0, q0);
            ++L;
        }
        if (issuer) ptx::tma_store_wait_read<0>();
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_base, 512);
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

CUtensorMapSwizzle swizzle_for(int row_bytes)
{
    return row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
         : row_bytes == 64  ? CU_TENSOR_MAP_SWIZZLE_64B
         : row_bytes == 32  ? CU_TENSOR_MAP_SWIZZLE_32B
                            : CU_TENSOR_MAP_SWIZZLE_NONE;
}

// Driver quirk handled the way CUTLASS does it (cute/atom/copy_traits_sm90_im2col.hpp, "driver_version <=
// 13010"): for tensors smaller than 128 KiB a descriptor bit must be cleared or the copy misbehaves.
void small_tensor_fixup(CUtensorMap* tm, size_t tensor_bytes, int driver_version)
{
    if (driver_version <= 13010 && tensor_bytes < 131072)
        reinterpret_cast<uint64_t*>(tm)[1] &= ~(1ull << 21);
}

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device attribute: one bit per device ordinal
std::mutex g_attr_mu;
uint64_t g_attr_devices[4] = {0, 0, 0, 0};

uint32_t round_up(uint32_t v, uint32_t a) { return (v + a - 1) / a * a; }

}  // namespace

// rank-4 (C, W, H, N) uint8 tiled map without swizzle, out-of-bounds elements read as zero (used by depthwise.cu)
lbc_status encode_tiled_u8_4d(CUtensorMap* tm, const void* base, const uint64_t dims[4], const uint32_t box[4])
{
    lbc_status st = resolve_driver_entry_points();
    if (st != LBC_OK) return st;
    const cuuint64_t d[4] = {dims[0], dims[1], dims[2], dims[3]};
    const cuuint64_t strides[3] = {dims[0], dims[0] * dims[1], dims[0] * dims[1] * dims[2]};
    const cuuint32_t b[4] = {box[0], box[1], box[2], box[3]};
    const cuuint32_t ones[4] = {1, 1, 1, 1};
    CUresult r = g_encode_tiled(tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, const_cast<void*>(base), d, strides, b, ones,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    LBC_REQUIRE(r == CUDA_SUCCESS, LBC_ERR_CUDA, "cuTensorMapEncodeTiled(rank 4) failed: %d", (int)r);
    return LBC_OK;
}

bool igemm_trace_compiled() { return LBC_TRACE != 0; }

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

static lbc_status make_config_impl(const ConvGeom& g, const DeviceInfo& dev, const lbc_plan_options& o, bool allow_pair,
                                   IgemmConfig* cfg);

lbc_status igemm_make_config(const ConvGeom& g, const DeviceInfo& dev, const lbc_plan_options& opt, IgemmConfig* cfg)
{
    // pair mode doubles the window footprint: where that does not fit (very wide rows), plan without it
    if (make_config_impl(g, dev, opt, true, cfg) == LBC_OK) return LBC_OK;
    return make_config_impl(g, dev, opt, false, cfg);
}

// `o`: explicit planner options (include/lowbit_cnn.h); tri-states are -1 = planner decides / 0 / 1, limits 0 = default
static lbc_status make_config_impl(const ConvGeom& g, const DeviceInfo& dev, const lbc_plan_options& o, bool allow_pair,
                                   IgemmConfig* cfg)
{
    const lbc_conv_desc& d = g.d;
    IgemmConfig c{};
    auto limit = [](int32_t v, int dflt) { return v > 0 ? (int)v : dflt; };
    // ---- N tile: the whole K_out when it fits one 256-wide tile, else an even split into the fewest tiles
    // (multiples of 16; a ragged last tile is handled by TMA zero-fill on loads and clipping on stores)
    int max_bn = std::max(16, std::min(256, limit(o.max_bn, 256)));
    // Small problems (a strong-scaled batch: 64 images per GPU leave ResNet-50's stage 4 with 25 M tiles): when 256-wide
    // tiles would occupy less than half of the SMs, 128-wide ones put twice as many CTAs to work (l4.x.conv2 at N = 64:
    // 16.8 -> 13.6 us, r02); the narrower MMA's lower efficiency does not matter on an under-filled chip.
    if (o.max_bn <= 0 && d.k > 128) {
        const int64_t m_tiles = (g.m_total + kBlockM - 1) / kBlockM;
        const int64_t n_tiles = (d.k + 255) / 256;
        if (2 * m_tiles * n_tiles <= (dev.sm_count > 0 ? dev.sm_count : 148)) max_bn = 128;
    }
    c.tiles_n = (d.k + max_bn - 1) / max_bn;
    c.bn = ((d.k + c.tiles_n - 1) / c.tiles_n + 15) / 16 * 16;
    c.tiles_n = (d.k + c.bn - 1) / c.bn;

    // ---- A mode
    const bool pure_gemm = d.r == 1 && d.s == 1 && d.stride_h == 1 && d.stride_w == 1 && d.pad_h == 0 && d.pad_w == 0;
    c.mode = pure_gemm ? A_TILED : A_IM2COL;
    if (o.force_im2col == 1) c.mode = A_IM2COL;
    const bool c16 = (d.c == 16);
    c.s_pad = d.s;
    c.rows_per_tile = c.cols_per_tile = c.row_tiles = c.col_tiles = c.wt = 0;
    if (!pure_gemm && d.stride_h == 1 && d.stride_w == 1 && o.force_im2col != 1 && o.window != 0) {
        // 16-byte pixels: taps are consumed in pairs and a B block is one filter row of 32/64/128 bytes
        const int s_eff = !c16 ? d.s : (d.s <= 2 ? 2 : d.s <= 4 ? 4 : 8);
        const int ext_w = (s_eff - 1) * d.dil_w, ext_h = (d.r - 1) * d.dil_h;
        int col_tiles = (g.q + kBlockM - 1) / kBlockM;
        int cols = (g.q + col_tiles - 1) / col_tiles;
        int wt = cols + ext_w;
        int rows = 1;
        while (rows < g.p && rows * wt + cols <= kBlockM) ++rows;   // (rows-1)*wt + cols <= 128
        const int row_tiles = (g.p + rows - 1) / rows;
        const double eff = (double)g.p * g.q / ((double)row_tiles * col_tiles * kBlockM);
        const bool s_ok = !c16 || d.s <= 8;
        // Narrow N tiles (one tile of 32 / 64 columns, int8 out) can drain with per-warp stores (warp_store = 1, see the
        // kernel's narrow warp-store path), which needs a warp's 32 accumulator lanes to be one run of pixels of one tile row: pad the
        // window pitch to a power of two >= 32 where that keeps the number of rows per tile (56 + 2 -> 64, 28 + 2 -> 32,
        // 112 + 3 -> 128; the extra pixels per window row are fetched but never read by an MMA row that is stored).
        if (d.out_mode == LBC_OUT_INT8 && (c.bn == 64 || c.bn == 32) && c.tiles_n == 1 && o.warp_store == 1 && o.small_teams != 0 &&
            o.four_acc != 0) {
            int wt2 = 32;
            while (wt2 < wt) wt2 <<= 1;
            int rows2 = 1;
            while (rows2 < g.p && rows2 * wt2 + cols <= kBlockM) ++rows2;
            if (wt2 <= kBlockM && rows2 == rows) wt = wt2;
        }
        if (eff >= 0.55 && wt <= 256 && rows + ext_h <= 256 && s_ok && (rows - 1) * wt + cols <= kBlockM) {
            c.mode = A_WINDOW;
            c.s_pad = s_eff;
            c.rows_per_tile = rows; c.cols_per_tile = cols; c.row_tiles = row_tiles; c.col_tiles = col_tiles; c.wt = wt;
        }
        // Wide layers that will run in CTA pairs (256-wide N tiles, >= 256 channels): a 128-cycle MMA consumes 4 KB of A,
        // which L2 sustains even with every tap fetched separately, so the im2col mode's fully used M tiles (no halo rows,
        // no ragged last row tile: 100% against 77-88% of the MMA rows) win - measured 4% (14x14x256), 10% (28x28x512)
        // and 17% (14x14x512) with pairs, while 128-wide tiles (64-cycle MMAs) stay 25% faster with windows.
        if (c.mode == A_WINDOW && c.bn == 256 && d.c >= 256 && d.c % 128 == 0 && o.cta_pairs != 0 && o.keep_window != 1) {
            c.mode = A_IM2COL;
            c.s_pad = d.s;
            c.rows_per_tile = c.cols_per_tile = c.row_tiles = c.col_tiles = c.wt = 0;
        }
    }

    // ---- K chunking
    if (c.mode == A_WINDOW && c16) {
        c.bkc = 16; c.c_pad = 16;
        c.bkb = c.s_pad * 16;
        c.cblocks = 1;
        c.inner = d.r;
    } else {
        c.bkc = (d.c % 128 == 0) ? 128 : (d.c % 64 == 0) ? 64 : 32;
        c.c_pad = (d.c + c.bkc - 1) / c.bkc * c.bkc;
        c.bkb = c.bkc;
        c.cblocks = c.c_pad / c.bkc;
        c.inner = (c.mode == A_TILED) ? 1 : d.r * d.s;
    }
    c.k_blocks = c.cblocks * c.inner;
    c.packed_row_bytes = (size_t)c.k_blocks * c.bkb;

    // ---- pair mode (see IgemmParams::pair): window layers whose filter matrix is too large to stay resident.
    // N tiles of <= 128 columns leave TMEM room for the two accumulators of a pair, double-buffered.
    c.pair = 0;
    if (c.mode == A_WINDOW && !c16 && (size_t)d.k * c.packed_row_bytes > 80u * 1024u && (d.k <= 128 || d.k % 128 == 0) &&
        d.n * c.row_tiles * c.col_tiles >= 2 && allow_pair && o.paired_tiles == 1) {   // opt-in: measured gains are marginal (r01)
        c.pair = 1;
        c.bn = d.k <= 128 ? d.k : 128;
        c.tiles_n = d.k / c.bn;
    }

    // ---- M tiles
    if (c.mode == A_WINDOW) c.tiles_m = d.n * c.row_tiles * c.col_tiles;
    else c.tiles_m = (int32_t)((g.m_total + kBlockM - 1) / kBlockM);

    bool pair_res = false;
    // ---- CTA-pair mode (see IgemmParams::cta2): layers whose filter matrix streams through the ring (too large to stay
    // resident, or several N tiles).  Their MMAs run at 128 operand bytes per cycle out of shared memory while TMA fills
    // the same banks, and ~80% tensor-pipe utilisation was the measured ceiling (profiles/r01_*); sharing B between
    // the two SMs of a TPC removes a third of that traffic.
    {
        const size_t full_b = (size_t)c.k_blocks * c.bn * c.bkb;
        const bool streams = !((size_t)c.tiles_n * full_b <= 80u * 1024u) || o.resident_filter == 0;
        // (a small matrix pairs only on request - cta_pairs = 1, window A - and then stays resident as two halves)
        const bool small_on_request = !streams && o.cta_pairs == 1 && c.mode == A_WINDOW && c.tiles_n == 1 && o.resident_filter != 0;
        const bool possible = !c.pair && !c16 && (streams || small_on_request) && c.bn % 32 == 0 && c.tiles_m >= 2;
        // Pairing couples the two CTAs' pipelines (one MMA stream waits for both producers and both epilogues), which
        // costs where the epilogue is the bound: short K loops (1x1 channel expansions) measured 10-15% slower in pairs,
        // long ones (3x3, wide 1x1 reductions) 5-25% faster.  Cross-over: MMA time per tile (128 cycles per N=256
        // instruction) against ~16 cycles per output column of epilogue.
        const int mma_cycles = c.k_blocks * (c.bkb / 32) * (c.bn / 2);
        bool want = mma_cycles >= 16 * c.bn;
        // A pair can keep the matrix resident as two halves (see res_b_ok below), which removes the filter stream from L2.
        // Ring modes (the kernels exist: 128-byte channel chunks) only with resident_filter = 3: measured slower there -
        // 512 -> 256 65.5 -> 69.7 us, 1024 -> 256 38.9 -> 43.0 us, 3x3 stride 2 71.7 -> 73.8 us (r02, N=512, cold) - those
        // layers wait for A from HBM, not for L2; window layers gain 4% (128 -> 128 @28x28) to 24% (@112x112, VGG conv2_2).
        pair_res = possible && c.tiles_n == 1 && o.resident_filter != 0 && o.resident_filter != 2 &&
                   (c.mode == A_WINDOW || (c.bkb == 128 && o.resident_filter == 3)) &&
                   full_b / 2 <= (size_t)limit(o.resident_kb, c.mode == A_WINDOW ? 80 : 128) * 1024u;
        if (pair_res && o.resident_filter == 3 && mma_cycles >= 8 * c.bn) want = true;
        if (o.cta_pairs >= 0) want = o.cta_pairs != 0;       // test / tuning override
        c.cta2 = (possible && want) ? 1 : 0;
    }

    // ---- output staging (int8 mode): panels of <= 128 bytes per row
    if (c.bn % 128 == 0) c.panel_bytes = 128;
    else if (c.bn == 64 || c.bn == 32) c.panel_bytes = c.bn;
    else c.panel_bytes = c.bn;                      // unswizzled single panel (rare channel counts)
    // ring modes with int8 output: per-warp staging and stores (see the kernel's epilogue) wherever the N tile splits into
    // whole panels per warp - 256 -> 2 x 128 B, 128 -> 2 x 64 B for the two warp sets of a team, <= 64 -> one panel
    // Measured (r01): +5-10% on 256-wide tiles with a short K loop (the 1x1 channel expansions, whose epilogue is the
    // bound); a loss where the extra staging bytes cost ring depth (long K loops) and on narrow tiles (2 KB stores).
    // Narrow tiles (r02): one N tile of 32 / 64 columns can drain through the kernel's narrow warp-store path in every A
    // mode (window tiles need the padded pitch above).  It removes the team paths' per-tile barriers and store waits
    // (2450 -> ~1700 cycles per team and tile in the traces) - and changes nothing: with 64-column MMAs the tensor pipe
    // re-reads the 4 KB A operand for every 2 KB of B, 48 cycles of shared-memory bandwidth per 32-cycle MMA, and that
    // (plus the window and staging traffic through the same banks) is what paces the stems and the 64-channel layers.
    // Measured (r02, N=512): conv1 222 -> 228 us, 64->64 3x3 62.5 -> 66.6 us, 1x1 64->64 / 256->64 unchanged.  Opt-in.
    c.warp_store = 0;
    const bool narrow_ws = (c.bn == 64 || c.bn == 32) && c.tiles_n == 1 && !c.pair && !c.cta2 && o.small_teams != 0 && o.four_acc != 0 &&
                           (c.mode != A_WINDOW || (c.wt >= 32 && (c.wt & (c.wt - 1)) == 0));
    {
        bool want = c.bn == 256 && !c.cta2 && c.k_blocks * (c.bkb / 32) <= 8;
        if (o.warp_store >= 0) want = o.warp_store != 0;     // test / tuning override (the only way onto the narrow path)
        if ((c.mode != A_WINDOW || narrow_ws) && d.out_mode == LBC_OUT_INT8 && want &&
            (c.bn == 256 || c.bn == 128 || c.bn == 64 || c.bn == 32)) {
            c.warp_store = 1;
            if (c.bn == 128) c.panel_bytes = 64;
        }
    }
    // ---- column-split epilogue (see IgemmParams::epi_split): N tiles with only two TMEM accumulator stages
    c.epi_split = 0;
    {
        const bool two_acc = !(4 * c.bn <= 512 && (c.pair || o.four_acc != 0));
        bool can = two_acc && !c.pair;
        if (d.out_mode == LBC_OUT_INT8) {
            if (c.warp_store) can = can && c.bn == 256;                     // four 64-byte panels: one per (team, warp set)
            else can = can && (c.bn / c.panel_bytes) % 2 == 0;              // panels dealt alternately to the two teams
        }
        // Measured (r02, ResNet-50 N=512): a loss where the epilogue is the bound - the 1x1 channel expansions went from
        // 105 to 150 us (64->256) and 62 to 82 us (128->512): with all 16 warps on one tile the per-tile fixed costs (waits,
        // fences, store issue: ~1000 cycles) are no longer hidden behind the other team's drain, and each warp converts
        // half as many columns per tile - and neutral (+-2%) on the MMA-bound layers.  Kept as an option, off by default.
        bool want = false;
        if (o.epi_split >= 0) want = o.epi_split != 0;
        if (can && want) {
            c.epi_split = 1;
            if (c.warp_store && d.out_mode == LBC_OUT_INT8) c.panel_bytes = 64;
        }
    }
    c.n_panels = c.bn / c.panel_bytes;
    c.panel_swz_bits = c.panel_bytes == 128 ? 3 : c.panel_bytes == 64 ? 2 : c.panel_bytes == 32 ? 1 : 0;

    // ---- accumulator stages and epilogue teams
    c.n_acc = (4 * c.bn <= 512 && (c.pair || o.four_acc != 0)) ? 4 : 2;
    c.team_warps = (c.n_acc == 4 && c.bn <= 64 && o.small_teams != 0) ? 4 : 8;
    // narrow tiles: 8 accumulator stages and two tiles per epilogue iteration (see the kernel's tpi == 2 path)
    c.tpi = 1;
    if (c.team_warps == 4 && c.tiles_n == 1 && !c.pair && !c.cta2 && !c.warp_store && c.bn == c.panel_bytes && o.tiles_per_iter2 != 0) {
        c.tpi = 2;
        c.n_acc = 8;
    }
    // narrow warp-store path: 8 stages, one tile per team step
    if (c.warp_store && narrow_ws && c.team_warps == 4) c.n_acc = 8;
    const int n_teams = kEpiWarps / c.team_warps;

    // ---- smem carve-up: [A ring | window ring][B ring][output staging][control]
    // Prefetch depth is what hides the ~2 us loaded HBM latency, so the A side (ring stages, or windows) gets every
    // byte left over after two or three B stages (B comes from L2 and needs little run-ahead) and the staging panels.
    const uint32_t ctl_bytes = round_up((uint32_t)sizeof(Ctl), 256);
    c.a_block_bytes = (c.mode == A_WINDOW) ? 0 : (uint32_t)(kBlockM * c.bkc);
    c.b_block_bytes = (uint32_t)((c.cta2 ? c.bn / 2 : c.bn) * c.bkb);     // a pair's CTAs hold half of the B rows each
    // tuning limits (lbc_plan_options; the defaults are what ships)
    const uint32_t tps_cap = (uint32_t)limit(o.tps_kb, 48) * 1024u;
    const int max_win = std::min(kMaxWinStages, limit(o.max_win_stages, kMaxWinStages));
    const int max_stages = std::min(kMaxStages, limit(o.max_stages, kMaxStages));
    const int max_bufs = std::max(1, std::min(3, limit(o.stage_bufs, 3)));
    // blocks per ring stage: group small B blocks (window mode) so one mbarrier round trip feeds several MMAs
    bool fits = false;
    uint32_t stage_bytes = 0;
    // a pair issues twice the MMAs per block: small stages, deep ring; otherwise try the large grouping first and fall
    // back to single blocks when shared memory is tight (wide windows)
    const uint32_t caps[2] = {c.pair ? 16u * 1024u : tps_cap, 16u * 1024u};
    for (int ci = 0; ci < 2 && !fits; ++ci) {
    c.tps = 1;
    if (c.mode == A_WINDOW) {
        for (int t = c.inner; t >= 1; --t)
            if (c.inner % t == 0 && (uint32_t)t * c.b_block_bytes <= caps[ci]) { c.tps = t; break; }
    }
    c.a_stage_bytes = c.tps * c.a_block_bytes;
    c.b_stage_bytes = c.tps * c.b_block_bytes;
    c.win_stage_bytes = c.win_tx_bytes = 0;
    if (c.mode == A_WINDOW) {
        const int s_eff = c.s_pad;
        const int ext_w = (s_eff - 1) * d.dil_w, ext_h = (d.r - 1) * d.dil_h;
        const uint32_t box_bytes = (uint32_t)(c.wt * (c.rows_per_tile + ext_h) * c.bkc);
        const uint32_t reach = (uint32_t)((kBlockM + ext_h * c.wt + ext_w + 1) * c.bkc);   // furthest row an MMA reads
        c.win_tx_bytes = box_bytes;
        c.win_sub_bytes = round_up(std::max(box_bytes, reach), 1024);
        c.win_stage_bytes = c.pair ? 2 * c.win_sub_bytes : c.win_sub_bytes;
    }
    // Resident filter matrix: one N tile and the whole packed matrix small enough to leave room for a deep A side.
    // It removes the per-block ring handshake (~400 cycles each, measured) and the L2 re-fetch of B for every tile.
    c.b_total_bytes = (uint32_t)c.tiles_n * c.k_blocks * c.b_block_bytes;      // all N tiles
    bool res_b_ok = !c.pair && !c.cta2 && c.b_total_bytes <= (uint32_t)limit(o.resident_kb, 80) * 1024u && o.resident_filter != 0;
    // CTA pairs with window A: each CTA holds half of the filter rows, so a matrix of up to twice the limit stays resident
    // (128 -> 128 3x3: 2 x 72 KB).  Streaming it cost 74 KB of L2 reads per 2304-cycle tile and CTA - with the windows
    // 42 B/clk/SM, which IS the chip's L2 bandwidth (~6300 B/clk over 148 SMs): the layer ran at 0.62 of the tensor peak.
    if (c.cta2 && pair_res) res_b_ok = true;
    // N-stationary: the matrix as a whole is too large, but one N tile fits and the persistent grid can be a multiple of
    // tiles_n, so every CTA keeps "its" N tile for the whole launch (see IgemmParams::res_one)
    c.res_one = 0;
    {
        const uint32_t one_tile = (uint32_t)c.k_blocks * c.b_block_bytes;
        int grid0 = std::min(dev.sm_count > 0 ? dev.sm_count : 148, c.tiles_m * c.tiles_n);
        if (o.max_grid > 0) grid0 = std::max(1, std::min(grid0, (int)o.max_grid));
        const int grid_ns = grid0 - grid0 % c.tiles_n;
        // Measured (r02): -5% on 256->1024 @14x14 and 256->512 stride 2 (64 KB tiles); 128 KB tiles (512->1024, 512->2048)
        // leave room for only three A stages and came out 0 to 7% slower, so the planner's own limit is 64 KB
        // (n_stationary = 1 raises it to 128 KB for the tests).
        const uint32_t one_limit = (o.n_stationary == 1 ? 128u : 64u) * 1024u;
        if (!res_b_ok && !c.pair && !c.cta2 && c.tiles_n > 1 && one_tile <= one_limit && grid_ns >= c.tiles_n &&
            o.resident_filter != 0 && o.n_stationary != 0) {
            c.res_one = 1;
            c.b_total_bytes = one_tile;
            res_b_ok = true;
        }
    }
    for (int pass = 0; pass < 2 && !fits; ++pass)
    for (int bufs = c.warp_store ? std::min(2, max_bufs) : max_bufs; bufs >= 1 && !fits; --bufs)
    for (int fold_try = 1; fold_try >= 0 && !fits; --fold_try) {     // with the bias-fold blocks if they fit, else without   // three staging panels per team when they fit, else two, else one
        c.res_b = (pass == 0 && res_b_ok) ? 1 : 0;
        if (pass == 0 && !res_b_ok) break;
        c.stage_bufs = bufs;
        stage_bytes = d.out_mode == LBC_OUT_INT8 ? round_up((uint32_t)(n_teams * bufs * kBlockM * c.panel_bytes), 1024) : 0;
        if (c.warp_store) stage_bytes = (uint32_t)(kEpiWarps * 32 * c.panel_bytes * bufs);   // 32-row panels, `bufs` per epilogue warp
        // bias folded into the MMA (resident filter matrix only): a 4 KB constant A block + bn x 32 B of bias digits
        // Enabled for N tiles >= 128 columns: -11% time on the 64->256 expansions, -2% on 128-wide tiles; narrow tiles
        // gain nothing (their per-tile bookkeeping dominates) and pay for the extra MMA per tile.  LBC_FOLD=0/1 overrides.
        {
            bool want = c.bn >= 128;
            if (o.fold_bias >= 0) want = o.fold_bias != 0;
            if (c.cta2) want = false;                           // (no bias-fold variant of the CTA-pair kernels)
            c.fold = (c.res_b && want && fold_try) ? 1 : 0;
            if (!fold_try && !(c.res_b && want)) continue;      // nothing to drop: this variant was already tried
        }
        const uint32_t fold_bytes = c.fold ? round_up(4096u + (uint32_t)(c.res_one ? c.bn : c.bn * c.tiles_n) * 32u, 1024) : 0u;
        if (stage_bytes + ctl_bytes + fold_bytes >= 227u * 1024u) continue;
        const uint32_t budget = 227 * 1024 - stage_bytes - ctl_bytes - fold_bytes;
        uint32_t win_total = 0;
        int stages;
        uint32_t b_region;
        if (c.res_b) {
            if (c.b_total_bytes + 2 * std::max(c.win_stage_bytes, c.a_stage_bytes) > budget) continue;
            b_region = round_up(c.b_total_bytes, 1024);
            // a third staging panel only pays where the epilogue is the bottleneck and must not cost A-side depth
            if (c.mode == A_WINDOW) {
                stages = 0;   // no ring
                c.win_stages = (int)std::min<uint32_t>(max_win, (budget - b_region) / c.win_stage_bytes);
                // Window depth is what hides the HBM latency of the next tiles' halos (a tile lasts ~1300 cycles, a loaded
                // window takes 2-4 k cycles to arrive): 64->64 3x3 @56x56 ran 16% faster with 6 windows and two staging panels
                // per team than with 4 and three (r02), so the third / second panel is only taken once 8 / 6 windows fit.
                if (c.win_stages < (bufs == 3 ? 8 : bufs == 2 ? 6 : 2)) continue;
                win_total = c.win_stages * c.win_stage_bytes;
            } else {
                c.win_stages = 0;
                stages = (int)std::min<uint32_t>(max_stages, (budget - b_region) / c.a_stage_bytes);
                if (stages < (bufs == 3 ? 5 : 3)) continue;
            }
        } else if (bufs == 3 || (bufs == 2 && c.bn > 128 && c.k_blocks > 8 && !c.pair)) {
            continue;   // streaming B with a long K loop: operand depth matters more than extra staging panels
        } else if (c.mode == A_WINDOW) {
            if (2 * c.win_stage_bytes + 2 * c.b_stage_bytes > budget) continue;
            stages = ((budget - 2 * c.win_stage_bytes) / c.b_stage_bytes >= 3 && max_stages >= 3) ? 3 : 2;
            if (c.pair) stages = (int)std::min<uint32_t>(std::min(max_stages, 6), (budget - 2 * c.win_stage_bytes) / c.b_stage_bytes);
            c.win_stages = (int)std::min<uint32_t>(max_win, (budget - stages * c.b_stage_bytes) / c.win_stage_bytes);
            win_total = c.win_stages * c.win_stage_bytes;
            b_region = (uint32_t)stages * c.b_stage_bytes;
        } else {
            c.win_stages = 0;
            stages = (int)std::min<uint32_t>(max_stages, budget / (c.a_stage_bytes + c.b_stage_bytes));
            if (stages < 2) continue;
            b_region = (uint32_t)stages * c.b_stage_bytes;
        }
        c.stages = stages;
        c.off_b = (c.mode == A_WINDOW) ? win_total : (uint32_t)stages * c.a_stage_bytes;
        c.off_stage = c.off_b + b_region;
        c.off_fold = c.off_stage + stage_bytes;
        c.off_ctl = c.off_fold + fold_bytes;
        c.smem_bytes = c.off_ctl + ctl_bytes;
        fits = c.smem_bytes <= 227 * 1024;
    }
    }   // tps candidates
    LBC_REQUIRE(fits, LBC_ERR_UNSUPPORTED, "igemm: operand rings do not fit in shared memory");
    // Two MMA-issuing warps (alternate tiles, half of the A-side ring each) where a single warp's issue rate is the
    // bound: narrow N tiles with a resident filter matrix.  Wide tiles (128 cycles per MMA) gain nothing.
    c.n_mma = 1;
    // issue-bound test: ~75 cycles per MMA for one warp (traces) against the tile's share of HBM time (22.5 B/clk/SM)
    const double issue_cycles = 450.0 + 75.0 * c.k_blocks * (c.bkb / 32);   // + per-tile fixed cost of the issuing warp
    const double a_bytes = c.mode == A_WINDOW ? (double)c.win_tx_bytes * c.cblocks : (double)kBlockM * c.c_pad * (c.mode == A_TILED ? 1 : d.r * d.s);
    const double hbm_cycles = (a_bytes + (double)kBlockM * c.bn) / 22.5;
    if (c.res_b && !c.cta2 && c.bn <= 128 && issue_cycles > hbm_cycles && o.two_mma_warps != 0) {
        if (c.mode == A_WINDOW && c.win_stages >= 4) { c.n_mma = 2; c.win_stages &= ~1; }
        else if (c.mode != A_WINDOW && c.stages >= 4) { c.n_mma = 2; c.stages &= ~1; }
    }

    // ---- MMA issue table (A-descriptor offsets in 16-byte units, one channel chunk)
    {
        const int ks = c.bkb / 32;
        if (c.mode != A_WINDOW) {
            c.n_tab = ks;
            for (int k = 0; k < ks; ++k) c.a_tab[k] = (uint16_t)(2 * k);
        } else {
            c.n_tab = c.inner * ks;
            LBC_REQUIRE(c.n_tab <= (int)(sizeof(c.a_tab) / sizeof(c.a_tab[0])), LBC_ERR_UNSUPPORTED,
                        "igemm: %d MMAs per channel chunk exceed the issue table", c.n_tab);
            const uint32_t s_step16 = (uint32_t)(d.dil_w * c.bkc) >> 4;            // next tap in the filter row
            const uint32_t r_step16 = (uint32_t)(d.dil_h * c.wt * c.bkc) >> 4;     // next filter row
            for (int t = 0; t < c.inner; ++t)
                for (int k = 0; k < ks; ++k) {
                    uint32_t off;
                    if (c.bkc == 16) off = (uint32_t)t * r_step16 + (uint32_t)k * 2u * s_step16;   // block = filter row, K-step = two taps
                    else off = (uint32_t)(t / d.s) * r_step16 + (uint32_t)(t % d.s) * s_step16 + 2u * (uint32_t)k;
                    const uint32_t boff = (uint32_t)t * (c.b_block_bytes >> 4) + 2u * (uint32_t)k;
                    LBC_REQUIRE(off < 65536u && boff < 65536u, LBC_ERR_UNSUPPORTED, "igemm: window too large for the issue table");
                    c.a_tab[t * ks + k] = (uint16_t)off;
                    c.b_tab[t * ks + k] = (uint16_t)boff;
                }
        }
    }
    uint32_t cols = 32;
    while (cols < (uint32_t)(c.n_acc * c.bn)) cols <<= 1;
    c.tmem_cols = cols;
    if (!c.res_b) c.res_one = 0;       // the resident variant did not fit: plain streaming
    c.grid = std::min(dev.sm_count > 0 ? dev.sm_count : 148, c.tiles_m * c.tiles_n);
    c.it_imgs = d.n;
    if (c.cta2) {
        // the two CTAs of a cluster take the consecutive tiles 2q, 2q+1 of an N tile: pad the M-tile space to an even count
        c.it_imgs = c.mode == A_WINDOW ? d.n : c.tiles_m;
        const int32_t per_img = c.mode == A_WINDOW ? c.row_tiles * c.col_tiles : 1;
        if ((per_img * c.it_imgs) & 1) ++c.it_imgs;
        const int32_t sms = (dev.sm_count > 0 ? dev.sm_count : 148) & ~1;
        c.grid = std::min(sms, per_img * c.it_imgs * c.tiles_n);
    }
    if (c.pair) {
        // pad the image radix so that the tiles of one N tile come in whole pairs; CTAs walk pairs
        if ((c.row_tiles * c.col_tiles * c.it_imgs) & 1) ++c.it_imgs;
        const int32_t pairs = c.row_tiles * c.col_tiles * c.it_imgs / 2 * c.tiles_n;
        c.grid = std::min(dev.sm_count > 0 ? dev.sm_count : 148, pairs);
    }
    if (o.max_grid > 0) c.grid = std::max(1, std::min(c.grid, (int)o.max_grid));   // tests: many tiles per CTA
    if (c.cta2) c.grid = std::max(2, c.grid & ~1);
    if (c.res_one) c.grid -= c.grid % c.tiles_n;      // >= tiles_n by the planner's check above
    c.reverse = o.reverse == 1 ? 1 : 0;
    c.pdl = o.pdl != 0 ? 1 : 0;
    // ---- split last round (see IgemmParams::tail_first)
    c.tail_first = -1; c.tail_count = 0; c.tail_m0 = 0;
    if (c.cta2 && c.mode != A_WINDOW && c.tiles_n == 1 && c.bn == 256 && !c.res_b && !c.epi_split && !c.warp_store &&
        d.out_mode == LBC_OUT_INT8 && o.tail_split != 0) {
        const int pairs = c.grid / 2, steps = c.it_imgs / 2;           // (it_imgs: the M tiles, padded to an even count)
        const int full = steps / pairs, rest = steps - full * pairs;
        if (full >= 1 && rest > 0 && 2 * rest <= pairs) {
            c.tail_first = full * c.grid;
            c.tail_count = 4 * rest;
            c.tail_m0 = 2 * full * pairs;
        }
    }
    *cfg = c;
    return LBC_OK;
}

lbc_status igemm_encode(const ConvGeom& g, const IgemmConfig& cfg, const DeviceInfo& dev, const int8_t* x,
                        const int8_t* w_packed, void* y, IgemmLaunch* out)
{
    lbc_status st = resolve_driver_entry_points();
    if (st != LBC_OK) return st;
    const lbc_conv_desc& d = g.d;
    LBC_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(w_packed) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(y) & 15) == 0,
                LBC_ERR_INVALID_ARG, "igemm: x, packed weights and y must be 16-byte aligned");
    out->cfg = cfg;
    const cuuint32_t ones[5] = {1, 1, 1, 1, 1};
    const CUtensorMapL2promotion promo = CU_TENSOR_MAP_L2_PROMOTION_L2_256B;

    // ---- B: [K_out rows][packed_row_bytes], box {bkb, bn}
    {
        const cuuint64_t dims[2] = {(cuuint64_t)cfg.packed_row_bytes, (cuuint64_t)d.k};
        const cuuint64_t strides[1] = {(cuuint64_t)cfg.packed_row_bytes};
        const cuuint32_t box[2] = {(cuuint32_t)cfg.bkb, (cuuint32_t)(cfg.cta2 ? cfg.bn / 2 : cfg.bn)};
        CUresult r = g_encode_tiled(&out->tm_b, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, (void*)w_packed, dims, strides, box,
                                    ones, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(cfg.bkb), promo,
                                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        LBC_REQUIRE(r == CUDA_SUCCESS, LBC_ERR_CUDA, "cuTensorMapEncodeTiled(B) failed: %d", (int)r);
    }
    // ---- A
    if (cfg.mode == A_TILED) {
        const cuuint64_t dims[2] = {(cuuint64_t)d.c, (cuuint64_t)g.m_total};
        const cuuint64_t strides[1] = {(cuuint64_t)d.c};
        const cuuint32_t box[2] = {(cuuint32_t)cfg.bkc, (cuuint32_t)kBlockM};
        CUresult r = g_encode_tiled(&out->tm_a, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, (void*)x, dims, strides, box, ones,
                                    CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(cfg.bkc), promo,
                                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        LBC_REQUIRE(r == CUDA_SUCCESS, LBC_ERR_CUDA, "cuTensorMapEncodeTiled(A) failed: %d", (int)r);
    } else if (cfg.mode == A_IM2COL) {
        // rank-4 (C, W, H, N); corners in (W, H) order.  lower = -pad ; upper = pad - (filter-1)*dilation.
        const cuuint64_t dims[4] = {(cuuint64_t)d.c, (cuuint64_t)d.w, (cuuint64_t)d.h, (cuuint64_t)d.n};
        const cuuint64_t strides[3] = {(cuuint64_t)d.c, (cuuint64_t)d.c * d.w, (cuuint64_t)d.c * d.w * d.h};
        const int lower[2] = {-d.pad_w, -d.pad_h};
        const int upper[2] = {d.pad_w - (d.s - 1) * d.dil_w, d.pad_h - (d.r - 1) * d.dil_h};
        const cuuint32_t trav[4] = {1, (cuuint32_t)d.stride_w, (cuuint32_t)d.stride_h, 1};
        CUresult r = g_encode_im2col(&out->tm_a, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, (void*)x, dims, strides, lower,
                                     upper, (cuuint32_t)cfg.bkc, (cuuint32_t)kBlockM, trav,
                                     CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(cfg.bkc), promo,
                                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        LBC_REQUIRE(r == CUDA_SUCCESS, LBC_ERR_CUDA, "cuTensorMapEncodeIm2col(A) failed: %d", (int)r);
        small_tensor_fixup(&out->tm_a, (size_t)d.n * d.h * d.w * d.c, dev.driver_version);
    } else {
        // WINDOW: rank-4 tiled (C, W, H, N), box {bkc, wt, rows + (R-1)*dil_h, 1}; OOB zero-fill is the padding
        const cuuint64_t dims[4] = {(cuuint64_t)d.c, (cuuint64_t)d.w, (cuuint64_t)d.h, (cuuint64_t)d.n};
        const cuuint64_t strides[3] = {(cuuint64_t)d.c, (cuuint64_t)d.c * d.w, (cuuint64_t)d.c * d.w * d.h};
        const cuuint32_t box[4] = {(cuuint32_t)cfg.bkc, (cuuint32_t)cfg.wt,
                                   (cuuint32_t)(cfg.rows_per_tile + (d.r - 1) * d.dil_h), 1};
        CUresult r = g_encode_tiled(&out->tm_a, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, (void*)x, dims, strides, box, ones,
                                    CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(cfg.bkc), promo,
                                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        LBC_REQUIRE(r == CUDA_SUCCESS, LBC_ERR_CUDA, "cuTensorMapEncodeTiled(A window) failed: %d", (int)r);
    }
    // ---- output (int8 mode): staged tile -> TMA store
    if (d.out_mode == LBC_OUT_INT8) {
        const CUtensorMapSwizzle oswz = cfg.panel_swz_bits ? swizzle_for(cfg.panel_bytes) : CU_TENSOR_MAP_SWIZZLE_NONE;
        CUresult r;
        bool have_out2 = false;
        if (cfg.mode == A_WINDOW) {
            const cuuint64_t dims[4] = {(cuuint64_t)d.k, (cuuint64_t)g.q, (cuuint64_t)g.p, (cuuint64_t)d.n};
            const cuuint64_t strides[3] = {(cuuint64_t)d.k, (cuuint64_t)d.k * g.q, (cuuint64_t)d.k * g.q * g.p};
            // per-warp stores: runs of 32 pixels of one tile row (and the ragged last run of a row)
            const cuuint32_t box[4] = {(cuuint32_t)cfg.panel_bytes, (cuuint32_t)(cfg.warp_store ? 32 : cfg.cols_per_tile),
                                       (cuuint32_t)(cfg.warp_store ? 1 : cfg.rows_per_tile), 1};
            r = g_encode_tiled(&out->tm_out, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, y, dims, strides, box, ones,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, oswz, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r == CUDA_SUCCESS && cfg.warp_store && cfg.cols_per_tile % 32 != 0) {
                const cuuint32_t box2[4] = {(cuuint32_t)cfg.panel_bytes, (cuuint32_t)(cfg.cols_per_tile % 32), 1, 1};
                r = g_encode_tiled(&out->tm_out2, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, y, dims, strides, box2, ones,
                                   CU_TENSOR_MAP_INTERLEAVE_NONE, oswz, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
                have_out2 = true;
            }
        } else {
            const cuuint64_t dims[2] = {(cuuint64_t)d.k, (cuuint64_t)g.m_total};
            const cuuint64_t strides[1] = {(cuuint64_t)d.k};
            const cuuint32_t box[2] = {(cuuint32_t)cfg.panel_bytes, (cuuint32_t)(cfg.warp_store ? 32 : kBlockM)};
            r = g_encode_tiled(&out->tm_out, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, y, dims, strides, box, ones,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, oswz, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        }
        LBC_REQUIRE(r == CUDA_SUCCESS, LBC_ERR_CUDA, "cuTensorMapEncodeTiled(out) failed: %d", (int)r);
        if (!have_out2) out->tm_out2 = out->tm_out;
    } else {
        out->tm_out = out->tm_b;   // unused in int32 mode; keep the parameter a valid descriptor
        out->tm_out2 = out->tm_b;
    }
    return LBC_OK;
}

// kernel parameters of a launch: geometry, tiling, smem carve-up, iteration digits
static void fill_params(const ConvGeom& g, const IgemmLaunch& l, const EpilogueParams& ep, const IgemmRuntime& rt, IgemmParams& prm)
{
    const IgemmConfig& c = l.cfg;
    const lbc_conv_desc& d = g.d;
    prm.m_total = g.m_total;
    prm.k_out = d.k; prm.n_img = d.n; prm.p = g.p; prm.q = g.q;
    prm.mode = c.mode; prm.bn = c.bn; prm.bkc = c.bkc; prm.bkb = c.bkb;
    prm.tiles_m = c.tiles_m; prm.tiles_n = c.tiles_n; prm.cblocks = c.cblocks; prm.inner = c.inner;
    prm.s_taps = d.s;
    prm.stride_h = d.stride_h; prm.stride_w = d.stride_w; prm.pad_h = d.pad_h; prm.pad_w = d.pad_w;
    prm.dil_h = d.dil_h; prm.dil_w = d.dil_w;
    prm.stages = c.stages; prm.a_stage_bytes = c.a_stage_bytes; prm.b_stage_bytes = c.b_stage_bytes;
    prm.tps = c.tps; prm.a_block_bytes = c.a_block_bytes; prm.b_block_bytes = c.b_block_bytes;
    prm.win_stages = c.win_stages; prm.win_stage_bytes = c.win_stage_bytes; prm.win_tx_bytes = c.win_tx_bytes;
    prm.wt = c.wt; prm.rows_per_tile = c.rows_per_tile; prm.cols_per_tile = c.cols_per_tile;
    prm.row_tiles = c.row_tiles; prm.col_tiles = c.col_tiles;
    prm.relu = ep.relu; prm.out_mode = ep.out_mode;
    prm.tmem_cols = c.tmem_cols; prm.n_acc = c.n_acc; prm.tpi = c.tpi;
    prm.panel_bytes = c.panel_bytes; prm.panel_swz_bits = c.panel_swz_bits; prm.n_panels = c.n_panels;
    prm.stage_bufs = c.stage_bufs; prm.k_mod = c.k_mod; prm.team_warps = c.team_warps; prm.warp_store = c.warp_store;
    prm.epi_split = c.epi_split;
    prm.n_tab = c.n_tab;
    for (int i = 0; i < c.n_tab; ++i) { prm.a_tab[i] = c.a_tab[i]; prm.b_tab[i] = c.b_tab[i]; }
    prm.res_b = c.res_b; prm.res_one = c.res_one; prm.b_total_bytes = c.b_total_bytes; prm.n_mma = c.n_mma;
    // digits of the CTA stride in the (img | rt | ct | n_blk) tile numbering
    prm.it_cols = c.mode == A_WINDOW ? c.col_tiles : 1;
    prm.it_rows = c.mode == A_WINDOW ? c.row_tiles : 1;
    prm.pair = c.pair; prm.it_imgs = c.it_imgs; prm.win_sub_bytes = c.win_sub_bytes;
    prm.cta2 = c.cta2; prm.n_major = (c.pair || c.cta2) ? 1 : 0;
    prm.tile_stride = c.pair ? 2 * c.grid : c.grid;
    if (prm.n_major) {   // (n_blk | img | rt | ct), CTA stride = 2 * grid tiles (pair) or grid tiles (CTA pairs)
        int32_t v = prm.tile_stride;
        prm.step_ct = v % prm.it_cols; v /= prm.it_cols;
        prm.step_rt = v % prm.it_rows; v /= prm.it_rows;
        prm.step_img = v % prm.it_imgs; v /= prm.it_imgs;
        prm.step_nb = v;
    } else {
        int32_t v = c.grid;
        prm.step_nb = v % c.tiles_n; v /= c.tiles_n;
        prm.step_ct = v % prm.it_cols; v /= prm.it_cols;
        prm.step_rt = v % prm.it_rows; v /= prm.it_rows;
        prm.step_img = v;
    }
    {   // an epilogue team's stride through the same numbering
        const int32_t n_teams = kEpiWarps / c.team_warps;
        int32_t v = prm.team_stride = n_teams * c.tpi * prm.tile_stride;
        if (prm.n_major) {
            prm.tstep_ct = v % prm.it_cols; v /= prm.it_cols;
            prm.tstep_rt = v % prm.it_rows; v /= prm.it_rows;
            prm.tstep_img = v % prm.it_imgs; v /= prm.it_imgs;
            prm.tstep_nb = v;
        } else {
            prm.tstep_nb = v % c.tiles_n; v /= c.tiles_n;
            prm.tstep_ct = v % prm.it_cols; v /= prm.it_cols;
            prm.tstep_rt = v % prm.it_rows; v /= prm.it_rows;
            prm.tstep_img = v;
        }
    }
    prm.off_b = c.off_b; prm.off_stage = c.off_stage; prm.off_ctl = c.off_ctl;
    prm.fold = c.fold; prm.off_fold = c.off_fold;
    // reversed traversal (IgemmLaunch::reverse, set by the network runner; lbc_plan_options::reverse for single layers)
    prm.rev_m = (l.reverse || c.reverse) ? (c.mode == A_WINDOW ? d.n : c.tiles_m) : 0;
    prm.early_b = (l.early_b && c.pdl) ? l.early_b : 0;          // 2: resident matrices only (lbc_plan_options::early_weights = 2)
    prm.tail_first = c.tail_first; prm.tail_count = c.tail_count; prm.tail_m0 = c.tail_m0;
    prm.trace = rt.trace; prm.trace_tiles = rt.trace_tiles;
    prm.flag = rt.flag;
    // TILED/IM2COL consume cblocks*inner ring blocks in [tap][chunk] order: present them to the MMA loop as one
    // "channel chunk" of cblocks*inner blocks.
    if (c.mode == A_WINDOW) { prm.mma_outer = c.cblocks; prm.mma_inner = c.inner; }
    else { prm.mma_outer = 1; prm.mma_inner = c.cblocks * c.inner; }
}

lbc_status igemm_launch(const ConvGeom& g, const IgemmLaunch& l, const EpilogueParams& ep, void* y,
                        const IgemmRuntime& rt, cudaStream_t stream)
{
    const IgemmConfig& c = l.cfg;
    IgemmParams prm{};
    fill_params(g, l, ep, rt, prm);
    const int km = c.mode == A_WINDOW ? (c.bkc == 16 ? 3 : 2) : c.mode;
    const int ks = c.bkb / 32;
    using KernelFn = void (*)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap, const IgemmParams,
                              const int32_t*, const float*, void*);
    // last index: 0 streaming B, 1 resident B, 2 streaming B in CTA pairs (not for 16-byte pixels), 3 resident B with the
    // bias-fold variant of the tile loops
    // ... 4: resident B in CTA pairs (window A with >= 32-byte pixels; ring modes with 128-byte channel chunks)
#define LBC_KERNELS(KM_, KS_) {igemm_i8_kernel<KM_, KS_, false, false, false>, igemm_i8_kernel<KM_, KS_, true, false, false>, \
                               (KM_ == 3 ? (KernelFn) nullptr : (KernelFn)igemm_i8_kernel<(KM_ == 3 ? 2 : KM_), KS_, false, true, false>), \
                               igemm_i8_kernel<KM_, KS_, true, false, true>, \
                               ((KM_ == 3 || (KM_ != 2 && KS_ != 4)) ? (KernelFn) nullptr \
                                    : (KernelFn)igemm_i8_kernel<(KM_ == 3 ? 2 : KM_), ((KM_ != 2 && KS_ != 4) ? 4 : KS_), true, true, false>)}
    static const KernelFn table[4][3][5] = {
        {LBC_KERNELS(0, 1), LBC_KERNELS(0, 2), LBC_KERNELS(0, 4)},
        {LBC_KERNELS(1, 1), LBC_KERNELS(1, 2), LBC_KERNELS(1, 4)},
        {LBC_KERNELS(2, 1), LBC_KERNELS(2, 2), LBC_KERNELS(2, 4)},
        {LBC_KERNELS(3, 1), LBC_KERNELS(3, 2), LBC_KERNELS(3, 4)},
    };
#undef LBC_KERNELS
    LBC_REQUIRE(ks == 1 || ks == 2 || ks == 4, LBC_ERR_UNSUPPORTED, "igemm: unsupported K block of %d bytes", c.bkb);
    const KernelFn fn = table[km][ks == 4 ? 2 : ks - 1][c.cta2 ? (c.res_b ? 4 : 2) : c.res_b ? (c.fold ? 3 : 1) : 0];
    LBC_REQUIRE(fn != nullptr, LBC_ERR_UNSUPPORTED, "igemm: no kernel for this configuration");
    {
        int dev_ord = 0;
        LBC_CUDA_TRY(cudaGetDevice(&dev_ord));
        std::lock_guard<std::mutex> lk(g_attr_mu);
        if (first_use_on_device(g_attr_devices, dev_ord)) {
            for (int i = 0; i < 4; ++i)
                for (int j = 0; j < 3; ++j)
                    for (int r = 0; r < 5; ++r)
                        if (table[i][j][r])
                            LBC_CUDA_TRY(cudaFuncSetAttribute(table[i][j][r], cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        }
    }
    // programmatic stream serialisation: this kernel's on-chip prologue may overlap the previous kernel's tail; its
    // griddepcontrol.wait orders every global access after the previous grid (see the kernel)
    cudaLaunchConfig_t lc{};
    lc.gridDim = dim3((unsigned)c.grid);
    lc.blockDim = dim3(kNumThreads);
    lc.dynamicSmemBytes = c.smem_bytes;
    lc.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = c.pdl ? 1 : 0;
    lc.attrs = attr;
    lc.numAttrs = 1;
    if (c.cta2) {
        attr[1].id = cudaLaunchAttributeClusterDimension;
        attr[1].val.clusterDim.x = 2; attr[1].val.clusterDim.y = 1; attr[1].val.clusterDim.z = 1;
        lc.numAttrs = 2;
    }
    const int32_t* bias_p = ep.bias;
    const float* scale_p = ep.scale;
    LBC_CUDA_TRY(cudaLaunchKernelEx(&lc, fn, l.tm_a, l.tm_b, l.tm_out, l.tm_out2, prm, bias_p, scale_p, y));
    LBC_CUDA_TRY(cudaGetLastError());
    return LBC_OK;
}


// ---- fused bottleneck tail: host side --------------------------------------------------------------------------------
bool fused_tail_supported(const ConvGeom& ga, const IgemmConfig& ca, const ConvGeom& gb, const IgemmConfig& cb, std::string* why)
{
    auto no = [&](const char* m) { if (why) *why = m; return false; };
    const lbc_conv_desc& a = ga.d;
    const lbc_conv_desc& b = gb.d;
    if (ca.mode != A_WINDOW || ca.bkc == 16) return no("conv A is not a window-mode layer");
    if (!ca.res_b || ca.tiles_n != 1 || ca.bn != 64 || a.k != 64) return no("conv A needs 64 output channels and a resident filter");
    if (ca.pair || ca.cta2 || ca.res_one) return no("conv A must not run in pairs");
    if (a.out_mode != LBC_OUT_INT8 || b.out_mode != LBC_OUT_INT8) return no("both convolutions must produce int8");
    if (b.r != 1 || b.s != 1 || b.stride_h != 1 || b.stride_w != 1 || b.pad_h != 0 || b.pad_w != 0 || b.groups != 1)
        return no("conv B is not a plain 1x1");
    if (b.c != a.k || b.n != a.n || b.h != ga.p || b.w != ga.q) return no("conv B does not consume conv A's output");
    if (b.k != 256) return no("conv B needs 256 output channels");
    if (cb.mode != A_TILED || cb.bkc != 64 || cb.cblocks != 1) return no("conv B's filter is not packed as [256][64]");
    return true;
}

lbc_status fused_tail_encode(const ConvGeom& ga, const IgemmConfig& ca, const ConvGeom& gb, const DeviceInfo& dev, const int8_t* x,
                             const int8_t* wa_packed, const int8_t* wb_packed, void* y, FusedLaunch* out)
{
    lbc_status st = resolve_driver_entry_points();
    if (st != LBC_OK) return st;
    LBC_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(wa_packed) | reinterpret_cast<uintptr_t>(wb_packed) |
                  reinterpret_cast<uintptr_t>(y)) & 15) == 0,
                LBC_ERR_INVALID_ARG, "fused tail: x, packed weights and y must be 16-byte aligned");
    const lbc_conv_desc& a = ga.d;
    const lbc_conv_desc& b = gb.d;
    FusedLaunch f{};
    f.cfg_a = ca;
    f.k2 = b.k;
    f.relu2 = b.relu;
    // shared memory: [window ring][B1][B2][A2 x 2][output staging: 2 teams x bufs x 16 KB][control]
    f.b1_bytes = (uint32_t)ca.k_blocks * ca.b_block_bytes;
    f.b2_bytes = (uint32_t)b.k * 64u;
    const uint32_t ctl_bytes = round_up((uint32_t)sizeof(FusedCtl), 1024);
    const uint32_t fixed = round_up(f.b1_bytes, 1024) + round_up(f.b2_bytes, 1024) + 2u * kBlockM * 64u + ctl_bytes;
    bool fits = false;
    for (int bufs = 2; bufs >= 1 && !fits; --bufs) {
        const uint32_t stage = 2u * (uint32_t)bufs * kBlockM * 128u;
        if (fixed + stage + 4u * ca.win_stage_bytes > 227u * 1024u) continue;
        const int wins = (int)std::min<uint32_t>(kMaxWinStages, (227u * 1024u - fixed - stage) / ca.win_stage_bytes);
        if (wins < (bufs == 2 ? 5 : 4)) continue;
        f.stage_bufs2 = bufs;
        f.win_stages = wins;
        fits = true;
    }
    LBC_REQUIRE(fits, LBC_ERR_UNSUPPORTED, "fused tail: the window ring does not fit next to both filter matrices");
    f.off_b1 = (uint32_t)f.win_stages * ca.win_stage_bytes;
    f.off_b2 = f.off_b1 + round_up(f.b1_bytes, 1024);
    f.off_a2 = f.off_b2 + round_up(f.b2_bytes, 1024);
    f.off_stage2 = f.off_a2 + 2u * kBlockM * 64u;
    f.off_ctl2 = f.off_stage2 + 2u * (uint32_t)f.stage_bufs2 * kBlockM * 128u;
    f.smem_bytes = f.off_ctl2 + ctl_bytes;
    f.grid = std::min(dev.sm_count > 0 ? dev.sm_count : 148, ca.tiles_m);
    const cuuint32_t ones[4] = {1, 1, 1, 1};
    const CUtensorMapL2promotion promo = CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
    {   // B1: conv A's packed filter matrix [64 rows][packed_row_bytes], box {bkb, 64}
        const cuuint64_t dims[2] = {(cuuint64_t)ca.packed_row_bytes, (cuuint64_t)a.k};
        const cuuint64_t strides[1] = {(cuuint64_t)ca.packed_row_bytes};
        const cuuint32_t box[2] = {(cuuint32_t)ca.bkb, (cuuint32_t)ca.bn};
        CUresult r = g_encode_tiled(&f.tm_b1, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, (void*)wa_packed, dims, strides, box, ones,
                                    CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(ca.bkb), promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        LBC_REQUIRE(r == CUDA_SUCCESS, LBC_ERR_CUDA, "cuTensorMapEncodeTiled(B1) failed: %d", (int)r);
    }
    {   // B2: conv B's packed filter matrix [256 rows][64 B], one box
        const cuuint64_t dims[2] = {64, (cuuint64_t)b.k};
        const cuuint64_t strides[1] = {64};
        const cuuint32_t box[2] = {64, (cuuint32_t)b.k};
        CUresult r = g_encode_tiled(&f.tm_b2, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, (void*)wb_packed, dims, strides, box, ones,
                                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        LBC_REQUIRE(r == CUDA_SUCCESS, LBC_ERR_CUDA, "cuTensorMapEncodeTiled(B2) failed: %d", (int)r);
    }
    {   // A: the halo window of conv A's input (as igemm_encode's window map)
        const cuuint64_t dims[4] = {(cuuint64_t)a.c, (cuuint64_t)a.w, (cuuint64_t)a.h, (cuuint64_t)a.n};
        const cuuint64_t strides[3] = {(cuuint64_t)a.c, (cuuint64_t)a.c * a.w, (cuuint64_t)a.c * a.w * a.h};
        const cuuint32_t box[4] = {(cuuint32_t)ca.bkc, (cuuint32_t)ca.wt, (cuuint32_t)(ca.rows_per_tile + (a.r - 1) * a.dil_h), 1};
        CUresult r = g_encode_tiled(&f.tm_a, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, (void*)x, dims, strides, box, ones,
                                    CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(ca.bkc), promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        LBC_REQUIRE(r == CUDA_SUCCESS, LBC_ERR_CUDA, "cuTensorMapEncodeTiled(A window) failed: %d", (int)r);
    }
    {   // output of conv B: (K2, Q, P, N), one 128-byte panel of a window tile per store
        const cuuint64_t dims[4] = {(cuuint64_t)b.k, (cuuint64_t)gb.q, (cuuint64_t)gb.p, (cuuint64_t)b.n};
        const cuuint64_t strides[3] = {(cuuint64_t)b.k, (cuuint64_t)b.k * gb.q, (cuuint64_t)b.k * gb.q * gb.p};
        const cuuint32_t box[4] = {128, (cuuint32_t)ca.cols_per_tile, (cuuint32_t)ca.rows_per_tile, 1};
        CUresult r = g_encode_tiled(&f.tm_out, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, y, dims, strides, box, ones,
                                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        LBC_REQUIRE(r == CUDA_SUCCESS, LBC_ERR_CUDA, "cuTensorMapEncodeTiled(out) failed: %d", (int)r);
    }
    *out = f;
    return LBC_OK;
}

lbc_status fused_tail_launch(const ConvGeom& ga, const ConvGeom& gb, const FusedLaunch& l, const EpilogueParams& epa,
                             const EpilogueParams& epb, const IgemmRuntime& rt, cudaStream_t stream)
{
    FusedParams fp{};
    IgemmLaunch la{};
    la.cfg = l.cfg_a;
    la.cfg.win_stages = l.win_stages;
    la.cfg.n_mma = 1;
    la.cfg.grid = l.grid;
    la.cfg.tpi = 1;
    la.cfg.team_warps = 8;
    la.reverse = l.reverse;
    fill_params(ga, la, epa, rt, fp.a);
    fp.a.off_b = l.off_b1;
    fp.k2 = l.k2;
    fp.relu2 = l.relu2;
    fp.c_mid = ga.d.k;
    fp.stage_bufs2 = l.stage_bufs2;
    fp.off_b2 = l.off_b2; fp.off_a2 = l.off_a2; fp.off_stage2 = l.off_stage2; fp.off_ctl2 = l.off_ctl2;
    fp.b1_bytes = l.b1_bytes; fp.b2_bytes = l.b2_bytes;
    using FusedFn = void (*)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap, const FusedParams, const int32_t*,
                             const float*, const int32_t*, const float*);
    const FusedFn fn = igemm_fused_tail_kernel;
    {
        static std::mutex mu;
        static uint64_t done[4] = {0, 0, 0, 0};
        int dev_ord = 0;
        LBC_CUDA_TRY(cudaGetDevice(&dev_ord));
        std::lock_guard<std::mutex> lk(mu);
        if (first_use_on_device(done, dev_ord)) {
            LBC_CUDA_TRY(cudaFuncSetAttribute(igemm_fused_tail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        }
    }
    cudaLaunchConfig_t lc{};
    lc.gridDim = dim3((unsigned)l.grid);
    lc.blockDim = dim3(kNumThreads);
    lc.dynamicSmemBytes = l.smem_bytes;
    lc.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = l.cfg_a.pdl ? 1 : 0;
    lc.attrs = attr;
    lc.numAttrs = 1;
    (void)gb;
    LBC_CUDA_TRY(cudaLaunchKernelEx(&lc, fn, l.tm_a, l.tm_b1, l.tm_b2, l.tm_out, fp, epa.bias, epa.scale, epb.bias, epb.scale));
    LBC_CUDA_TRY(cudaGetLastError());
    return LBC_OK;
}

}  // namespace lbc
