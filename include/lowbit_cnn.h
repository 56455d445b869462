/*
 * lowbit_cnn.h — C ABI of liblowbit-cnn (B200 / sm_100a int8 convolution library).
 *
 * This is the drop-in boundary for the reference's int8 convolution path.  The reference
 * (alnfedorov/lowbitdnn-project) has no shared library: its operators are header-only C++
 * templates over at::Tensor that are #included into the drivers.  Every entry point below
 * names the reference interface it replaces (paths relative to the reference root):
 *
 *   lbc_conv_plan_create / lbc_conv_run   <- conv2DForward3x3<batch,inC,outC,inH,inW,outH,outW>(Tensor,Tensor)
 *                                            cpp/int8conv/conv2DForward3x3TensorCores.cuh:695-751 (host wrapper)
 *                                            and the kernel it launches, :537-693
 *   lbc_conv_prepack_weights              <- to_vect_c(kernel) + .contiguous() at
 *                                            cpp/int8conv/check.cu:72,77 and ...TensorCores.cuh:715-716
 *   lbc_to_vect_c / lbc_from_vect_c       <- to_vect_c / from_vect_c, cpp/int8conv/utils.cuh:11-26
 *   epilogue (bias, scale, RNE, clamp)    <- quantize(), cpp/int8conv/conv2DForward3x3WinogradFused.cuh:39-46
 *                                            _Quantize.forward, python/qtorch/nn/functional/quantization.py:27-49
 *   elapsed_ms out-parameter              <- std::tuple<Tensor,float> second element (cudaEvent timing),
 *                                            ...TensorCores.cuh:734-750
 *   lbc_net_*                             <- the conv->relu chains of python/tmp.py:43-56 driven as one unit
 *                                            (benchmark apps: cpp/apps/benchmark.cpp:54-81)
 *
 * Conventions
 *   - Activations: int8, NHWC, dense.  Weights: int8, [K][R][S][C/groups] ("KRSC") or OIHW on input to prepack.
 *   - Accumulation: exact int32 (two's-complement wraparound), cross-correlation (no kernel flip),
 *     same orientation as refConv2DForwardImpl (cpp/int8conv/refConv2DForward.hpp:24-51).
 *   - Epilogue (LBC_OUT_INT8):  t = acc + bias[k] (int32 wraparound); f = (float)t (round-to-nearest-even);
 *     f = f * scale[k] (one fp32 multiply, no FMA); q = rint(f) (round-half-to-even);
 *     q = clamp(q, relu ? 0 : -128, 127); store int8 NHWC.
 *     LBC_OUT_INT32 stores t = acc + bias[k] (bias may be NULL -> 0) as int32 NHWC, no scaling.
 *   - All buffers are caller-owned DEVICE pointers unless the function name ends in _host.
 *   - Every function returns an lbc_status; nothing throws across this boundary.
 *   - There is no CPU fallback: every compute entry point fails with LBC_ERR_NO_DEVICE when no
 *     sm_100 device is usable.
 */
#ifndef LOWBIT_CNN_H
#define LOWBIT_CNN_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LBC_VERSION_MAJOR 0
#define LBC_VERSION_MINOR 1

typedef enum lbc_status {
    LBC_OK = 0,
    LBC_ERR_INVALID_ARG = 1,     /* NULL pointer, zero dimension, inconsistent descriptor            */
    LBC_ERR_UNSUPPORTED = 2,     /* descriptor is valid but no kernel covers it                       */
    LBC_ERR_NO_DEVICE = 3,       /* no CUDA device / not compute capability 10.x                      */
    LBC_ERR_CUDA = 4,            /* a CUDA runtime/driver call failed; see lbc_last_error_string()    */
    LBC_ERR_ALLOC = 5,           /* host or device allocation failed                                  */
    LBC_ERR_KERNEL_TIMEOUT = 6   /* a device-side pipeline wait exceeded its watchdog (bug guard)     */
} lbc_status;

typedef enum lbc_out_mode {
    LBC_OUT_INT8 = 0,            /* fused bias + per-channel scale + RNE + [relu] + saturate -> int8  */
    LBC_OUT_INT32 = 1            /* raw accumulators (+ optional bias) -> int32                       */
} lbc_out_mode;

typedef enum lbc_weight_layout {
    LBC_W_KRSC = 0,              /* [K][R][S][C/groups]  (NHWC-style filters)                         */
    LBC_W_OIHW = 1               /* [K][C/groups][R][S]  (reference layout, refConv2DForward.hpp:27)  */
} lbc_weight_layout;

typedef enum lbc_kernel_kind {
    LBC_KERNEL_AUTO = 0,         /* planner decides                                                   */
    LBC_KERNEL_DIRECT = 1,       /* CUDA-core direct convolution (any shape)                          */
    LBC_KERNEL_IGEMM_TC = 2,     /* tcgen05 implicit GEMM (TMA im2col + TMEM accumulators)            */
    LBC_KERNEL_DEPTHWISE = 3,    /* CUDA-core depthwise (groups == C == K)                            */
    LBC_KERNEL_STEM_TC = 4       /* tiny C (C*stride^2 <= 16): zero-pad + space-to-depth pre-pass into 16-channel
                                    pixels at stride 1, then the tcgen05 kernel's 16-byte-pixel window mode      */
} lbc_kernel_kind;

/* Plain-old-data convolution descriptor.  All sizes in elements. */
typedef struct lbc_conv_desc {
    int32_t n, h, w, c;          /* input  NHWC                                                       */
    int32_t k, r, s;             /* K filters of R x S x (C/groups)                                   */
    int32_t stride_h, stride_w;
    int32_t pad_h, pad_w;        /* symmetric zero padding                                            */
    int32_t dil_h, dil_w;
    int32_t groups;
    int32_t relu;                /* 0/1, only used by LBC_OUT_INT8                                    */
    int32_t out_mode;            /* lbc_out_mode                                                      */
} lbc_conv_desc;

/* Planner options.  Every field has a "planner decides" value, so a zero-initialised struct passed through
 * lbc_plan_options_init() reproduces lbc_conv_plan_create().  They exist for the parity tests (which force every
 * code path on every shape) and for tuning sweeps; the library reads NO environment variables.
 * Tri-state fields: -1 = planner decides, 0 = off, 1 = on (where the mechanism is structurally possible).
 * Limit fields: 0 = planner decides. */
typedef struct lbc_plan_options {
    int32_t struct_size;         /* sizeof(lbc_plan_options); set by lbc_plan_options_init                         */
    int32_t cta_pairs;           /* two-CTA clusters, cta_group::2 MMAs, half of the filter rows per CTA           */
    int32_t warp_store;          /* per-warp staging + TMA stores in the epilogue (ring modes)                      */
    int32_t fold_bias;           /* bias through the first MMA of every tile (resident filter matrix)               */
    int32_t paired_tiles;        /* two M tiles per CTA step share every B block (window A, streaming B); opt-in    */
    int32_t resident_filter;     /* filter matrix kept in shared memory when it fits (2: but not as halves in CTA pairs) */
    int32_t window;              /* shifted-window A operand for stride-1 RxS layers (0: TMA im2col instead)        */
    int32_t keep_window;         /* 1: keep the window mode where the planner would prefer im2col + CTA pairs        */
    int32_t force_im2col;        /* 1: TMA im2col A operand even for pure GEMMs                                     */
    int32_t pixel_groups;        /* pixel-group rewrite of narrow pointwise layers                                  */
    int32_t dw_tiled;            /* TMA-staged depthwise 3x3 kernel (0: direct global-memory kernel)                */
    int32_t reverse;             /* 1: walk the tiles last-to-first (single-layer API; networks alternate)          */
    int32_t pdl;                 /* programmatic dependent launch                                                   */
    int32_t two_mma_warps;       /* 0: a single MMA-issuing warp                                                    */
    int32_t tiles_per_iter2;     /* 0: one tile per epilogue-team iteration on narrow tiles                         */
    int32_t small_teams;         /* 0: 2 x 8-warp epilogue teams even on narrow tiles                               */
    int32_t four_acc;            /* 0: two TMEM accumulator stages where four would fit                             */
    int32_t n_stationary;        /* CTAs keep one N tile of the filter matrix resident and walk M (multi-N-tile)    */
    int32_t epi_pipeline;        /* reserved (the pipelined TMEM drain is a compile-time switch, LBC_EPI_PIPE)      */
    int32_t max_grid;            /* cap on persistent CTAs (tests: many tiles per CTA)                              */
    int32_t max_bn;              /* cap on the N tile width                                                         */
    int32_t max_stages;          /* cap on operand ring stages                                                      */
    int32_t max_win_stages;      /* cap on window ring stages                                                       */
    int32_t stage_bufs;          /* cap on output staging panels per epilogue team (1..3)                           */
    int32_t tps_kb;              /* cap (KB) on the B bytes grouped into one ring stage in window mode              */
    int32_t resident_kb;         /* largest filter matrix (KB) kept resident                                        */
    int32_t epi_split;           /* tri-state: both epilogue teams drain every tile (column split) on > 128-wide tiles */
    int32_t fuse;                /* networks: 1 = run conv(R x S -> 64) -> conv(1x1 -> 256) pairs as one fused launch (opt-in) */
    int32_t early_weights;       /* networks: filter blocks are fetched before the programmatic-dependency wait (2: resident matrices only) */
    int32_t tail_split;          /* CTA pairs: the leftover steps of the last round run as half-width tiles on twice the pairs */
    int32_t reserved[4];
} lbc_plan_options;

/* int8 NHWC pooling window (max-pool).  Output size: 1 + (in + 2*pad - window) / stride, as the reference's cuDNN
 * pooling computes it (python/qtorch/cpp/pool2d.cuh:76-78). */
typedef struct lbc_pool_desc {
    int32_t n, h, w, c;
    int32_t kh, kw;
    int32_t stride_h, stride_w;
    int32_t pad_h, pad_w;
} lbc_pool_desc;

/* One node of a network graph (lbc_net_create_graph). */
typedef enum lbc_node_kind {
    LBC_NODE_CONV = 0,           /* `conv`: a convolution (planned by the tile/layout planner)                      */
    LBC_NODE_MAXPOOL = 1,        /* `pool`: int8 max-pool of the producer's output                                  */
    LBC_NODE_ADD = 2             /* saturating int8 add of two producers (+ ReLU when `relu`): the residual join    */
} lbc_node_kind;
typedef struct lbc_node {
    int32_t kind;                /* lbc_node_kind                                                                   */
    int32_t input_of;            /* producer node index, or -1: fed from outside (network input / resident buffer)  */
    int32_t input2_of;           /* LBC_NODE_ADD: the second operand's producer (>= 0)                              */
    int32_t relu;                /* LBC_NODE_ADD: clamp at zero after the add                                       */
    lbc_conv_desc conv;          /* LBC_NODE_CONV                                                                   */
    lbc_pool_desc pool;          /* LBC_NODE_MAXPOOL (n/h/w/c must equal the producer's output)                     */
} lbc_node;

typedef struct lbc_fused_plan lbc_fused_plan;   /* opaque: two convolutions run as one launch (lbc_fused_tail_*)            */
typedef struct lbc_plan lbc_plan;     /* opaque; immutable after creation and safe to share between threads and
                                         streams (per-run scratch is allocated stream-ordered per call)            */
typedef struct lbc_net  lbc_net;      /* opaque: a fixed chain/list of planned convolutions           */
typedef void* lbc_stream;             /* a cudaStream_t (NULL = legacy default stream)                */

/* ---- library / device --------------------------------------------------------------------------- */
int         lbc_version(void);                       /* major*1000 + minor                            */
const char* lbc_last_error_string(void);             /* thread-local, never NULL                      */
lbc_status  lbc_device_info(int device, int* sm_count, int* cc_major, int* cc_minor, size_t* hbm_bytes);

/* ---- shape helpers (pure host arithmetic; usable without a GPU) ------------------------------- */
/* (in + 2p - (d(k-1)+1))/s + 1 : cpp/int8conv/cudnn2DConvolution.cuh:33-36 */
lbc_status  lbc_conv_out_shape(const lbc_conv_desc* d, int32_t* p, int32_t* q);
/* ops = 2*N*P*Q*K*(C/g)*R*S ; bytes = N*H*W*C + K*(C/g)*R*S + N*P*Q*K*out_elt + 8*K   */
lbc_status  lbc_conv_work(const lbc_conv_desc* d, double* ops, double* bytes);

/* ---- planning --------------------------------------------------------------------------------- */
/* Chooses kernel + tiling for `d` on the current device.  `force` = LBC_KERNEL_AUTO for the planner's
 * choice, or a specific kind (fails with LBC_ERR_UNSUPPORTED if that kernel cannot run the shape). */
lbc_status  lbc_conv_plan_create(const lbc_conv_desc* d, int32_t force, lbc_plan** plan);
/* The same with explicit planner options (NULL = defaults). */
void        lbc_plan_options_init(lbc_plan_options* opt);
lbc_status  lbc_conv_plan_create_ex(const lbc_conv_desc* d, int32_t force, const lbc_plan_options* opt, lbc_plan** plan);
lbc_status  lbc_conv_plan_destroy(lbc_plan* plan);
/* Dry run of the planner for a B200 with `sm_count` SMs (0 = 148): no CUDA call, usable without a GPU.  Reports the
 * kernel kind and the planner's description (tile, K chunk, stages, modes, shared memory) or the reason it refuses. */
lbc_status  lbc_conv_plan_dry(const lbc_conv_desc* d, int32_t force, int32_t sm_count, int32_t* kind, char* buf, size_t buf_len);
lbc_status  lbc_conv_plan_dry_ex(const lbc_conv_desc* d, int32_t force, const lbc_plan_options* opt, int32_t sm_count,
                                 int32_t* kind, char* buf, size_t buf_len);
lbc_status  lbc_conv_plan_kernel(const lbc_plan* plan, int32_t* kind);           /* lbc_kernel_kind   */
lbc_status  lbc_conv_plan_describe(const lbc_plan* plan, char* buf, size_t buf_len); /* human-readable */
/* Number of kernels one lbc_conv_run() launches (for launch accounting in the harness). */
lbc_status  lbc_conv_plan_launches(const lbc_plan* plan, int32_t* launches);

/* ---- weights ---------------------------------------------------------------------------------- */
/* Size in bytes of the packed-weight buffer the plan's kernel wants. */
lbc_status  lbc_conv_packed_weight_bytes(const lbc_plan* plan, size_t* bytes);
/* Re-lays `w_dev` (device, int8, `layout`) into the plan's kernel layout at `dst_dev` (device). */
lbc_status  lbc_conv_prepack_weights(const lbc_plan* plan, const int8_t* w_dev, int32_t layout,
                                     void* dst_dev, lbc_stream stream);

/* ---- execution -------------------------------------------------------------------------------- */
/* Asynchronous on `stream`.  If elapsed_ms != NULL the call brackets the launch with events on
 * `stream`, synchronises, and returns the device time (the reference's `(out, ms)` convention). */
lbc_status  lbc_conv_run(const lbc_plan* plan, const int8_t* x_nhwc, const void* w_packed,
                         const int32_t* bias, const float* scale, void* y_nhwc,
                         lbc_stream stream, float* elapsed_ms);

/* Device-side status of the asynchronous paths: every pipeline wait in the kernels is bounded by a watchdog; a trip
 * sets the plan's (network's) device flag.  Call after synchronising the stream: reads and clears the flag and returns
 * LBC_ERR_KERNEL_TIMEOUT if any launch of this plan since the last check gave up waiting.  The timed / _host entry
 * points check it themselves. */
lbc_status  lbc_conv_plan_check(const lbc_plan* plan);

/* Same, but x / y are HOST buffers (pinned or pageable); H2D and D2H copies are issued on `stream`
 * around the kernel and the call returns after the result is in y_host.  w/bias/scale stay on device. */
lbc_status  lbc_conv_run_host(const lbc_plan* plan, const int8_t* x_host, const void* w_packed,
                              const int32_t* bias, const float* scale, void* y_host,
                              lbc_stream stream, float* elapsed_ms);

/* ---- fused bottleneck tail ----------------------------------------------------------------------- */
/* conv_a (R x S, stride 1, 64 output channels, int8 out) followed by conv_b (1x1, 64 -> 256, int8 out) on conv_a's output,
 * in ONE kernel: conv_a's requantised tile stays in shared memory as the second GEMM's operand, the middle tensor never
 * touches HBM (the conv -> relu -> conv chain of python/tmp.py:43-56; ResNet-50's stage-1 conv2 -> conv3).  The result is
 * bit-identical to running the two convolutions one after the other.  LBC_ERR_UNSUPPORTED for any other pair.
 * The weights are packed with the two single-layer plans (lbc_fused_tail_plan_parts + lbc_conv_prepack_weights). */
lbc_status  lbc_fused_tail_plan_create(const lbc_conv_desc* conv_a, const lbc_conv_desc* conv_b, lbc_fused_plan** plan);
lbc_status  lbc_fused_tail_plan_destroy(lbc_fused_plan* plan);
lbc_status  lbc_fused_tail_plan_parts(const lbc_fused_plan* plan, const lbc_plan** plan_a, const lbc_plan** plan_b);
lbc_status  lbc_fused_tail_run(const lbc_fused_plan* plan, const int8_t* x_nhwc, const void* wa_packed, const int32_t* bias_a,
                               const float* scale_a, const void* wb_packed, const int32_t* bias_b, const float* scale_b,
                               void* y_nhwc, lbc_stream stream, float* elapsed_ms);

/* ---- int8 ops between convolutions (NHWC int8 device buffers) ----------------------------------- */
/* Max-pool: replaces max_pool2d(input, kernel, stride, padding) of python/qtorch/cpp/pool2d.cuh:54-92 (cuDNN
 * CUDNN_POOLING_MAX_DETERMINISTIC on int8): padding never wins, plain integer max. */
lbc_status  lbc_pool_out_shape(const lbc_pool_desc* d, int32_t* p, int32_t* q);
lbc_status  lbc_maxpool2d_run(const lbc_pool_desc* d, const int8_t* x_nhwc, int8_t* y_nhwc, lbc_stream stream);
/* Residual join: y[i] = clamp(a[i] + b[i], relu ? 0 : -128, 127) over n_elements int8 values (y may alias a or b). */
lbc_status  lbc_add_relu_run(const int8_t* a, const int8_t* b, int8_t* y, size_t n_elements, int32_t relu, lbc_stream stream);
/* Global average pool over H*W with the convolutions' requantisation rule: y[n][c] = sat_int8(rint(sum * scale)). */
lbc_status  lbc_global_avgpool_run(const int8_t* x_nhwc, int32_t n, int32_t hw, int32_t c, float scale, int8_t* y_nc,
                                   lbc_stream stream);

/* ---- layout converters (reference tensor formats) -------------------------------------------- */
/* NCHW int8/int32 -> [N][C/V][H][W][V]  (utils.cuh:20-26) and back (utils.cuh:11-17). elt = 1 or 4. */
lbc_status  lbc_to_vect_c(const void* src_nchw, void* dst_vect, int32_t n, int32_t c, int32_t h, int32_t w,
                          int32_t v, int32_t elt_bytes, lbc_stream stream);
lbc_status  lbc_from_vect_c(const void* src_vect, void* dst_nchw, int32_t n, int32_t c, int32_t h, int32_t w,
                            int32_t v, int32_t elt_bytes, lbc_stream stream);
/* NHWC <-> [N][C/V][H][W][V]: the regroup the reference-signature shims need around lbc_conv_run. */
lbc_status  lbc_nhwc_to_vect_c(const void* src_nhwc, void* dst_vect, int32_t n, int32_t c, int32_t h, int32_t w,
                               int32_t v, int32_t elt_bytes, lbc_stream stream);
lbc_status  lbc_vect_c_to_nhwc(const void* src_vect, void* dst_nhwc, int32_t n, int32_t c, int32_t h, int32_t w,
                               int32_t v, int32_t elt_bytes, lbc_stream stream);
/* NCHW <-> NHWC (callers holding refConv2DForward-format tensors). */
lbc_status  lbc_nchw_to_nhwc(const void* src, void* dst, int32_t n, int32_t c, int32_t h, int32_t w,
                             int32_t elt_bytes, lbc_stream stream);
lbc_status  lbc_nhwc_to_nchw(const void* src, void* dst, int32_t n, int32_t c, int32_t h, int32_t w,
                             int32_t elt_bytes, lbc_stream stream);

/* ---- backward passes as int8 convolutions ------------------------------------------------------ */
/* The reference computes both gradients of its QConv2D with the SAME int8 forward convolution on re-laid operands
 * (python/qtorch/nn/functional/qconv2d.py:90-114; the dp4a drafts cpp/int8conv/conv2DBackwardData3x3.cuh:61-64,126-127 and
 * conv2DBackwardWeights3x3.cuh:15-100 state the same sums).  These helpers derive the descriptors and operand layouts so
 * that lbc_conv_plan_create / lbc_conv_run (LBC_OUT_INT32) produce the exact int32 gradients.  Stride 1, dilation 1,
 * groups 1, padding <= filter - 1, as in the reference (qconv2d.py:84-88).
 *   data gradient    dx[n,h,w,c] = sum_{k,r,s} dy[n,h+pad-r,w+pad-s,k] * w[k,r,s,c]
 *                    = conv(dy as input [N,P,Q,K], filter w'[c][r'][s'][k] = w[k][R-1-r'][S-1-s'][c], padding R-1-pad)
 *   weight gradient  dw[k,r,s,c] = sum_{n,p,q} dy[n,p,q,k] * x[n,p-pad+r,q-pad+s,c]
 *                    = conv(x^T as input [C,H,W,N], filter dy^T [K][P][Q][N], padding pad) -> int32 [C][R][S][K];
 *                    lbc_nhwc_to_chwn turns it into [K][R][S][C]. */
lbc_status  lbc_conv_dgrad_desc(const lbc_conv_desc* fwd, lbc_conv_desc* dgrad);
lbc_status  lbc_conv_wgrad_desc(const lbc_conv_desc* fwd, lbc_conv_desc* wgrad);
/* w [K][R][S][C] (device) -> the data-gradient filter [C][R][S][K], rotated by 180 degrees. */
lbc_status  lbc_conv_dgrad_weights(const lbc_conv_desc* fwd, const int8_t* w_krsc, int8_t* w_dgrad, lbc_stream stream);
/* [N][H][W][C] -> [C][H][W][N] (elt_bytes 1 or 4): the operand transposes of the weight gradient, and its result. */
lbc_status  lbc_nhwc_to_chwn(const void* src, void* dst, int32_t n, int32_t h, int32_t w, int32_t c, int32_t elt_bytes,
                             lbc_stream stream);

/* ---- networks: a list of convolutions run back to back (benchmark apps) ------------------------ */
/* `input_of[i]` = index of the layer whose OUTPUT feeds layer i, or -1 for the network input.  The
 * network owns its packed weights, bias, scale and activation buffers (synthetic or caller-loaded). */
lbc_status  lbc_net_create(const lbc_conv_desc* descs, const int32_t* input_of, int32_t n_layers, lbc_net** net);
/* With planner options applied to every layer (NULL = defaults); opt->reverse == 0 switches the alternating traversal
 * direction off. */
lbc_status  lbc_net_create_ex(const lbc_conv_desc* descs, const int32_t* input_of, int32_t n_layers,
                              const lbc_plan_options* opt, lbc_net** net);
lbc_status  lbc_net_check(lbc_net* net);              /* as lbc_conv_plan_check, for every layer of the network */
/* A network as a graph of convolutions, max-pools and residual adds (a whole int8 ResNet runs on the device: the conv ->
 * pool -> relu chains of python/tmp.py:43-56 and the bottleneck's residual join).  Nodes must be listed in topological
 * order.  Every lbc_net_* call below works on graph networks; per-layer parameters only exist for LBC_NODE_CONV nodes. */
lbc_status  lbc_net_create_graph(const lbc_node* nodes, int32_t n_nodes, const lbc_plan_options* opt, lbc_net** net);
lbc_status  lbc_net_destroy(lbc_net* net);
lbc_status  lbc_net_layer_plan(const lbc_net* net, int32_t layer, const lbc_plan** plan);
/* *into = the layer whose fused launch absorbs `layer` (its own output is then never materialised), or -1. */
lbc_status  lbc_net_layer_fused_into(const lbc_net* net, int32_t layer, int32_t* into);
/* Load parameters for one layer from HOST memory (weights in `layout`, bias int32[K], scale f32[K]). */
lbc_status  lbc_net_set_params_host(lbc_net* net, int32_t layer, const int8_t* w_host, int32_t layout,
                                    const int32_t* bias_host, const float* scale_host);
/* Fill the resident input buffer of a layer whose input_of == -1 from HOST memory (NHWC int8). */
lbc_status  lbc_net_set_input_host(lbc_net* net, int32_t layer, const int8_t* x_host);
/* Copy the first `max_bytes` bytes (0 = all) of one layer's resident output (NHWC, int8 or int32 per its descriptor:
 * image-major, so a prefix is a whole number of images) to HOST memory; synchronises the device. */
lbc_status  lbc_net_read_output_host(const lbc_net* net, int32_t layer, void* y_host, size_t max_bytes);
/* Device pointers of a layer's input/output activations (valid until lbc_net_destroy). */
lbc_status  lbc_net_layer_io(const lbc_net* net, int32_t layer, const void** x_dev, void** y_dev);
/* Run all layers on `stream` with the network input already resident in HBM (x_dev NHWC of layer(s)
 * with input_of == -1).  per_layer_ms (may be NULL) receives n_layers event-timed durations. */
lbc_status  lbc_net_run(lbc_net* net, const int8_t* x_dev, lbc_stream stream, float* per_layer_ms, float* total_ms);
/* End to end: x_host -> (H2D) -> all layers -> (D2H) -> y_host (the last layer's output). */
lbc_status  lbc_net_run_host(lbc_net* net, const int8_t* x_host, void* y_host, lbc_stream stream, float* total_ms);
/* Pipelined end to end: enqueue H2D(x_host) -> all layers -> D2H(y_host) and return without waiting.  The copies
 * run on the network's own copy streams and the network input is double-buffered, so the upload of step i+1 and the
 * download of step i overlap the layers of the neighbouring steps (a serving loop's steady state).  x_host / y_host
 * must stay valid - and be pinned for the copies to overlap - until lbc_net_sync_host() returns, which waits for
 * everything submitted and reports the device time from the first submit to the last download. */
lbc_status  lbc_net_submit_host(lbc_net* net, const int8_t* x_host, void* y_host, lbc_stream stream);
lbc_status  lbc_net_sync_host(lbc_net* net, float* elapsed_ms);
lbc_status  lbc_net_launches(const lbc_net* net, int32_t* launches);

/* ---- measurement helpers (cpp/libbenchmark role) ---------------------------------------------- */
/* MMA-only tcgen05 kind::i8 peak probe: returns achieved dense int8 TOPS on the current device. */
lbc_status  lbc_probe_int8_mma_peak(int32_t iters, double* tops, lbc_stream stream);
/* Streaming-copy probe (int4 loads/stores), GB/s read+write. */
lbc_status  lbc_probe_hbm_copy(size_t bytes, int32_t iters, double* gbs, lbc_stream stream);
/* Development aid, per plan: when device_buf != NULL, CTA 0 of the plan's tcgen05 kernel records clock64 stamps of its
 * pipeline events (16 int64 slots per local tile, `tiles` tiles, then 2 x grid CTA start/end stamps, then 4 x grid
 * %globaltimer stamps: kernel entry, set-up done, past the dependency wait, last store issued) into it.
 * NULL switches tracing off.  Not thread-safe against concurrent runs of the same plan.  The stamps are compiled in only
 * with -DLBC_TRACE=1 (lib/liblowbit_cnn_trace.so, built next to the product library); the product build returns
 * LBC_ERR_UNSUPPORTED for a non-NULL buffer. */
lbc_status  lbc_conv_plan_set_trace(lbc_plan* plan, void* device_buf, int32_t tiles);
/* Writes `bytes` of zeros to an internal scratch buffer to evict L2 (timing hygiene). */
lbc_status  lbc_flush_l2(lbc_stream stream);

#ifdef __cplusplus
}
#endif
#endif /* LOWBIT_CNN_H */
