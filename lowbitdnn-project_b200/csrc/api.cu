// api.cu — the extern "C" surface declared in include/lowbit_cnn.h: planner, weight pre-pack, run,
// layout converters, network runner and probes.  Host-side C++; every compute path ends in a CUDA launch
// (there is no CPU fallback — without an sm_100 device every compute entry point returns LBC_ERR_NO_DEVICE).
#include "common.cuh"

#include <stdarg.h>
#include <string.h>

#include <array>
#include <mutex>
#include <string>
#include <utility>
#include <vector>

namespace lbc {

// ---- errors ------------------------------------------------------------------------------------------
static thread_local char t_err[512] = "no error";

void set_error(const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(t_err, sizeof t_err, fmt, ap);
    va_end(ap);
}
const char* last_error() { return t_err; }

// ---- geometry ----------------------------------------------------------------------------------------
static int32_t out_dim(int32_t in, int32_t pad, int32_t dil, int32_t k, int32_t stride)
{
    return (in + 2 * pad - (dil * (k - 1) + 1)) / stride + 1;   // cudnn2DConvolution.cuh:33-36
}

lbc_status make_geom(const lbc_conv_desc* d, ConvGeom* g)
{
    LBC_REQUIRE(d && g, LBC_ERR_INVALID_ARG, "null descriptor");
    LBC_REQUIRE(d->n > 0 && d->h > 0 && d->w > 0 && d->c > 0 && d->k > 0 && d->r > 0 && d->s > 0,
                LBC_ERR_INVALID_ARG, "descriptor has a non-positive extent");
    LBC_REQUIRE(d->stride_h > 0 && d->stride_w > 0 && d->dil_h > 0 && d->dil_w > 0 && d->pad_h >= 0 && d->pad_w >= 0,
                LBC_ERR_INVALID_ARG, "descriptor has a bad stride/dilation/padding");
    LBC_REQUIRE(d->groups > 0 && d->c % d->groups == 0 && d->k % d->groups == 0, LBC_ERR_INVALID_ARG,
                "groups must divide C and K");
    LBC_REQUIRE(d->out_mode == LBC_OUT_INT8 || d->out_mode == LBC_OUT_INT32, LBC_ERR_INVALID_ARG, "bad out_mode");
    g->d = *d;
    g->p = out_dim(d->h, d->pad_h, d->dil_h, d->r, d->stride_h);
    g->q = out_dim(d->w, d->pad_w, d->dil_w, d->s, d->stride_w);
    LBC_REQUIRE(g->p > 0 && g->q > 0, LBC_ERR_INVALID_ARG, "empty output (%d x %d)", g->p, g->q);
    LBC_REQUIRE(d->h + 2 * d->pad_h >= d->dil_h * (d->r - 1) + 1 && d->w + 2 * d->pad_w >= d->dil_w * (d->s - 1) + 1,
                LBC_ERR_INVALID_ARG, "filter larger than the padded input");
    g->cg = d->c / d->groups;
    g->kg = d->k / d->groups;
    g->m_total = (int64_t)d->n * g->p * g->q;
    return LBC_OK;
}

// ---- device ------------------------------------------------------------------------------------------
lbc_status current_device(DeviceInfo* info)
{
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        cudaGetLastError();
        set_error("no CUDA device (%s); liblowbit-cnn has no CPU fallback", e == cudaSuccess ? "count == 0" : cudaGetErrorString(e));
        return LBC_ERR_NO_DEVICE;
    }
    int dev = 0;
    LBC_CUDA_TRY(cudaGetDevice(&dev));
    static std::mutex mu;
    static std::vector<DeviceInfo> cache;
    std::lock_guard<std::mutex> lk(mu);
    if ((int)cache.size() <= dev) cache.resize(dev + 1);
    if (cache[dev].device < 0) {
        cudaDeviceProp p;
        LBC_CUDA_TRY(cudaGetDeviceProperties(&p, dev));
        DeviceInfo di;
        di.device = dev;
        di.sm_count = p.multiProcessorCount;
        di.cc_major = p.major;
        di.cc_minor = p.minor;
        di.hbm_bytes = p.totalGlobalMem;
        cudaDriverGetVersion(&di.driver_version);
        cache[dev] = di;
    }
    *info = cache[dev];
    if (info->cc_major != 10) {
        set_error("device %d is sm_%d%d; liblowbit-cnn is built for sm_100a only", dev, info->cc_major, info->cc_minor);
        return LBC_ERR_NO_DEVICE;
    }
    return LBC_OK;
}

bool first_use_on_device(uint64_t (&mask)[4], int device)
{
    if (device < 0 || device >= 256) return true;      // out of the bitmap: set the attributes every time
    const bool first = !((mask[device >> 6] >> (device & 63)) & 1ull);
    mask[device >> 6] |= 1ull << (device & 63);
    return first;
}

// planner defaults: every tri-state "planner decides", every limit "default"
void default_options(lbc_plan_options* o)
{
    memset(o, 0, sizeof *o);
    o->struct_size = (int32_t)sizeof *o;
    o->cta_pairs = o->warp_store = o->fold_bias = o->resident_filter = o->window = o->pixel_groups = o->dw_tiled = -1;
    o->paired_tiles = 0;
    o->keep_window = o->force_im2col = 0;
    o->reverse = -1;
    o->pdl = o->two_mma_warps = o->tiles_per_iter2 = o->small_teams = o->four_acc = -1;
    o->n_stationary = o->epi_pipeline = o->epi_split = o->fuse = -1;
    o->early_weights = -1;
    o->tail_split = -1;
}

}  // namespace lbc

using namespace lbc;

// ---- opaque types ------------------------------------------------------------------------------------
// A fully resolved launch: everything needed to put the layer on a stream.
struct ResolvedLaunch {
    const lbc_plan* plan = nullptr;
    const int8_t* x = nullptr;
    const void* w = nullptr;
    EpilogueParams ep{};
    void* y = nullptr;
    IgemmLaunch ig{};
    DwLaunch dw{};
};

struct lbc_plan {
    ConvGeom g;
    int32_t kind;
    IgemmConfig cfg;
    DeviceInfo dev;
    lbc_plan_options opt;
    // device watchdog word: owned by the plan, or borrowed from the network the plan belongs to
    int* flag = nullptr;
    bool own_flag = false;
    // pipeline trace (development aid, lbc_conv_plan_set_trace)
    long long* trace = nullptr;
    int32_t trace_tiles = 0;
    // LBC_KERNEL_STEM_TC, and LBC_KERNEL_IGEMM_TC with pw_factor > 1: the rewritten problem the tcgen05 kernel runs
    ConvGeom g_inner;
    // Pixel-group rewrite of a pointwise (1x1, stride 1) layer whose C or K is not a multiple of 16 (MobileNetV2's
    // 24-channel tensors): f consecutive pixels form one GEMM row of f*C channels against the block-diagonal
    // [f*K][f*C] filter, so every TMA stride is a multiple of 16 bytes again and the [M/f][f*K] result IS the NHWC
    // output.  The MMA does f times the useful work, which is free on these HBM-bound layers.
    int32_t pw_factor = 1;
    int32_t stem_sh = 1, stem_sw = 1;
    // [N][Hs][Ws][16] transformed input of the small-C path, owned by the plan.  It is the one piece of mutable state a
    // run touches: runs of the same plan on different streams are ordered on the device through `stem_last` (the
    // second stream waits for the first run's igemm before its transform overwrites the buffer); no host blocking.
    void* stem_x = nullptr;
    mutable std::mutex stem_mu;
    mutable cudaEvent_t stem_last = nullptr;
    mutable cudaStream_t stem_stream = nullptr;
    mutable bool stem_used = false;
    // encoded launches (three CUtensorMaps each) of the most recent (x, w, bias, scale, y) argument sets of lbc_conv_run
    mutable std::mutex cache_mu;
    mutable std::vector<ResolvedLaunch> cache;
    // scratch for lbc_conv_run_host
    mutable std::mutex mu;
    mutable void* x_dev = nullptr;
    mutable void* y_dev = nullptr;
};

struct lbc_fused_plan {
    lbc_plan* a = nullptr;       // the R x S convolution (window mode, resident filter)
    lbc_plan* b = nullptr;       // the 1x1 convolution
    mutable std::mutex cache_mu;
    mutable std::vector<std::pair<std::array<const void*, 4>, FusedLaunch>> cache;   // (x, wa, wb, y) -> encoded launch
};

namespace {

size_t out_elt(const ConvGeom& g) { return g.d.out_mode == LBC_OUT_INT32 ? 4 : 1; }
size_t in_bytes(const ConvGeom& g) { return (size_t)g.d.n * g.d.h * g.d.w * g.d.c; }
size_t out_bytes(const ConvGeom& g) { return (size_t)g.m_total * g.d.k * out_elt(g); }

size_t packed_weight_bytes(const lbc_plan* p)
{
    const lbc_conv_desc& d = p->g.d;
    switch (p->kind) {
        case LBC_KERNEL_IGEMM_TC:
        case LBC_KERNEL_STEM_TC: return (size_t)d.k * p->pw_factor * p->cfg.packed_row_bytes;
        case LBC_KERNEL_DEPTHWISE: return (size_t)d.r * d.s * d.c;
        default: return (size_t)d.k * d.r * d.s * p->g.cg;
    }
}

lbc_status resolve(const lbc_plan* plan, const int8_t* x, const void* w, const int32_t* bias, const float* scale,
                   void* y, ResolvedLaunch* out)
{
    LBC_REQUIRE(plan && x && w && y, LBC_ERR_INVALID_ARG, "lbc_conv_run: null plan/x/w/y");
    LBC_REQUIRE(plan->g.d.out_mode == LBC_OUT_INT32 || scale, LBC_ERR_INVALID_ARG,
                "lbc_conv_run: int8 output needs a per-channel scale");
    out->plan = plan;
    out->x = x;
    out->w = w;
    out->y = y;
    out->ep.bias = bias;
    out->ep.scale = scale;
    out->ep.relu = plan->g.d.relu;
    out->ep.out_mode = plan->g.d.out_mode;
    if (plan->kind == LBC_KERNEL_IGEMM_TC)
        return igemm_encode(plan->pw_factor > 1 ? plan->g_inner : plan->g, plan->cfg, plan->dev, x, (const int8_t*)w, y, &out->ig);
    if (plan->kind == LBC_KERNEL_STEM_TC)
        return igemm_encode(plan->g_inner, plan->cfg, plan->dev, (const int8_t*)plan->stem_x, (const int8_t*)w, y, &out->ig);
    if (plan->kind == LBC_KERNEL_DEPTHWISE) return depthwise_encode(plan->g, plan->opt, x, &out->dw);
    return LBC_OK;
}

// lbc_conv_run: the encoded launch for this argument set, from the plan's small cache when it has been seen before
// (cuTensorMapEncode* costs a few microseconds per map; a serving loop calls with the same buffers every step)
lbc_status resolve_cached(const lbc_plan* plan, const int8_t* x, const void* w, const int32_t* bias, const float* scale,
                          void* y, ResolvedLaunch* out)
{
    {
        std::lock_guard<std::mutex> lk(plan->cache_mu);
        for (const ResolvedLaunch& c : plan->cache)
            if (c.x == x && c.w == w && c.y == y && c.ep.bias == bias && c.ep.scale == scale) {
                *out = c;
                return LBC_OK;
            }
    }
    lbc_status st = resolve(plan, x, w, bias, scale, y, out);
    if (st != LBC_OK) return st;
    std::lock_guard<std::mutex> lk(plan->cache_mu);
    if (plan->cache.size() >= 4) plan->cache.erase(plan->cache.begin());
    plan->cache.push_back(*out);
    return LBC_OK;
}

IgemmRuntime runtime_of(const lbc_plan* p)
{
    IgemmRuntime rt;
    rt.flag = p->flag;
    rt.trace = p->trace;
    rt.trace_tiles = p->trace_tiles;
    return rt;
}

// reads and clears a device watchdog word
lbc_status check_flag(int* flag)
{
    if (!flag) return LBC_OK;
    int v = 0;
    LBC_CUDA_TRY(cudaMemcpy(&v, flag, sizeof v, cudaMemcpyDeviceToHost));
    if (v) {
        cudaMemset(flag, 0, sizeof v);
        set_error(v == 2 ? "igemm: dynamic shared memory base is not 1024-byte aligned"
                         : "device pipeline watchdog fired (a bounded mbarrier wait exceeded 2 s)");
        return LBC_ERR_KERNEL_TIMEOUT;
    }
    return LBC_OK;
}

lbc_status launch(const ResolvedLaunch& l, cudaStream_t stream)
{
    const lbc_plan* p = l.plan;
    switch (p->kind) {
        case LBC_KERNEL_IGEMM_TC: return igemm_launch(p->pw_factor > 1 ? p->g_inner : p->g, l.ig, l.ep, l.y, runtime_of(p), stream);
        case LBC_KERNEL_STEM_TC: {
            const lbc_conv_desc& d = p->g.d;
            // the transformed-input scratch is per plan: order this run after the previous one if that was on another stream
            std::lock_guard<std::mutex> lk(p->stem_mu);
            if (p->stem_used && p->stem_stream != stream) LBC_CUDA_TRY(cudaStreamWaitEvent(stream, p->stem_last, 0));
            lbc_status st = launch_stem_xform(l.x, p->stem_x, d.n, d.h, d.w, d.c, p->g_inner.d.h, p->g_inner.d.w, p->stem_sh,
                                              p->stem_sw, d.pad_h, d.pad_w, stream);
            if (st != LBC_OK) return st;
            st = igemm_launch(p->g_inner, l.ig, l.ep, l.y, runtime_of(p), stream);
            if (st != LBC_OK) return st;
            LBC_CUDA_TRY(cudaEventRecord(p->stem_last, stream));
            p->stem_stream = stream;
            p->stem_used = true;
            return LBC_OK;
        }
        case LBC_KERNEL_DEPTHWISE: return launch_depthwise(p->g, l.x, (const int8_t*)l.w, l.ep, l.y, &l.dw, p->flag, stream);
        case LBC_KERNEL_DIRECT: return launch_direct_conv(p->g, l.x, (const int8_t*)l.w, l.ep, l.y, stream);
        default: set_error("plan has unknown kernel kind %d", p->kind); return LBC_ERR_UNSUPPORTED;
    }
}

struct EventPair {
    cudaEvent_t a = nullptr, b = nullptr;
    lbc_status init()
    {
        LBC_CUDA_TRY(cudaEventCreate(&a));
        LBC_CUDA_TRY(cudaEventCreate(&b));
        return LBC_OK;
    }
    ~EventPair()
    {
        if (a) cudaEventDestroy(a);
        if (b) cudaEventDestroy(b);
    }
};

}  // namespace

// ======================================================================================================
extern "C" {

int lbc_version(void) { return LBC_VERSION_MAJOR * 1000 + LBC_VERSION_MINOR; }
const char* lbc_last_error_string(void) { return last_error(); }

lbc_status lbc_device_info(int device, int* sm_count, int* cc_major, int* cc_minor, size_t* hbm_bytes)
{
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) {
        cudaGetLastError();
        set_error("no CUDA device %d", device);
        return LBC_ERR_NO_DEVICE;
    }
    cudaDeviceProp p;
    LBC_CUDA_TRY(cudaGetDeviceProperties(&p, device));
    if (sm_count) *sm_count = p.multiProcessorCount;
    if (cc_major) *cc_major = p.major;
    if (cc_minor) *cc_minor = p.minor;
    if (hbm_bytes) *hbm_bytes = p.totalGlobalMem;
    return LBC_OK;
}

lbc_status lbc_conv_out_shape(const lbc_conv_desc* d, int32_t* p, int32_t* q)
{
    ConvGeom g;
    lbc_status st = make_geom(d, &g);
    if (st != LBC_OK) return st;
    if (p) *p = g.p;
    if (q) *q = g.q;
    return LBC_OK;
}

lbc_status lbc_conv_work(const lbc_conv_desc* d, double* ops, double* bytes)
{
    ConvGeom g;
    lbc_status st = make_geom(d, &g);
    if (st != LBC_OK) return st;
    if (ops) *ops = 2.0 * (double)g.m_total * d->k * g.cg * d->r * d->s;
    if (bytes)
        *bytes = (double)in_bytes(g) + (double)d->k * g.cg * d->r * d->s + (double)out_bytes(g) + 8.0 * d->k;
    return LBC_OK;
}

// The planner proper: pure host logic for a given device description.  `dry` plans without touching CUDA (no scratch
// allocation): lbc_conv_plan_dry() uses it to make every tiling decision checkable on a machine without a GPU.
static lbc_status plan_build(const lbc_conv_desc* d, int32_t force, const DeviceInfo& dev, const lbc_plan_options* opt_in,
                             bool dry, int* shared_flag, lbc_plan** plan)
{
    LBC_REQUIRE(plan, LBC_ERR_INVALID_ARG, "null plan out-pointer");
    *plan = nullptr;
    lbc_plan_options opt;
    default_options(&opt);
    if (opt_in) {
        LBC_REQUIRE(opt_in->struct_size == (int32_t)sizeof(lbc_plan_options), LBC_ERR_INVALID_ARG,
                    "lbc_plan_options.struct_size is %d, this library expects %d (call lbc_plan_options_init first)",
                    opt_in->struct_size, (int)sizeof(lbc_plan_options));
        opt = *opt_in;
    }
    ConvGeom g;
    lbc_status st = make_geom(d, &g);
    if (st != LBC_OK) return st;

    std::string why;
    const bool tc_ok = igemm_supported(g, &why);
    const bool dw_ok = (d->groups == d->c && d->k == d->c && d->c % 4 == 0);
    // small-C rewrite (stems): C*stride^2 <= 16 channels after zero-pad + space-to-depth
    ConvGeom gi{};
    IgemmConfig stem_cfg{};
    bool stem_ok = d->groups == 1 && d->c < 16 && d->k % 16 == 0 && d->dil_h == 1 && d->dil_w == 1 &&
                   d->stride_h == d->stride_w && (d->stride_h == 1 || d->stride_h == 2) &&
                   d->c * d->stride_h * d->stride_w <= 16;
    if (stem_ok) {
        lbc_conv_desc di = *d;
        const int sh = d->stride_h, sw = d->stride_w;
        di.r = (d->r + sh - 1) / sh;
        di.s = (d->s + sw - 1) / sw;
        di.h = g.p + di.r - 1;
        di.w = g.q + di.s - 1;
        di.c = 16;
        di.stride_h = di.stride_w = 1;
        di.pad_h = di.pad_w = 0;
        stem_ok = make_geom(&di, &gi) == LBC_OK && gi.p == g.p && gi.q == g.q && igemm_supported(gi, nullptr) &&
                  igemm_make_config(gi, dev, opt, &stem_cfg) == LBC_OK && stem_cfg.mode == 2 && stem_cfg.bkc == 16;
    }
    // pixel-group rewrite for pointwise layers the tensor-core path cannot tile directly
    ConvGeom gp{};
    IgemmConfig pw_cfg{};
    int32_t pw_factor = 1;
    // ... and for narrow pointwise layers it CAN tile: a 128 x 16 output tile pays the per-tile costs (barriers, waits,
    // TMA issue) for 2048 outputs; grouping f pixels makes the tile 128 x 16f over the same bytes.
    const bool pointwise = d->groups == 1 && d->r == 1 && d->s == 1 && d->stride_h == 1 && d->stride_w == 1 &&
                           d->pad_h == 0 && d->pad_w == 0;
    const bool narrow = tc_ok && pointwise && (d->k < 64 || d->c < 32) && opt.pixel_groups != 0;
    if (pointwise && (!tc_ok || narrow)) {
        for (int32_t f = 2; f <= 16 && pw_factor == 1; f *= 2) {
            if ((f * d->c) % 16 || (f * d->k) % 16 || g.m_total % f || (int64_t)f * d->c > 4096) continue;
            if (narrow && (f * d->k > 256 || f * d->c > 512)) break;
            if (narrow && (f * d->k < 64 || f * d->c < 32) && 2 * f * d->k <= 256 && 2 * f * d->c <= 512 && g.m_total % (2 * f) == 0)
                continue;   // a larger group still fits: keep growing
            lbc_conv_desc di = *d;
            di.n = 1; di.h = 1; di.w = (int32_t)(g.m_total / f); di.c = f * d->c; di.k = f * d->k;
            if (g.m_total / f >= (1ll << 31)) continue;
            if (make_geom(&di, &gp) == LBC_OK && igemm_supported(gp, nullptr) && igemm_make_config(gp, dev, opt, &pw_cfg) == LBC_OK) {
                pw_factor = f;
                pw_cfg.k_mod = d->k;
            }
        }
    }
    const bool pw_ok = pw_factor > 1;
    bool use_pw = pw_ok && !tc_ok;   // a forced igemm_tc on a shape it tiles directly stays un-grouped
    int32_t kind = force;
    if (force == LBC_KERNEL_AUTO) {
        // Tile/layout planner: tensor cores for dense contractions, CUDA cores where they do not pay.
        if (dw_ok) kind = LBC_KERNEL_DEPTHWISE;
        else if (tc_ok || pw_ok) kind = LBC_KERNEL_IGEMM_TC;
        else if (stem_ok) kind = LBC_KERNEL_STEM_TC;
        else kind = LBC_KERNEL_DIRECT;
        use_pw = pw_ok;
    }
    LBC_REQUIRE(kind == LBC_KERNEL_DIRECT || kind == LBC_KERNEL_IGEMM_TC || kind == LBC_KERNEL_DEPTHWISE ||
                    kind == LBC_KERNEL_STEM_TC,
                LBC_ERR_UNSUPPORTED, "kernel kind %d is not available", kind);
    LBC_REQUIRE(kind != LBC_KERNEL_STEM_TC || stem_ok, LBC_ERR_UNSUPPORTED,
                "small-C tensor-core path needs groups 1, C*stride^2 <= 16, stride 1 or 2, dilation 1, K %% 16 == 0");
    LBC_REQUIRE(kind != LBC_KERNEL_IGEMM_TC || tc_ok || pw_ok, LBC_ERR_UNSUPPORTED,
                "tcgen05 implicit GEMM cannot run this shape: %s", why.c_str());
    LBC_REQUIRE(kind != LBC_KERNEL_DEPTHWISE || dw_ok, LBC_ERR_UNSUPPORTED,
                "depthwise kernel needs groups == C == K and C %% 4 == 0");

    lbc_plan* p = new (std::nothrow) lbc_plan();
    LBC_REQUIRE(p, LBC_ERR_ALLOC, "out of host memory");
    p->g = g;
    p->kind = kind;
    p->dev = dev;
    p->opt = opt;
    if (shared_flag) {
        p->flag = shared_flag;
    } else if (!dry) {
        if (cudaMalloc((void**)&p->flag, sizeof(int)) != cudaSuccess || cudaMemset(p->flag, 0, sizeof(int)) != cudaSuccess) {
            cudaGetLastError();
            delete p;
            set_error("cannot allocate the plan's device status word");
            return LBC_ERR_ALLOC;
        }
        p->own_flag = true;
    }
    if (kind == LBC_KERNEL_IGEMM_TC && use_pw) {
        p->pw_factor = pw_factor;
        p->g_inner = gp;
        p->cfg = pw_cfg;
    } else if (kind == LBC_KERNEL_IGEMM_TC) {
        st = igemm_make_config(g, dev, opt, &p->cfg);
        if (st != LBC_OK) {
            lbc_conv_plan_destroy(p);
            return st;
        }
    }
    if (kind == LBC_KERNEL_STEM_TC) {
        p->g_inner = gi;
        p->cfg = stem_cfg;
        p->stem_sh = d->stride_h;
        p->stem_sw = d->stride_w;
        const size_t bytes = (size_t)gi.d.n * gi.d.h * gi.d.w * 16;
        if (!dry && (cudaMalloc(&p->stem_x, bytes) != cudaSuccess ||
                     cudaEventCreateWithFlags(&p->stem_last, cudaEventDisableTiming) != cudaSuccess)) {
            cudaGetLastError();
            lbc_conv_plan_destroy(p);
            set_error("small-C path: cannot allocate the %zu-byte transformed input", bytes);
            return LBC_ERR_ALLOC;
        }
    }
    *plan = p;
    return LBC_OK;
}

void lbc_plan_options_init(lbc_plan_options* opt)
{
    if (opt) default_options(opt);
}

lbc_status lbc_conv_plan_create_ex(const lbc_conv_desc* d, int32_t force, const lbc_plan_options* opt, lbc_plan** plan)
{
    LBC_REQUIRE(plan, LBC_ERR_INVALID_ARG, "null plan out-pointer");
    *plan = nullptr;
    DeviceInfo dev;
    lbc_status st = current_device(&dev);
    if (st != LBC_OK) return st;
    return plan_build(d, force, dev, opt, false, nullptr, plan);
}

lbc_status lbc_conv_plan_create(const lbc_conv_desc* d, int32_t force, lbc_plan** plan)
{
    return lbc_conv_plan_create_ex(d, force, nullptr, plan);
}

lbc_status lbc_conv_plan_dry(const lbc_conv_desc* d, int32_t force, int32_t sm_count, int32_t* kind, char* buf, size_t buf_len)
{
    return lbc_conv_plan_dry_ex(d, force, nullptr, sm_count, kind, buf, buf_len);
}

lbc_status lbc_conv_plan_dry_ex(const lbc_conv_desc* d, int32_t force, const lbc_plan_options* opt, int32_t sm_count,
                                int32_t* kind, char* buf, size_t buf_len)
{
    DeviceInfo dev;
    dev.device = 0;
    dev.sm_count = sm_count > 0 ? sm_count : 148;
    dev.cc_major = 10;
    dev.cc_minor = 0;
    dev.hbm_bytes = (size_t)180 << 30;
    dev.driver_version = 13000;
    lbc_plan* p = nullptr;
    lbc_status st = plan_build(d, force, dev, opt, true, nullptr, &p);
    if (st != LBC_OK) return st;
    if (kind) *kind = p->kind;
    if (buf && buf_len) st = lbc_conv_plan_describe(p, buf, buf_len);
    delete p;
    return st;
}

lbc_status lbc_conv_plan_destroy(lbc_plan* plan)
{
    if (!plan) return LBC_OK;
    if (plan->x_dev) cudaFree(plan->x_dev);
    if (plan->y_dev) cudaFree(plan->y_dev);
    if (plan->stem_x) cudaFree(plan->stem_x);
    if (plan->stem_last) cudaEventDestroy(plan->stem_last);
    if (plan->own_flag && plan->flag) cudaFree(plan->flag);
    delete plan;
    return LBC_OK;
}

lbc_status lbc_conv_plan_set_trace(lbc_plan* plan, void* device_buf, int32_t tiles)
{
    LBC_REQUIRE(plan, LBC_ERR_INVALID_ARG, "null plan");
    LBC_REQUIRE(!device_buf || igemm_trace_compiled(), LBC_ERR_UNSUPPORTED,
                "pipeline tracing is compiled out of this build: load lib/liblowbit_cnn_trace.so (built with -DLBC_TRACE=1)");
    plan->trace = device_buf ? reinterpret_cast<long long*>(device_buf) : nullptr;
    plan->trace_tiles = device_buf ? tiles : 0;
    return LBC_OK;
}

lbc_status lbc_conv_plan_check(const lbc_plan* plan)
{
    LBC_REQUIRE(plan, LBC_ERR_INVALID_ARG, "null plan");
    return check_flag(plan->flag);
}

lbc_status lbc_conv_plan_kernel(const lbc_plan* plan, int32_t* kind)
{
    LBC_REQUIRE(plan && kind, LBC_ERR_INVALID_ARG, "null argument");
    *kind = plan->kind;
    return LBC_OK;
}

lbc_status lbc_conv_plan_describe(const lbc_plan* plan, char* buf, size_t buf_len)
{
    LBC_REQUIRE(plan && buf && buf_len > 0, LBC_ERR_INVALID_ARG, "null argument");
    const lbc_conv_desc& d = plan->g.d;
    if (plan->kind == LBC_KERNEL_IGEMM_TC || plan->kind == LBC_KERNEL_STEM_TC) {
        const IgemmConfig& c = plan->cfg;
        static const char* modes[] = {"tiled", "im2col", "window"};
        snprintf(buf, buf_len,
                 "%s N%d %dx%dx%d->%d %dx%d s%d p%d | M=%lld tile 128x%d kchunk %dB x%d kblocks stages %dx%d+%dw b=%s "
                 "a=%s(%dx%d px/tile) tiles %dx%d grid %d smem %zu tmem %u sbufs %d epi=%s",
                 plan->kind == LBC_KERNEL_STEM_TC ? "stem_tc(s2d->16ch)" : plan->pw_factor > 1 ? "igemm_tc(pixel-groups)" : "igemm_tc",
                 d.n, d.h, d.w, d.c, d.k, d.r, d.s,
                 d.stride_h, d.pad_h, (long long)plan->g.m_total, c.bn, c.bkc, c.k_blocks, c.stages, c.tps, c.win_stages,
                 c.res_b ? (c.cta2 ? "resident,cta-pair" : c.res_one ? "resident,n-stationary" : c.n_mma == 2 ? "resident,2mma" : "resident") : c.pair ? "ring,paired-tiles" : c.cta2 ? "ring,cta-pair" : "ring", modes[c.mode], c.rows_per_tile, c.cols_per_tile, c.tiles_m,
                 c.tiles_n, c.grid, c.smem_bytes, c.tmem_cols, c.stage_bufs,
                 d.out_mode != LBC_OUT_INT8 ? (c.fold ? "int32,bias-in-mma" : "int32")
                 : (c.warp_store && c.team_warps == 4) ? (c.fold ? "narrow-warp-stores,bias-in-mma" : "narrow-warp-stores")
                 : c.warp_store ? (c.epi_split ? (c.fold ? "warp-stores,split,bias-in-mma" : "warp-stores,split")
                                               : (c.fold ? "warp-stores,bias-in-mma" : "warp-stores"))
                 : c.tail_first >= 0 ? "2x8-warp-teams,tail-split"
                 : (c.epi_split && c.team_warps == 8) ? (c.fold ? "2x8-warp-teams,split,bias-in-mma" : "2x8-warp-teams,split")
                 : c.team_warps == 4 ? (c.fold ? "4x4-warp-teams,bias-in-mma" : "4x4-warp-teams")
                                     : (c.fold ? "2x8-warp-teams,bias-in-mma" : "2x8-warp-teams"));
    } else {
        snprintf(buf, buf_len, "%s N%d %dx%dx%d->%d %dx%d s%d p%d g%d | M=%lld",
                 plan->kind == LBC_KERNEL_DEPTHWISE ? "depthwise" : "direct", d.n, d.h, d.w, d.c, d.k, d.r, d.s,
                 d.stride_h, d.pad_h, d.groups, (long long)plan->g.m_total);
    }
    return LBC_OK;
}

lbc_status lbc_conv_plan_launches(const lbc_plan* plan, int32_t* launches)
{
    LBC_REQUIRE(plan && launches, LBC_ERR_INVALID_ARG, "null argument");
    *launches = plan->kind == LBC_KERNEL_STEM_TC ? 2 : 1;
    return LBC_OK;
}

lbc_status lbc_conv_packed_weight_bytes(const lbc_plan* plan, size_t* bytes)
{
    LBC_REQUIRE(plan && bytes, LBC_ERR_INVALID_ARG, "null argument");
    *bytes = packed_weight_bytes(plan);
    return LBC_OK;
}

lbc_status lbc_conv_prepack_weights(const lbc_plan* plan, const int8_t* w_dev, int32_t layout, void* dst_dev,
                                    lbc_stream stream)
{
    LBC_REQUIRE(plan && w_dev && dst_dev, LBC_ERR_INVALID_ARG, "null argument");
    LBC_REQUIRE(layout == LBC_W_KRSC || layout == LBC_W_OIHW, LBC_ERR_INVALID_ARG, "bad weight layout %d", layout);
    const lbc_conv_desc& d = plan->g.d;
    cudaStream_t s = (cudaStream_t)stream;
    switch (plan->kind) {
        case LBC_KERNEL_IGEMM_TC:
            if (plan->pw_factor > 1) {
                // [K][C] (KRSC and OIHW coincide for 1x1) -> block-diagonal [f*K][f*C] -> the kernel's filter matrix
                const int32_t f = plan->pw_factor;
                int8_t* tmp = nullptr;
                LBC_CUDA_TRY(cudaMallocAsync((void**)&tmp, (size_t)f * d.k * f * d.c, s));
                lbc_status st = launch_blockdiag(w_dev, tmp, d.k, d.c, f, s);
                if (st == LBC_OK)
                    st = launch_prepack_igemm(tmp, LBC_W_KRSC, (int8_t*)dst_dev, f * d.k, 1, 1, f * d.c, 1, plan->cfg.bkc,
                                              plan->cfg.cblocks, 0, s);
                cudaFreeAsync(tmp, s);
                return st;
            }
            return launch_prepack_igemm(w_dev, layout, (int8_t*)dst_dev, d.k, d.r, d.s, plan->g.cg, plan->cfg.s_pad,
                                        plan->cfg.bkc, plan->cfg.cblocks, plan->cfg.mode == 2, s);
        case LBC_KERNEL_STEM_TC:
            return launch_prepack_stem(w_dev, layout, (int8_t*)dst_dev, d.k, d.r, d.s, d.c, plan->stem_sh, plan->stem_sw,
                                       plan->g_inner.d.r, plan->cfg.s_pad, s);
        case LBC_KERNEL_DEPTHWISE:
            return launch_prepack_depthwise(w_dev, layout, (int8_t*)dst_dev, d.c, d.r, d.s, s);
        default:
            return launch_prepack_krsc(w_dev, layout, (int8_t*)dst_dev, d.k, d.r, d.s, plan->g.cg, plan->g.cg, s);
    }
}

lbc_status lbc_conv_run(const lbc_plan* plan, const int8_t* x, const void* w_packed, const int32_t* bias,
                        const float* scale, void* y, lbc_stream stream, float* elapsed_ms)
{
    ResolvedLaunch l;
    lbc_status st = resolve_cached(plan, x, w_packed, bias, scale, y, &l);
    if (st != LBC_OK) return st;
    cudaStream_t s = (cudaStream_t)stream;
    if (!elapsed_ms) return launch(l, s);
    EventPair ev;
    st = ev.init();
    if (st != LBC_OK) return st;
    LBC_CUDA_TRY(cudaEventRecord(ev.a, s));
    st = launch(l, s);
    if (st != LBC_OK) return st;
    LBC_CUDA_TRY(cudaEventRecord(ev.b, s));
    LBC_CUDA_TRY(cudaEventSynchronize(ev.b));
    LBC_CUDA_TRY(cudaEventElapsedTime(elapsed_ms, ev.a, ev.b));
    return check_flag(plan->flag);
}

lbc_status lbc_conv_run_host(const lbc_plan* plan, const int8_t* x_host, const void* w_packed, const int32_t* bias,
                             const float* scale, void* y_host, lbc_stream stream, float* elapsed_ms)
{
    LBC_REQUIRE(plan && x_host && y_host, LBC_ERR_INVALID_ARG, "null argument");
    std::lock_guard<std::mutex> lk(plan->mu);
    if (!plan->x_dev) {
        if (cudaMalloc(&plan->x_dev, in_bytes(plan->g)) != cudaSuccess || cudaMalloc(&plan->y_dev, out_bytes(plan->g)) != cudaSuccess) {
            cudaGetLastError();
            set_error("lbc_conv_run_host: device allocation failed");
            return LBC_ERR_ALLOC;
        }
    }
    cudaStream_t s = (cudaStream_t)stream;
    EventPair ev;
    lbc_status st = ev.init();
    if (st != LBC_OK) return st;
    ResolvedLaunch l;
    st = resolve(plan, (const int8_t*)plan->x_dev, w_packed, bias, scale, plan->y_dev, &l);
    if (st != LBC_OK) return st;
    LBC_CUDA_TRY(cudaEventRecord(ev.a, s));
    LBC_CUDA_TRY(cudaMemcpyAsync(plan->x_dev, x_host, in_bytes(plan->g), cudaMemcpyHostToDevice, s));
    st = launch(l, s);
    if (st != LBC_OK) return st;
    LBC_CUDA_TRY(cudaMemcpyAsync(y_host, plan->y_dev, out_bytes(plan->g), cudaMemcpyDeviceToHost, s));
    LBC_CUDA_TRY(cudaEventRecord(ev.b, s));
    LBC_CUDA_TRY(cudaEventSynchronize(ev.b));
    if (elapsed_ms) LBC_CUDA_TRY(cudaEventElapsedTime(elapsed_ms, ev.a, ev.b));
    return check_flag(plan->flag);
}

// ---- fused bottleneck tail ----------------------------------------------------------------------------
lbc_status lbc_fused_tail_plan_destroy(lbc_fused_plan* plan)
{
    if (!plan) return LBC_OK;
    lbc_conv_plan_destroy(plan->a);
    lbc_conv_plan_destroy(plan->b);
    delete plan;
    return LBC_OK;
}

lbc_status lbc_fused_tail_plan_create(const lbc_conv_desc* conv_a, const lbc_conv_desc* conv_b, lbc_fused_plan** out)
{
    LBC_REQUIRE(conv_a && conv_b && out, LBC_ERR_INVALID_ARG, "null argument");
    *out = nullptr;
    lbc_fused_plan* f = new (std::nothrow) lbc_fused_plan();
    LBC_REQUIRE(f, LBC_ERR_ALLOC, "out of host memory");
    lbc_status st = lbc_conv_plan_create(conv_a, LBC_KERNEL_AUTO, &f->a);
    if (st == LBC_OK) st = lbc_conv_plan_create(conv_b, LBC_KERNEL_AUTO, &f->b);
    if (st == LBC_OK) {
        std::string why;
        if (f->a->kind != LBC_KERNEL_IGEMM_TC || f->b->kind != LBC_KERNEL_IGEMM_TC || f->a->pw_factor != 1 || f->b->pw_factor != 1 ||
            !fused_tail_supported(f->a->g, f->a->cfg, f->b->g, f->b->cfg, &why)) {
            set_error("this pair of convolutions cannot run fused: %s", why.empty() ? "not tensor-core layers" : why.c_str());
            st = LBC_ERR_UNSUPPORTED;
        }
    }
    if (st != LBC_OK) {
        char keep[512];
        strncpy(keep, last_error(), sizeof keep);
        keep[sizeof keep - 1] = 0;
        lbc_fused_tail_plan_destroy(f);
        set_error("%s", keep);
        return st;
    }
    *out = f;
    return LBC_OK;
}

lbc_status lbc_fused_tail_plan_parts(const lbc_fused_plan* plan, const lbc_plan** plan_a, const lbc_plan** plan_b)
{
    LBC_REQUIRE(plan, LBC_ERR_INVALID_ARG, "null plan");
    if (plan_a) *plan_a = plan->a;
    if (plan_b) *plan_b = plan->b;
    return LBC_OK;
}

lbc_status lbc_fused_tail_run(const lbc_fused_plan* plan, const int8_t* x, const void* wa, const int32_t* bias_a, const float* scale_a,
                              const void* wb, const int32_t* bias_b, const float* scale_b, void* y, lbc_stream stream, float* elapsed_ms)
{
    LBC_REQUIRE(plan && x && wa && wb && y && scale_a && scale_b, LBC_ERR_INVALID_ARG, "lbc_fused_tail_run: null argument");
    FusedLaunch fl;
    bool hit = false;
    const std::array<const void*, 4> key = {x, wa, wb, y};
    {
        std::lock_guard<std::mutex> lk(plan->cache_mu);
        for (const auto& c : plan->cache)
            if (c.first == key) { fl = c.second; hit = true; break; }
    }
    if (!hit) {
        lbc_status st = fused_tail_encode(plan->a->g, plan->a->cfg, plan->b->g, plan->a->dev, x, (const int8_t*)wa, (const int8_t*)wb, y, &fl);
        if (st != LBC_OK) return st;
        std::lock_guard<std::mutex> lk(plan->cache_mu);
        if (plan->cache.size() >= 4) plan->cache.erase(plan->cache.begin());
        plan->cache.emplace_back(key, fl);
    }
    EpilogueParams epa{bias_a, scale_a, plan->a->g.d.relu, LBC_OUT_INT8}, epb{bias_b, scale_b, plan->b->g.d.relu, LBC_OUT_INT8};
    cudaStream_t s = (cudaStream_t)stream;
    if (!elapsed_ms) return fused_tail_launch(plan->a->g, plan->b->g, fl, epa, epb, runtime_of(plan->a), s);
    EventPair ev;
    lbc_status st = ev.init();
    if (st != LBC_OK) return st;
    LBC_CUDA_TRY(cudaEventRecord(ev.a, s));
    st = fused_tail_launch(plan->a->g, plan->b->g, fl, epa, epb, runtime_of(plan->a), s);
    if (st != LBC_OK) return st;
    LBC_CUDA_TRY(cudaEventRecord(ev.b, s));
    LBC_CUDA_TRY(cudaEventSynchronize(ev.b));
    LBC_CUDA_TRY(cudaEventElapsedTime(elapsed_ms, ev.a, ev.b));
    return check_flag(plan->a->flag);
}

// ---- int8 ops between convolutions --------------------------------------------------------------------
static lbc_status pool_geom(const lbc_pool_desc* d, int32_t* p, int32_t* q)
{
    LBC_REQUIRE(d, LBC_ERR_INVALID_ARG, "null pooling descriptor");
    LBC_REQUIRE(d->n > 0 && d->h > 0 && d->w > 0 && d->c > 0 && d->kh > 0 && d->kw > 0 && d->stride_h > 0 && d->stride_w > 0 &&
                    d->pad_h >= 0 && d->pad_w >= 0,
                LBC_ERR_INVALID_ARG, "pooling descriptor has a bad extent");
    LBC_REQUIRE(d->pad_h < d->kh && d->pad_w < d->kw, LBC_ERR_INVALID_ARG, "pooling padding must be smaller than the window");
    LBC_REQUIRE(d->h + 2 * d->pad_h >= d->kh && d->w + 2 * d->pad_w >= d->kw, LBC_ERR_INVALID_ARG, "pooling window larger than the padded input");
    *p = 1 + (d->h + 2 * d->pad_h - d->kh) / d->stride_h;      // cudnnGetPooling2dForwardOutputDim (pool2d.cuh:76-78)
    *q = 1 + (d->w + 2 * d->pad_w - d->kw) / d->stride_w;
    return LBC_OK;
}

lbc_status lbc_pool_out_shape(const lbc_pool_desc* d, int32_t* p, int32_t* q)
{
    int32_t pp = 0, qq = 0;
    lbc_status st = pool_geom(d, &pp, &qq);
    if (st != LBC_OK) return st;
    if (p) *p = pp;
    if (q) *q = qq;
    return LBC_OK;
}

lbc_status lbc_maxpool2d_run(const lbc_pool_desc* d, const int8_t* x, int8_t* y, lbc_stream stream)
{
    int32_t p = 0, q = 0;
    lbc_status st = pool_geom(d, &p, &q);
    if (st != LBC_OK) return st;
    LBC_REQUIRE(x && y, LBC_ERR_INVALID_ARG, "lbc_maxpool2d_run: null x/y");
    DeviceInfo dev;
    st = current_device(&dev);
    if (st != LBC_OK) return st;
    return launch_maxpool(*d, p, q, x, y, dev.sm_count, (cudaStream_t)stream);
}

lbc_status lbc_add_relu_run(const int8_t* a, const int8_t* b, int8_t* y, size_t n_elements, int32_t relu, lbc_stream stream)
{
    LBC_REQUIRE(a && b && y && n_elements > 0, LBC_ERR_INVALID_ARG, "lbc_add_relu_run: null operand or empty tensor");
    DeviceInfo dev;
    lbc_status st = current_device(&dev);
    if (st != LBC_OK) return st;
    return launch_add_relu(a, b, y, n_elements, relu, dev.sm_count, (cudaStream_t)stream);
}

lbc_status lbc_global_avgpool_run(const int8_t* x, int32_t n, int32_t hw, int32_t c, float scale, int8_t* y, lbc_stream stream)
{
    LBC_REQUIRE(x && y && n > 0 && hw > 0 && c > 0, LBC_ERR_INVALID_ARG, "lbc_global_avgpool_run: bad argument");
    LBC_REQUIRE(n <= 65535, LBC_ERR_UNSUPPORTED, "lbc_global_avgpool_run: more than 65535 images");
    DeviceInfo dev;
    lbc_status st = current_device(&dev);
    if (st != LBC_OK) return st;
    return launch_global_avgpool(x, y, n, hw, c, scale, (cudaStream_t)stream);
}

// ---- layout converters -------------------------------------------------------------------------------
static lbc_status permute_checked(const void* src, void* dst, int32_t d0, int32_t d1, int32_t d2, int32_t d3,
                                  int32_t d4, int p0, int p1, int p2, int p3, int p4, int32_t elt, lbc_stream stream)
{
    DeviceInfo dev;
    lbc_status st = current_device(&dev);
    if (st != LBC_OK) return st;
    const int32_t dims[5] = {d0, d1, d2, d3, d4};
    const int32_t perm[5] = {p0, p1, p2, p3, p4};
    return launch_permute5(src, dst, dims, perm, elt, (cudaStream_t)stream);
}

#define LBC_VECT_ARGS_OK()                                                                                  \
    LBC_REQUIRE(n > 0 && c > 0 && h > 0 && w > 0 && v > 0 && c % v == 0, LBC_ERR_INVALID_ARG,               \
                "vect_c: C (%d) must be a positive multiple of V (%d)", c, v)

lbc_status lbc_to_vect_c(const void* src, void* dst, int32_t n, int32_t c, int32_t h, int32_t w, int32_t v,
                         int32_t elt, lbc_stream stream)
{   // [N, C/V, V, H, W] -> [N, C/V, H, W, V]   (utils.cuh:20-26)
    LBC_VECT_ARGS_OK();
    return permute_checked(src, dst, n, c / v, v, h, w, 0, 1, 3, 4, 2, elt, stream);
}

lbc_status lbc_from_vect_c(const void* src, void* dst, int32_t n, int32_t c, int32_t h, int32_t w, int32_t v,
                           int32_t elt, lbc_stream stream)
{   // [N, C/V, H, W, V] -> [N, C/V, V, H, W]   (utils.cuh:11-17)
    LBC_VECT_ARGS_OK();
    return permute_checked(src, dst, n, c / v, h, w, v, 0, 1, 4, 2, 3, elt, stream);
}

lbc_status lbc_nhwc_to_vect_c(const void* src, void* dst, int32_t n, int32_t c, int32_t h, int32_t w, int32_t v,
                              int32_t elt, lbc_stream stream)
{   // [N, H, W, C/V, V] -> [N, C/V, H, W, V]
    LBC_VECT_ARGS_OK();
    return permute_checked(src, dst, n, h, w, c / v, v, 0, 3, 1, 2, 4, elt, stream);
}

lbc_status lbc_vect_c_to_nhwc(const void* src, void* dst, int32_t n, int32_t c, int32_t h, int32_t w, int32_t v,
                              int32_t elt, lbc_stream stream)
{   // [N, C/V, H, W, V] -> [N, H, W, C/V, V]
    LBC_VECT_ARGS_OK();
    return permute_checked(src, dst, n, c / v, h, w, v, 0, 2, 3, 1, 4, elt, stream);
}

lbc_status lbc_nchw_to_nhwc(const void* src, void* dst, int32_t n, int32_t c, int32_t h, int32_t w, int32_t elt,
                            lbc_stream stream)
{
    LBC_REQUIRE(n > 0 && c > 0 && h > 0 && w > 0, LBC_ERR_INVALID_ARG, "bad extents");
    return permute_checked(src, dst, n, c, h, w, 1, 0, 2, 3, 1, 4, elt, stream);
}

lbc_status lbc_nhwc_to_nchw(const void* src, void* dst, int32_t n, int32_t c, int32_t h, int32_t w, int32_t elt,
                            lbc_stream stream)
{
    LBC_REQUIRE(n > 0 && c > 0 && h > 0 && w > 0, LBC_ERR_INVALID_ARG, "bad extents");
    return permute_checked(src, dst, n, h, w, c, 1, 0, 3, 1, 2, 4, elt, stream);
}

// ---- backward passes as int8 convolutions --------------------------------------------------------------
static lbc_status backward_geom(const lbc_conv_desc* fwd, ConvGeom* g)
{
    lbc_status st = make_geom(fwd, g);
    if (st != LBC_OK) return st;
    const lbc_conv_desc& d = *fwd;
    LBC_REQUIRE(d.stride_h == 1 && d.stride_w == 1 && d.dil_h == 1 && d.dil_w == 1 && d.groups == 1, LBC_ERR_UNSUPPORTED,
                "backward convolutions need stride 1, dilation 1, groups 1 (as the reference: qconv2d.py:84-88)");
    LBC_REQUIRE(d.pad_h <= d.r - 1 && d.pad_w <= d.s - 1, LBC_ERR_UNSUPPORTED, "backward convolutions need padding <= filter - 1");
    return LBC_OK;
}

lbc_status lbc_conv_dgrad_desc(const lbc_conv_desc* fwd, lbc_conv_desc* dg)
{
    LBC_REQUIRE(fwd && dg, LBC_ERR_INVALID_ARG, "null descriptor");
    ConvGeom g;
    lbc_status st = backward_geom(fwd, &g);
    if (st != LBC_OK) return st;
    *dg = *fwd;
    dg->h = g.p; dg->w = g.q; dg->c = fwd->k; dg->k = fwd->c;
    dg->pad_h = fwd->r - 1 - fwd->pad_h;
    dg->pad_w = fwd->s - 1 - fwd->pad_w;
    dg->relu = 0;
    dg->out_mode = LBC_OUT_INT32;
    return LBC_OK;
}

lbc_status lbc_conv_wgrad_desc(const lbc_conv_desc* fwd, lbc_conv_desc* wg)
{
    LBC_REQUIRE(fwd && wg, LBC_ERR_INVALID_ARG, "null descriptor");
    ConvGeom g;
    lbc_status st = backward_geom(fwd, &g);
    if (st != LBC_OK) return st;
    *wg = *fwd;
    wg->n = fwd->c; wg->c = fwd->n; wg->k = fwd->k;     // "images" = input channels, channels = batch
    wg->r = g.p; wg->s = g.q;                           // the filter is dy: P x Q taps
    wg->relu = 0;
    wg->out_mode = LBC_OUT_INT32;
    return LBC_OK;
}

lbc_status lbc_conv_dgrad_weights(const lbc_conv_desc* fwd, const int8_t* w_krsc, int8_t* w_dgrad, lbc_stream stream)
{
    LBC_REQUIRE(fwd && w_krsc && w_dgrad, LBC_ERR_INVALID_ARG, "null argument");
    ConvGeom g;
    lbc_status st = backward_geom(fwd, &g);
    if (st != LBC_OK) return st;
    DeviceInfo dev;
    st = current_device(&dev);
    if (st != LBC_OK) return st;
    return launch_dgrad_weights(w_krsc, w_dgrad, fwd->k, fwd->r, fwd->s, fwd->c, (cudaStream_t)stream);
}

lbc_status lbc_nhwc_to_chwn(const void* src, void* dst, int32_t n, int32_t h, int32_t w, int32_t c, int32_t elt, lbc_stream stream)
{
    LBC_REQUIRE(n > 0 && c > 0 && h > 0 && w > 0, LBC_ERR_INVALID_ARG, "bad extents");
    return permute_checked(src, dst, n, h, w, c, 1, 3, 1, 2, 0, 4, elt, stream);
}

// ---- probes ------------------------------------------------------------------------------------------
lbc_status lbc_probe_int8_mma_peak(int32_t iters, double* tops, lbc_stream stream)
{
    return probe_int8_mma_peak(iters, tops, (cudaStream_t)stream);
}
lbc_status lbc_probe_hbm_copy(size_t bytes, int32_t iters, double* gbs, lbc_stream stream)
{
    return probe_hbm_copy(bytes, iters, gbs, (cudaStream_t)stream);
}
lbc_status lbc_flush_l2(lbc_stream stream)
{
    DeviceInfo dev;
    lbc_status st = current_device(&dev);
    if (st != LBC_OK) return st;
    return flush_l2((cudaStream_t)stream);
}

}  // extern "C"

// ======================================================================================================
// Networks
// ======================================================================================================
struct lbc_net {
    struct Layer {
        int32_t kind = LBC_NODE_CONV;
        lbc_plan* plan = nullptr;        // LBC_NODE_CONV only
        lbc_pool_desc pool{};            // LBC_NODE_MAXPOOL
        int32_t add_relu = 0;            // LBC_NODE_ADD
        int32_t input2_of = -1;          // LBC_NODE_ADD: second operand
        // fused bottleneck tail: `fused_into` >= 0 -> this convolution is absorbed by that (1x1) layer's launch and its own
        // output is never written; `fused_from` >= 0 on the absorbing layer, whose launch then reads `fused_from`'s input
        int32_t fused_into = -1, fused_from = -1;
        FusedLaunch fl{};
        const void* fl_x = nullptr;
        // activation geometry of every node kind: input NHWC (conv / pool; add: both operands) and output NHWC
        int32_t in_n = 0, in_h = 0, in_w = 0, in_c = 0, out_p = 0, out_q = 0, out_k = 0, out_elt = 1;
        size_t in_bytes() const { return (size_t)in_n * in_h * in_w * in_c; }
        size_t out_bytes() const { return (size_t)in_n * out_p * out_q * out_k * out_elt; }
        int32_t input_of = -1;
        void* w = nullptr;       // packed weights
        int32_t* bias = nullptr;
        float* scale = nullptr;
        void* x_own = nullptr;   // own input buffer when input_of == -1
        void* y = nullptr;
        ResolvedLaunch rl;
        bool resolved = false;
        // tcgen05 layers alternate their traversal direction along producer -> consumer edges, so a consumer starts on
        // the images its producer wrote last (still in L2) instead of on the ones written first (evicted)
        bool reverse = false;
    };
    std::vector<Layer> layers;
    std::vector<cudaEvent_t> events;   // n_layers + 1
    std::mutex mu;
    int* flag = nullptr;               // device watchdog word shared by every layer's plan
    int sm_count = 148;
    // layer 0 resolved against each of the pipelined host path's two input buffers (see Pipe)
    ResolvedLaunch first_rl[2];
    bool first_rl_ok[2] = {false, false};
    // pipelined host path (lbc_net_submit_host): copy streams, second input buffer, hand-over events
    struct Pipe {
        cudaStream_t h2d = nullptr, d2h = nullptr;
        void* x_buf[2] = {nullptr, nullptr};          // [0] aliases layer 0's resident input buffer
        cudaEvent_t up_done[2] = {nullptr, nullptr};  // upload into x_buf[i] finished
        cudaEvent_t x_used[2] = {nullptr, nullptr};   // layer 0 has consumed x_buf[i]
        cudaEvent_t comp_done = nullptr, d2h_done = nullptr, t_first = nullptr, t_last = nullptr;
        uint64_t submitted = 0;
        bool ready = false;
    } pipe;
};

namespace {

const void* layer_input(const lbc_net* net, int i)
{
    const lbc_net::Layer& L = net->layers[i];
    return L.input_of < 0 ? L.x_own : net->layers[L.input_of].y;
}

lbc_status net_resolve(lbc_net* net, int i, const int8_t* x_override)
{
    lbc_net::Layer& L = net->layers[i];
    if (L.kind != LBC_NODE_CONV || L.fused_into >= 0 || L.fused_from >= 0) return LBC_OK;   // nothing to encode here
    const int8_t* x = x_override ? x_override : (const int8_t*)layer_input(net, i);
    if (L.resolved && L.rl.x == x) return LBC_OK;
    lbc_status st = resolve(L.plan, x, L.w, L.bias, L.scale, L.y, &L.rl);
    if (st == LBC_OK) {
        L.resolved = true;
        L.rl.ig.reverse = L.reverse ? 1 : 0;
        L.rl.ig.early_b = L.plan->opt.early_weights == 2 ? 2 : L.plan->opt.early_weights != 0 ? 1 : 0;
        L.rl.dw.reverse = (L.reverse && L.rl.dw.tiled) ? 1 : 0;
    }
    return st;
}

// put node i on the stream; `rl` = the encoded launch to use for a convolution node (its own, or layer 0's per-buffer one),
// `x_override` = the input buffer of a non-conv node 0 in the pipelined host path
lbc_status net_launch_node(lbc_net* net, int i, const ResolvedLaunch* rl, const void* x_override, cudaStream_t s)
{
    lbc_net::Layer& L = net->layers[i];
    switch (L.kind) {
        case LBC_NODE_CONV: {
            if (L.fused_into >= 0) return LBC_OK;                       // absorbed by the consumer's fused launch
            if (L.fused_from >= 0) {
                lbc_net::Layer& A = net->layers[L.fused_from];
                const void* x = (L.fused_from == 0 && x_override) ? x_override : layer_input(net, L.fused_from);
                if (L.fl_x != x) {
                    lbc_status st = fused_tail_encode(A.plan->g, A.plan->cfg, L.plan->g, A.plan->dev, (const int8_t*)x, (const int8_t*)A.w,
                                                      (const int8_t*)L.w, L.y, &L.fl);
                    if (st != LBC_OK) return st;
                    L.fl.reverse = A.reverse ? 1 : 0;
                    L.fl_x = x;
                }
                const EpilogueParams epa{A.bias, A.scale, A.plan->g.d.relu, LBC_OUT_INT8}, epb{L.bias, L.scale, L.plan->g.d.relu, LBC_OUT_INT8};
                return fused_tail_launch(A.plan->g, L.plan->g, L.fl, epa, epb, runtime_of(A.plan), s);
            }
            return launch(rl ? *rl : L.rl, s);
        }
        case LBC_NODE_MAXPOOL:
            return launch_maxpool(L.pool, L.out_p, L.out_q, (const int8_t*)(x_override ? x_override : layer_input(net, i)), (int8_t*)L.y,
                                  net->sm_count, s);
        case LBC_NODE_ADD:
            return launch_add_relu((const int8_t*)layer_input(net, i), (const int8_t*)net->layers[L.input2_of].y, (int8_t*)L.y,
                                   L.out_bytes(), L.add_relu, net->sm_count, s);
        default: set_error("node %d has unknown kind %d", i, L.kind); return LBC_ERR_INVALID_ARG;
    }
}

lbc_status net_enqueue(lbc_net* net, const int8_t* x_dev, cudaStream_t s, bool timed)
{
    const int n = (int)net->layers.size();
    if (timed) LBC_CUDA_TRY(cudaEventRecord(net->events[0], s));
    for (int i = 0; i < n; ++i) {
        lbc_status st = net_resolve(net, i, (i == 0 && x_dev) ? x_dev : nullptr);
        if (st != LBC_OK) return st;
        st = net_launch_node(net, i, nullptr, (i == 0 && x_dev) ? x_dev : nullptr, s);
        if (st != LBC_OK) return st;
        if (timed) LBC_CUDA_TRY(cudaEventRecord(net->events[i + 1], s));
    }
    return LBC_OK;
}

}  // namespace

extern "C" {

lbc_status lbc_net_destroy(lbc_net* net)
{
    if (!net) return LBC_OK;
    for (auto& L : net->layers) {
        if (L.w) cudaFree(L.w);
        if (L.bias) cudaFree(L.bias);
        if (L.scale) cudaFree(L.scale);
        if (L.x_own) cudaFree(L.x_own);
        if (L.y) cudaFree(L.y);
        lbc_conv_plan_destroy(L.plan);
    }
    for (auto e : net->events)
        if (e) cudaEventDestroy(e);
    if (net->flag) cudaFree(net->flag);
    {
        lbc_net::Pipe& pp = net->pipe;
        if (pp.h2d) cudaStreamDestroy(pp.h2d);
        if (pp.d2h) cudaStreamDestroy(pp.d2h);
        if (pp.x_buf[1]) cudaFree(pp.x_buf[1]);
        for (cudaEvent_t e : {pp.up_done[0], pp.up_done[1], pp.x_used[0], pp.x_used[1], pp.comp_done, pp.d2h_done, pp.t_first, pp.t_last})
            if (e) cudaEventDestroy(e);
    }
    delete net;
    return LBC_OK;
}

lbc_status lbc_net_create(const lbc_conv_desc* descs, const int32_t* input_of, int32_t n_layers, lbc_net** out)
{
    return lbc_net_create_ex(descs, input_of, n_layers, nullptr, out);
}

lbc_status lbc_net_create_ex(const lbc_conv_desc* descs, const int32_t* input_of, int32_t n_layers, const lbc_plan_options* opt_in,
                             lbc_net** out)
{
    LBC_REQUIRE(descs && out && n_layers > 0, LBC_ERR_INVALID_ARG, "null argument");
    std::vector<lbc_node> nodes((size_t)n_layers);
    for (int i = 0; i < n_layers; ++i) {
        lbc_node nd{};
        nd.kind = LBC_NODE_CONV;
        nd.input_of = input_of ? input_of[i] : (i == 0 ? -1 : i - 1);
        nd.input2_of = -1;
        nd.conv = descs[i];
        nodes[(size_t)i] = nd;
    }
    return lbc_net_create_graph(nodes.data(), n_layers, opt_in, out);
}

lbc_status lbc_net_create_graph(const lbc_node* nodes, int32_t n_layers, const lbc_plan_options* opt_in, lbc_net** out)
{
    LBC_REQUIRE(nodes && out && n_layers > 0, LBC_ERR_INVALID_ARG, "null argument");
    *out = nullptr;
    DeviceInfo dev;
    lbc_status st = current_device(&dev);
    if (st != LBC_OK) return st;
    lbc_plan_options opt;
    default_options(&opt);
    if (opt_in) {
        LBC_REQUIRE(opt_in->struct_size == (int32_t)sizeof(lbc_plan_options), LBC_ERR_INVALID_ARG,
                    "lbc_plan_options.struct_size is %d, this library expects %d", opt_in->struct_size, (int)sizeof(lbc_plan_options));
        opt = *opt_in;
    }
    const bool snake = opt.reverse != 0;     // the network sets every layer's direction itself
    opt.reverse = -1;
    lbc_net* net = new (std::nothrow) lbc_net();
    LBC_REQUIRE(net, LBC_ERR_ALLOC, "out of host memory");
    net->sm_count = dev.sm_count;
    if (cudaMalloc((void**)&net->flag, sizeof(int)) != cudaSuccess || cudaMemset(net->flag, 0, sizeof(int)) != cudaSuccess) {
        cudaGetLastError();
        delete net;
        set_error("cannot allocate the network's device status word");
        return LBC_ERR_ALLOC;
    }
    net->layers.resize(n_layers);
    for (int i = 0; i < n_layers && st == LBC_OK; ++i) {
        lbc_net::Layer& L = net->layers[i];
        const lbc_node& nd = nodes[i];
        L.kind = nd.kind;
        L.input_of = nd.input_of;
        L.input2_of = nd.kind == LBC_NODE_ADD ? nd.input2_of : -1;
        if (L.input_of >= i || L.input2_of >= i || L.input_of < -1) {
            set_error("node %d: inputs (%d, %d) must reference earlier nodes", i, L.input_of, L.input2_of);
            st = LBC_ERR_INVALID_ARG;
            break;
        }
        // what the producer hands over: int8 NHWC of (n, p, q, k)
        auto produced = [&](int32_t src, int32_t* n, int32_t* h, int32_t* w, int32_t* c) {
            const lbc_net::Layer& P = net->layers[src];
            *n = P.in_n; *h = P.out_p; *w = P.out_q; *c = P.out_k;
            return P.out_elt == 1;
        };
        if (nd.kind == LBC_NODE_CONV) {
            st = plan_build(&nd.conv, LBC_KERNEL_AUTO, dev, &opt, false, net->flag, &L.plan);
            if (st != LBC_OK) break;
            const ConvGeom& g = L.plan->g;
            L.in_n = g.d.n; L.in_h = g.d.h; L.in_w = g.d.w; L.in_c = g.d.c;
            L.out_p = g.p; L.out_q = g.q; L.out_k = g.d.k; L.out_elt = (int32_t)out_elt(g);
        } else if (nd.kind == LBC_NODE_MAXPOOL) {
            L.pool = nd.pool;
            st = pool_geom(&L.pool, &L.out_p, &L.out_q);
            if (st != LBC_OK) break;
            L.in_n = L.pool.n; L.in_h = L.pool.h; L.in_w = L.pool.w; L.in_c = L.pool.c;
            L.out_k = L.pool.c;
        } else if (nd.kind == LBC_NODE_ADD) {
            if (L.input_of < 0 || L.input2_of < 0) {
                set_error("node %d: an add needs two producer nodes", i);
                st = LBC_ERR_INVALID_ARG;
                break;
            }
            int32_t n2, h2, w2, c2;
            produced(L.input_of, &L.in_n, &L.in_h, &L.in_w, &L.in_c);
            const bool ok2 = produced(L.input2_of, &n2, &h2, &w2, &c2);
            if (!ok2 || n2 != L.in_n || h2 != L.in_h || w2 != L.in_w || c2 != L.in_c) {
                set_error("node %d: the operands of the add differ (%dx%dx%dx%d vs %dx%dx%dx%d)", i, L.in_n, L.in_h, L.in_w, L.in_c, n2, h2, w2, c2);
                st = LBC_ERR_INVALID_ARG;
                break;
            }
            L.add_relu = nd.relu;
            L.out_p = L.in_h; L.out_q = L.in_w; L.out_k = L.in_c;
        } else {
            set_error("node %d: unknown kind %d", i, nd.kind);
            st = LBC_ERR_INVALID_ARG;
            break;
        }
        if (L.input_of >= 0) {
            int32_t pn, ph, pw, pc;
            const bool int8_src = produced(L.input_of, &pn, &ph, &pw, &pc);
            if (!int8_src || pn != L.in_n || ph != L.in_h || pw != L.in_w || pc != L.in_c) {
                set_error("layer %d: input %dx%dx%dx%d does not match the output of layer %d (%dx%dx%dx%d)", i, L.in_n, L.in_h,
                          L.in_w, L.in_c, L.input_of, pn, ph, pw, pc);
                st = LBC_ERR_INVALID_ARG;
                break;
            }
        }
        {
            const bool tc = L.kind == LBC_NODE_CONV &&
                            (L.plan->kind == LBC_KERNEL_IGEMM_TC || L.plan->kind == LBC_KERNEL_STEM_TC ||
                             L.plan->kind == LBC_KERNEL_DEPTHWISE);      // (the tiled depthwise kernel mirrors its tile index)
            const bool producer_rev = L.input_of >= 0 && net->layers[L.input_of].reverse;
            // a producer that ran forwards finished on the last images: start there; CUDA-core kernels always run forwards.
            // A small-C (stem) layer's producer is its own space-to-depth pre-pass, which always runs forwards.
            const bool own_prepass = L.kind == LBC_NODE_CONV && L.plan->kind == LBC_KERNEL_STEM_TC;
            L.reverse = tc && (L.input_of >= 0 ? !producer_rev : own_prepass) && snake;
        }
        bool ok = cudaMalloc(&L.y, L.out_bytes()) == cudaSuccess;
        size_t wb = 0;
        if (ok && L.kind == LBC_NODE_CONV) {
            wb = packed_weight_bytes(L.plan);
            ok = cudaMalloc(&L.w, wb) == cudaSuccess && cudaMalloc((void**)&L.bias, sizeof(int32_t) * L.out_k) == cudaSuccess &&
                 cudaMalloc((void**)&L.scale, sizeof(float) * L.out_k) == cudaSuccess;
        }
        if (ok && L.input_of < 0) ok = cudaMalloc(&L.x_own, L.in_bytes()) == cudaSuccess;
        if (!ok) {
            cudaGetLastError();
            set_error("layer %d: device allocation failed", i);
            st = LBC_ERR_ALLOC;
            break;
        }
        if (L.kind == LBC_NODE_CONV) {
            cudaMemset(L.w, 0, wb);
            cudaMemset(L.bias, 0, sizeof(int32_t) * L.out_k);
            cudaMemset(L.scale, 0, sizeof(float) * L.out_k);
        }
        if (L.x_own) cudaMemset(L.x_own, 0, L.in_bytes());
    }
    // Opt-in (lbc_plan_options::fuse = 1).  Measured on ResNet-50 stage 1 at N = 512 (r02): a fused pair takes 160-169 us against
    // 56 + 89 us for the two launches - it removes 2 x 103 MB of HBM traffic per bottleneck, but both epilogues now run on the
    // same 16 warps one after the other and the per-tile fixed costs of two drains add up; until the epilogue is split into
    // dedicated epi1 / epi2 warp groups the planner does not choose it on its own.
    if (st == LBC_OK && opt.fuse == 1) {
        // fused bottleneck tails: a convolution whose ONLY consumer is a 1x1 convolution, both shapes the fused kernel covers
        std::vector<int> consumers((size_t)n_layers, 0);
        for (int i = 0; i < n_layers; ++i) {
            if (net->layers[i].input_of >= 0) ++consumers[(size_t)net->layers[i].input_of];
            if (net->layers[i].input2_of >= 0) ++consumers[(size_t)net->layers[i].input2_of];
        }
        for (int j = 1; j < n_layers; ++j) {
            lbc_net::Layer& B = net->layers[j];
            const int i = B.input_of;
            if (B.kind != LBC_NODE_CONV || i < 0) continue;
            lbc_net::Layer& A = net->layers[i];
            if (A.kind != LBC_NODE_CONV || consumers[(size_t)i] != 1 || A.fused_from >= 0 || A.fused_into >= 0) continue;
            if (A.plan->kind != LBC_KERNEL_IGEMM_TC || B.plan->kind != LBC_KERNEL_IGEMM_TC || A.plan->pw_factor != 1 || B.plan->pw_factor != 1)
                continue;
            if (!fused_tail_supported(A.plan->g, A.plan->cfg, B.plan->g, B.plan->cfg, nullptr)) continue;
            A.fused_into = j;
            B.fused_from = i;
        }
        // traversal directions again, now that a fused pair is one launch that walks conv A's tiles
        for (int i = 0; i < n_layers; ++i) {
            lbc_net::Layer& L = net->layers[i];
            if (L.fused_from >= 0) { L.reverse = net->layers[L.fused_from].reverse; continue; }
            const bool tc = L.kind == LBC_NODE_CONV &&
                            (L.plan->kind == LBC_KERNEL_IGEMM_TC || L.plan->kind == LBC_KERNEL_STEM_TC || L.plan->kind == LBC_KERNEL_DEPTHWISE);
            const bool producer_rev = L.input_of >= 0 && net->layers[L.input_of].reverse;
            const bool own_prepass = L.kind == LBC_NODE_CONV && L.plan->kind == LBC_KERNEL_STEM_TC;
            L.reverse = tc && (L.input_of >= 0 ? !producer_rev : own_prepass) && snake;
        }
    }
    if (st == LBC_OK) {
        net->events.assign(n_layers + 1, nullptr);
        for (auto& e : net->events)
            if (cudaEventCreate(&e) != cudaSuccess) {
                e = nullptr;
                set_error("cudaEventCreate failed");
                st = LBC_ERR_CUDA;
                break;
            }
    }
    if (st != LBC_OK) {
        char keep[512];
        strncpy(keep, last_error(), sizeof keep);
        keep[sizeof keep - 1] = 0;
        lbc_net_destroy(net);
        set_error("%s", keep);
        return st;
    }
    *out = net;
    return LBC_OK;
}

lbc_status lbc_net_layer_plan(const lbc_net* net, int32_t layer, const lbc_plan** plan)
{
    LBC_REQUIRE(net && plan && layer >= 0 && layer < (int)net->layers.size(), LBC_ERR_INVALID_ARG, "bad layer index");
    LBC_REQUIRE(net->layers[layer].kind == LBC_NODE_CONV, LBC_ERR_INVALID_ARG, "node %d is not a convolution (kind %d)", layer,
                net->layers[layer].kind);
    *plan = net->layers[layer].plan;
    return LBC_OK;
}

lbc_status lbc_net_layer_fused_into(const lbc_net* net, int32_t layer, int32_t* into)
{
    LBC_REQUIRE(net && into && layer >= 0 && layer < (int)net->layers.size(), LBC_ERR_INVALID_ARG, "bad layer index");
    *into = net->layers[layer].fused_into;
    return LBC_OK;
}

lbc_status lbc_net_set_params_host(lbc_net* net, int32_t layer, const int8_t* w_host, int32_t layout,
                                   const int32_t* bias_host, const float* scale_host)
{
    LBC_REQUIRE(net && layer >= 0 && layer < (int)net->layers.size(), LBC_ERR_INVALID_ARG, "bad layer index");
    lbc_net::Layer& L = net->layers[layer];
    LBC_REQUIRE(L.kind == LBC_NODE_CONV, LBC_ERR_INVALID_ARG, "node %d is not a convolution: it has no parameters", layer);
    const ConvGeom& g = L.plan->g;
    if (w_host) {
        const size_t raw = (size_t)g.d.k * g.d.r * g.d.s * g.cg;
        void* tmp = nullptr;
        if (cudaMalloc(&tmp, raw) != cudaSuccess) {
            cudaGetLastError();
            set_error("set_params: device allocation failed");
            return LBC_ERR_ALLOC;
        }
        cudaError_t e = cudaMemcpy(tmp, w_host, raw, cudaMemcpyHostToDevice);
        lbc_status st = e == cudaSuccess ? lbc_conv_prepack_weights(L.plan, (const int8_t*)tmp, layout, L.w, nullptr) : LBC_ERR_CUDA;
        if (st == LBC_OK) e = cudaDeviceSynchronize();
        cudaFree(tmp);
        if (st != LBC_OK) return st;
        LBC_CUDA_TRY(e);
    }
    if (bias_host) LBC_CUDA_TRY(cudaMemcpy(L.bias, bias_host, sizeof(int32_t) * g.d.k, cudaMemcpyHostToDevice));
    if (scale_host) LBC_CUDA_TRY(cudaMemcpy(L.scale, scale_host, sizeof(float) * g.d.k, cudaMemcpyHostToDevice));
    return LBC_OK;
}

lbc_status lbc_net_set_input_host(lbc_net* net, int32_t layer, const int8_t* x_host)
{
    LBC_REQUIRE(net && x_host && layer >= 0 && layer < (int)net->layers.size(), LBC_ERR_INVALID_ARG, "bad argument");
    lbc_net::Layer& L = net->layers[layer];
    LBC_REQUIRE(L.x_own, LBC_ERR_INVALID_ARG, "layer %d takes its input from layer %d, not from a resident buffer", layer, L.input_of);
    LBC_CUDA_TRY(cudaMemcpy(L.x_own, x_host, L.in_bytes(), cudaMemcpyHostToDevice));
    return LBC_OK;
}

lbc_status lbc_net_read_output_host(const lbc_net* net, int32_t layer, void* y_host, size_t max_bytes)
{
    LBC_REQUIRE(net && y_host && layer >= 0 && layer < (int)net->layers.size(), LBC_ERR_INVALID_ARG, "bad argument");
    const lbc_net::Layer& L = net->layers[layer];
    LBC_REQUIRE(L.fused_into < 0, LBC_ERR_UNSUPPORTED, "layer %d runs fused into layer %d: its output is never materialised", layer,
                L.fused_into);
    const size_t all = L.out_bytes();
    LBC_CUDA_TRY(cudaDeviceSynchronize());
    LBC_CUDA_TRY(cudaMemcpy(y_host, L.y, (max_bytes && max_bytes < all) ? max_bytes : all, cudaMemcpyDeviceToHost));
    return LBC_OK;
}

lbc_status lbc_net_layer_io(const lbc_net* net, int32_t layer, const void** x_dev, void** y_dev)
{
    LBC_REQUIRE(net && layer >= 0 && layer < (int)net->layers.size(), LBC_ERR_INVALID_ARG, "bad layer index");
    if (x_dev) *x_dev = layer_input(net, layer);
    if (y_dev) *y_dev = net->layers[layer].y;
    return LBC_OK;
}

lbc_status lbc_net_run(lbc_net* net, const int8_t* x_dev, lbc_stream stream, float* per_layer_ms, float* total_ms)
{
    LBC_REQUIRE(net, LBC_ERR_INVALID_ARG, "null net");
    std::lock_guard<std::mutex> lk(net->mu);
    cudaStream_t s = (cudaStream_t)stream;
    const bool timed = per_layer_ms || total_ms;
    lbc_status st = net_enqueue(net, x_dev, s, timed);
    if (st != LBC_OK) return st;
    if (!timed) return LBC_OK;
    const int n = (int)net->layers.size();
    LBC_CUDA_TRY(cudaEventSynchronize(net->events[n]));
    if (per_layer_ms)
        for (int i = 0; i < n; ++i) LBC_CUDA_TRY(cudaEventElapsedTime(&per_layer_ms[i], net->events[i], net->events[i + 1]));
    if (total_ms) LBC_CUDA_TRY(cudaEventElapsedTime(total_ms, net->events[0], net->events[n]));
    return check_flag(net->flag);
}

lbc_status lbc_net_run_host(lbc_net* net, const int8_t* x_host, void* y_host, lbc_stream stream, float* total_ms)
{
    LBC_REQUIRE(net && x_host && y_host, LBC_ERR_INVALID_ARG, "null argument");
    std::lock_guard<std::mutex> lk(net->mu);
    cudaStream_t s = (cudaStream_t)stream;
    const int n = (int)net->layers.size();
    lbc_net::Layer& first = net->layers[0];
    lbc_net::Layer& last = net->layers[n - 1];
    LBC_REQUIRE(first.x_own, LBC_ERR_INVALID_ARG, "layer 0 must take the network input");
    LBC_CUDA_TRY(cudaEventRecord(net->events[0], s));
    LBC_CUDA_TRY(cudaMemcpyAsync(first.x_own, x_host, first.in_bytes(), cudaMemcpyHostToDevice, s));
    lbc_status st = net_enqueue(net, nullptr, s, false);
    if (st != LBC_OK) return st;
    LBC_CUDA_TRY(cudaMemcpyAsync(y_host, last.y, last.out_bytes(), cudaMemcpyDeviceToHost, s));
    LBC_CUDA_TRY(cudaEventRecord(net->events[n], s));
    LBC_CUDA_TRY(cudaEventSynchronize(net->events[n]));
    if (total_ms) LBC_CUDA_TRY(cudaEventElapsedTime(total_ms, net->events[0], net->events[n]));
    return check_flag(net->flag);
}

static lbc_status pipe_init(lbc_net* net)
{
    lbc_net::Pipe& pp = net->pipe;
    if (pp.ready) return LBC_OK;
    lbc_net::Layer& first = net->layers[0];
    LBC_REQUIRE(first.x_own, LBC_ERR_INVALID_ARG, "layer 0 must take the network input");
    // every handle is created only if it is still null, so a call that failed half-way can simply be repeated
    // (lbc_net_destroy releases whatever exists)
    if (!pp.h2d) LBC_CUDA_TRY(cudaStreamCreateWithFlags(&pp.h2d, cudaStreamNonBlocking));
    if (!pp.d2h) LBC_CUDA_TRY(cudaStreamCreateWithFlags(&pp.d2h, cudaStreamNonBlocking));
    pp.x_buf[0] = first.x_own;
    if (!pp.x_buf[1] && cudaMalloc(&pp.x_buf[1], first.in_bytes()) != cudaSuccess) {
        cudaGetLastError();
        pp.x_buf[1] = nullptr;
        set_error("lbc_net_submit_host: cannot allocate the second input buffer");
        return LBC_ERR_ALLOC;
    }
    for (cudaEvent_t* e : {&pp.up_done[0], &pp.up_done[1], &pp.x_used[0], &pp.x_used[1], &pp.comp_done, &pp.d2h_done})
        if (!*e) LBC_CUDA_TRY(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
    if (!pp.t_first) LBC_CUDA_TRY(cudaEventCreate(&pp.t_first));
    if (!pp.t_last) LBC_CUDA_TRY(cudaEventCreate(&pp.t_last));
    pp.ready = true;
    return LBC_OK;
}

lbc_status lbc_net_submit_host(lbc_net* net, const int8_t* x_host, void* y_host, lbc_stream stream)
{
    LBC_REQUIRE(net && x_host && y_host, LBC_ERR_INVALID_ARG, "null argument");
    std::lock_guard<std::mutex> lk(net->mu);
    lbc_status st = pipe_init(net);
    if (st != LBC_OK) return st;
    lbc_net::Pipe& pp = net->pipe;
    cudaStream_t s = (cudaStream_t)stream;
    const int n = (int)net->layers.size();
    const int b = (int)(pp.submitted & 1);
    lbc_net::Layer& first = net->layers[0];
    lbc_net::Layer& last = net->layers[n - 1];
    // upload: the buffer is free once layer 0 of the step two submissions ago has run
    if (pp.submitted == 0) LBC_CUDA_TRY(cudaEventRecord(pp.t_first, pp.h2d));
    if (pp.submitted >= 2) LBC_CUDA_TRY(cudaStreamWaitEvent(pp.h2d, pp.x_used[b], 0));
    LBC_CUDA_TRY(cudaMemcpyAsync(pp.x_buf[b], x_host, first.in_bytes(), cudaMemcpyHostToDevice, pp.h2d));
    LBC_CUDA_TRY(cudaEventRecord(pp.up_done[b], pp.h2d));
    // compute: layer 0 waits for its upload; the last layer waits until the previous result has been downloaded
    LBC_CUDA_TRY(cudaStreamWaitEvent(s, pp.up_done[b], 0));
    for (int i = 0; i < n; ++i) {
        const ResolvedLaunch* rl = nullptr;
        if (i == 0 && net->layers[0].kind == LBC_NODE_CONV) {
            // layer 0 alternates between the two input buffers: keep one encoded launch per buffer
            if (!net->first_rl_ok[b]) {
                lbc_net::Layer& L = net->layers[0];
                st = resolve(L.plan, (const int8_t*)pp.x_buf[b], L.w, L.bias, L.scale, L.y, &net->first_rl[b]);
                if (st != LBC_OK) return st;
                net->first_rl[b].ig.reverse = L.reverse ? 1 : 0;
                net->first_rl[b].ig.early_b = L.plan->opt.early_weights == 2 ? 2 : L.plan->opt.early_weights != 0 ? 1 : 0;
                net->first_rl[b].dw.reverse = (L.reverse && net->first_rl[b].dw.tiled) ? 1 : 0;
                net->first_rl_ok[b] = true;
            }
            rl = &net->first_rl[b];
        } else {
            st = net_resolve(net, i, nullptr);
            if (st != LBC_OK) return st;
        }
        if (i == n - 1 && pp.submitted >= 1) LBC_CUDA_TRY(cudaStreamWaitEvent(s, pp.d2h_done, 0));
        st = net_launch_node(net, i, rl, i == 0 ? pp.x_buf[b] : nullptr, s);
        if (st != LBC_OK) return st;
        if (i == 0) LBC_CUDA_TRY(cudaEventRecord(pp.x_used[b], s));
    }
    LBC_CUDA_TRY(cudaEventRecord(pp.comp_done, s));
    // download
    LBC_CUDA_TRY(cudaStreamWaitEvent(pp.d2h, pp.comp_done, 0));
    LBC_CUDA_TRY(cudaMemcpyAsync(y_host, last.y, last.out_bytes(), cudaMemcpyDeviceToHost, pp.d2h));
    LBC_CUDA_TRY(cudaEventRecord(pp.d2h_done, pp.d2h));
    ++pp.submitted;
    return LBC_OK;
}

lbc_status lbc_net_sync_host(lbc_net* net, float* elapsed_ms)
{
    LBC_REQUIRE(net, LBC_ERR_INVALID_ARG, "null net");
    std::lock_guard<std::mutex> lk(net->mu);
    lbc_net::Pipe& pp = net->pipe;
    if (elapsed_ms) *elapsed_ms = 0.f;
    if (!pp.ready || pp.submitted == 0) return LBC_OK;
    LBC_CUDA_TRY(cudaEventRecord(pp.t_last, pp.d2h));
    LBC_CUDA_TRY(cudaEventSynchronize(pp.t_last));
    if (elapsed_ms) LBC_CUDA_TRY(cudaEventElapsedTime(elapsed_ms, pp.t_first, pp.t_last));
    pp.submitted = 0;
    return check_flag(net->flag);
}

lbc_status lbc_net_check(lbc_net* net)
{
    LBC_REQUIRE(net, LBC_ERR_INVALID_ARG, "null net");
    return check_flag(net->flag);
}

lbc_status lbc_net_launches(const lbc_net* net, int32_t* launches)
{
    LBC_REQUIRE(net && launches, LBC_ERR_INVALID_ARG, "null argument");
    int32_t n = 0;
    for (const auto& L : net->layers) n += L.fused_into >= 0 ? 0 : (L.kind == LBC_NODE_CONV && L.plan->kind == LBC_KERNEL_STEM_TC) ? 2 : 1;
    *launches = n;
    return LBC_OK;
}

}  // extern "C"
