/*
 * cpu_ref.c — CPU ORACLE for the int8 convolution hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * link or call this file.  The product path (lowbitdnn-project_b200/csrc) never does.
 *
 * What it restates (paths relative to the reference root, /root/reference in the authoring box):
 *   - loop nest, cross-correlation orientation, int32 accumulation:
 *         cpp/int8conv/refConv2DForward.hpp:24-51   (refConv2DForwardImpl)
 *   - zero-padding predicate:  cpp/int8conv/conv2DForward3x3.cuh:666-667
 *   - output-size formula:     cpp/int8conv/cudnn2DConvolution.cuh:33-36
 *   - VECT_C layout:           cpp/int8conv/utils.cuh:11-26
 *   - requantisation rule (round-to-nearest-even, then saturate to [-128,127]):
 *         cpp/int8conv/conv2DForward3x3WinogradFused.cuh:39-46
 *         python/qtorch/nn/functional/quantization.py:27-49
 *
 * Pinning: oracle_ref_conv_nchw_valid() is checked bit-for-bit against the reference's own
 * refConv2DForward (compiled unmodified into oracle/_ref/libref_conv.so by oracle/Makefile) in
 * tests/test_oracle.py, and against the golden vectors that library produced (tests/golden/).
 * The reference holds no golden vectors of its own.  The bias / per-channel scale /
 * ReLU epilogue, NHWC, stride, dilation and groups are north_star-defined extensions: they are
 * pinned only through oracle_conv_nhwc()'s agreement with the reference on the shapes both express.
 *
 * Build: gcc -O3 -march=x86-64-v3 -fopenmp -ffp-contract=off -shared -fPIC cpu_ref.c -o _build/libcpu_ref.so
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct oracle_conv_desc {
    int32_t n, h, w, c;
    int32_t k, r, s;
    int32_t stride_h, stride_w;
    int32_t pad_h, pad_w;
    int32_t dil_h, dil_w;
    int32_t groups;
    int32_t relu;
    int32_t out_mode; /* 0 = int8 requantised, 1 = int32 accumulators (+bias) */
} oracle_conv_desc;

int oracle_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* cudnn2DConvolution.cuh:33-36 */
int32_t oracle_out_dim(int32_t in, int32_t pad, int32_t dil, int32_t k, int32_t stride)
{
    return (in + 2 * pad - (dil * (k - 1) + 1)) / stride + 1;
}

/* Requantise one accumulator: the reference's rule quantize() =
 *   max(-128, min(__float2int_rn(v * scale), 127))          (conv2DForward3x3WinogradFused.cuh:39-46)
 * i.e. ROUND (half-to-even, saturating to int32, NaN -> 0 as cvt.rni.s32.f32 defines it) and THEN clamp; ReLU raises
 * the lower clamp to 0.  One fp32 multiply, no FMA (this TU is built with -ffp-contract=off). */
int8_t oracle_requant(int32_t acc, int32_t bias, float scale, int relu)
{
    int32_t t = (int32_t)((uint32_t)acc + (uint32_t)bias); /* int32 wraparound */
    volatile float f = (float)t;                           /* cvt.rn.f32.s32   */
    f = f * scale;
    float g = f;
    long q;
    if (isnan(g)) q = 0;                         /* __float2int_rn(NaN) == 0 */
    else if (g >= 2147483648.0f) q = 2147483647L;  /* saturating conversion */
    else if (g <= -2147483648.0f) q = -2147483647L - 1;
    else q = lrintf(g);                          /* default rounding mode: half-to-even */
    const long lo = relu ? 0 : -128;
    if (q < lo) q = lo;
    if (q > 127) q = 127;
    return (int8_t)q;
}

/*
 * Exact restatement of refConv2DForwardImpl (refConv2DForward.hpp:24-51): NCHW int8 input that the
 * caller has PRE-PADDED, OIHW int8 kernel, VALID, stride 1, dilation 1, groups 1, int32 NCHW output.
 */
void oracle_ref_conv_nchw_valid(int32_t batch, int32_t inC, int32_t inH, int32_t inW,
                                int32_t outC, int32_t outH, int32_t outW, int32_t kH, int32_t kW,
                                const int8_t* input, const int8_t* kernel, int32_t* output)
{
    for (int32_t b = 0; b < batch; ++b) {
#pragma omp parallel for
        for (int32_t o = 0; o < outC; ++o)
            for (int32_t y = 0; y < outH; ++y)
                for (int32_t x = 0; x < outW; ++x) {
                    int32_t result = 0;
                    for (int32_t i = 0; i < inC; ++i)
                        for (int32_t ky = 0; ky < kH; ++ky)
                            for (int32_t kx = 0; kx < kW; ++kx)
                                result += (int32_t)kernel[((o * inC + i) * kH + ky) * kW + kx] *
                                          (int32_t)input[((b * inC + i) * inH + (y + ky)) * inW + (x + kx)];
                    output[((b * outC + o) * outH + y) * outW + x] = result;
                }
    }
}

/*
 * General oracle: NHWC int8 activations, [K][R][S][C/groups] int8 weights, int32 bias (may be NULL),
 * fp32 per-output-channel scale, optional ReLU.  Writes int8 NHWC (out_mode 0) or int32 NHWC (1).
 * Returns 0 on success.
 */
int oracle_conv_nhwc(const oracle_conv_desc* d, const int8_t* x, const int8_t* w,
                     const int32_t* bias, const float* scale, void* y, int threads)
{
    if (!d || !x || !w || !y) return 1;
    if (d->groups <= 0 || d->c % d->groups || d->k % d->groups) return 1;
    if (d->out_mode == 0 && !scale) return 1;
    const int32_t P = oracle_out_dim(d->h, d->pad_h, d->dil_h, d->r, d->stride_h);
    const int32_t Q = oracle_out_dim(d->w, d->pad_w, d->dil_w, d->s, d->stride_w);
    if (P <= 0 || Q <= 0) return 1;
    const int32_t cg = d->c / d->groups, kg = d->k / d->groups;
    const int64_t rows = (int64_t)d->n * P;
#ifdef _OPENMP
    if (threads <= 0) threads = omp_get_max_threads();
#else
    (void)threads;
#endif
#pragma omp parallel for num_threads(threads) schedule(static)
    for (int64_t row = 0; row < rows; ++row) {
        const int32_t n = (int32_t)(row / P), p = (int32_t)(row % P);
        for (int32_t q = 0; q < Q; ++q) {
            for (int32_t k = 0; k < d->k; ++k) {
                const int32_t g = k / kg;
                int32_t acc = 0;
                for (int32_t r = 0; r < d->r; ++r) {
                    const int32_t ih = p * d->stride_h - d->pad_h + r * d->dil_h;
                    if (ih < 0 || ih >= d->h) continue; /* zero padding: conv2DForward3x3.cuh:666-667 */
                    for (int32_t s = 0; s < d->s; ++s) {
                        const int32_t iw = q * d->stride_w - d->pad_w + s * d->dil_w;
                        if (iw < 0 || iw >= d->w) continue;
                        const int8_t* xp = x + (((int64_t)n * d->h + ih) * d->w + iw) * d->c + (int64_t)g * cg;
                        const int8_t* wp = w + (((int64_t)k * d->r + r) * d->s + s) * cg;
                        int32_t part = 0;
                        for (int32_t c = 0; c < cg; ++c) part += (int32_t)xp[c] * (int32_t)wp[c];
                        acc += part;
                    }
                }
                const int64_t o = (((int64_t)n * P + p) * Q + q) * d->k + k;
                const int32_t b = bias ? bias[k] : 0;
                if (d->out_mode == 1)
                    ((int32_t*)y)[o] = (int32_t)((uint32_t)acc + (uint32_t)b);
                else
                    ((int8_t*)y)[o] = oracle_requant(acc, b, scale[k], d->relu);
            }
        }
    }
    return 0;
}

/* utils.cuh:20-26 — [N,C,H,W] -> [N,C/V,H,W,V] (materialised). elt = element size in bytes. */
void oracle_to_vect_c(const void* src, void* dst, int32_t n, int32_t c, int32_t h, int32_t w, int32_t v, int32_t elt)
{
    const char* s = (const char*)src;
    char* t = (char*)dst;
    for (int32_t in = 0; in < n; ++in)
        for (int32_t ic = 0; ic < c; ++ic)
            for (int32_t ih = 0; ih < h; ++ih)
                for (int32_t iw = 0; iw < w; ++iw) {
                    int64_t si = (((int64_t)in * c + ic) * h + ih) * w + iw;
                    int64_t di = (((((int64_t)in * (c / v) + ic / v) * h + ih) * w + iw) * v) + ic % v;
                    memcpy(t + di * elt, s + si * elt, (size_t)elt);
                }
}

/* utils.cuh:11-17 — inverse. */
void oracle_from_vect_c(const void* src, void* dst, int32_t n, int32_t c, int32_t h, int32_t w, int32_t v, int32_t elt)
{
    const char* s = (const char*)src;
    char* t = (char*)dst;
    for (int32_t in = 0; in < n; ++in)
        for (int32_t ic = 0; ic < c; ++ic)
            for (int32_t ih = 0; ih < h; ++ih)
                for (int32_t iw = 0; iw < w; ++iw) {
                    int64_t di = (((int64_t)in * c + ic) * h + ih) * w + iw;
                    int64_t si = (((((int64_t)in * (c / v) + ic / v) * h + ih) * w + iw) * v) + ic % v;
                    memcpy(t + di * elt, s + si * elt, (size_t)elt);
                }
}
