/*
 * vsiq_oracle.c -- CPU restatement of the reference's QAT fake-quantization
 * hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is the parity checker for the CUDA kernels under
 * vsiquantization_b200/csrc/.  Nothing in the product package may import,
 * link or call it; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs do (and there only as the checker or
 * as the timed CPU baseline).
 *
 * Parity pin: the reference (tranngocduvnvp/VSIQuantization) has no tests and
 * no golden vectors.  The pin is the reference's own Python code executed on
 * CPU in the build container (oracle/gen_golden.py imports it from
 * /root/reference and writes tests/golden/*.npz); tests/test_oracle_golden.py
 * checks every function below against those vectors.  The arithmetic the
 * reference relies on lives in PyTorch ATen (un-vendored third-party
 * dependency, no version pinned by the reference; vectors generated with
 * torch 2.11.0+cu128 on CPU).
 *
 * Every fp32 operation below is individually rounded (compile with
 * -ffp-contract=off, no -ffast-math): the reference composes separate ATen
 * elementwise kernels, so there is never an FMA contraction.
 *
 * Tensor layout convention used throughout: a contiguous fp32 tensor viewed
 * as [outer, C, inner]; channel(c) of flat index i is (i / inner) % C.
 *   per-tensor            : outer = 1, C = 1, inner = numel
 *   per-channel, ch_axis 0: outer = 1, C = shape[0], inner = numel / C   (OIHW weights)
 *   per-channel, ch_axis 1: outer = N, C = shape[1], inner = H*W         (NCHW activations)
 */
#include <math.h>
#include <stdint.h>
#include <stddef.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define VSIQ_ORACLE_VERSION 1

int vsiq_oracle_version(void) { return VSIQ_ORACLE_VERSION; }

void vsiq_oracle_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int vsiq_oracle_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* torch.clamp(x, lo, hi) on CPU: min(max(x, lo), hi) with NaN propagated and
 * -0.0 preserved when lo == 0 (reference: quantizers/uniform.py:95). */
static inline float clamp_torch(float r, float lo, float hi) {
    if (r != r) return r;
    if (r < lo) return lo;
    if (r > hi) return hi;
    return r;
}

/* Effective zero-point used by the forward pass.
 * learned == 0: zero_point is a constant (Python int in the reference,
 *               quantizers/quantization_manager.py:46,103).
 * learned == 1: zero_point is the float parameter z_f and the forward uses
 *               clamp(rint(z_f), qmin, qmax)  (quantizers/uniform.py:98-102,
 *               quantizers/lsq_module.py:354-358). */
static inline float effective_zp(float zf, int learned, float qmin, float qmax) {
    if (!learned) return zf;
    return clamp_torch(rintf(zf), qmin, qmax);
}

/* ------------------------------------------------------------------------
 * Fake-quant forward.
 * Reference: UniformQuantizer.quantize / discreate_tensor,
 *   quantizers/uniform.py:54-55 and :95
 *     x_int = clamp(round(x/scale + zero_point), qmin, qmax)
 *     y     = (x_int - zero_point) * scale
 * and the per-channel form LSQFakeQuantize.fake_quantize_per_channel_affine,
 *   quantizers/lsq_module.py:254-274 (qparams broadcast as [1,C,1,...]).
 * scale/zero_point arrive already rounded to fp32 (the reference's fp64
 * 0-dim Parameter / Python float never promotes the fp32 tensor).
 * codes (optional) receives x_int as fp32.
 * ---------------------------------------------------------------------- */
void vsiq_oracle_fake_quant_fwd(const float *x, float *y, float *codes,
                                int64_t outer, int64_t C, int64_t inner,
                                const float *scale, const float *zero_point,
                                int zp_learned, int qmin, int qmax) {
    const float lo = (float)qmin, hi = (float)qmax;
    const int64_t rows = outer * C;
#pragma omp parallel for schedule(static)
    for (int64_t row = 0; row < rows; ++row) {
        const int64_t c = row % C;
        const float s = scale[c];
        const float z = effective_zp(zero_point[c], zp_learned, lo, hi);
        const float *xr = x + row * inner;
        float *yr = y + row * inner;
        float *cr = codes ? codes + row * inner : NULL;
        for (int64_t i = 0; i < inner; ++i) {
            float v = xr[i] / s;          /* IEEE division, not x * (1/s) */
            float t = v + z;              /* always performed, even for z == 0 */
            float r = rintf(t);           /* round half to even */
            float q = clamp_torch(r, lo, hi);
            float d = q - z;
            yr[i] = d * s;
            if (cr) cr[i] = q;
        }
    }
}

/* ------------------------------------------------------------------------
 * Backward of the above as PyTorch autograd computes it for the reference
 * graph (quantizers/uniform.py:47-55, :242-271; per-channel:
 * quantizers/lsq_module.py:147-173, :317-340).
 *
 *   grad_d = g * s                                   (mul backward)
 *   grad_v = (qmin <= r <= qmax) ? grad_d : 0        (clamp backward, mask on the ROUNDED value, inclusive)
 *   dx     = grad_v / s                              (div backward)
 *   ds_c   = gs * [ sum g*(q - z)  +  sum -grad_v * ((x/s)/s) ]
 *   dz_c   = gs * c_z * [ sum grad_v - sum grad_d ]  (only when zero_point is learned;
 *            c_z = [qmin <= rint(z_f) <= qmax])
 *
 * dx is reproduced operation-for-operation in fp32 (bit-comparable).
 * The sums are accumulated here in double (the "truth"); the reference
 * accumulates in fp32 with ATen's own tree order, so tests compare ds/dz to
 * the golden reference values with a tolerance.
 * mask_mode 1 selects the reference's dead-code FunLSQ masking
 * (quantizers/uniform.py:144-150): masks on the UNROUNDED x/s, strict
 * inequalities, symmetric per-tensor only (z ignored).
 * ds / dz may be NULL (plain STE backward).
 * ---------------------------------------------------------------------- */
void vsiq_oracle_fake_quant_bwd(const float *x, const float *g, float *dx,
                                double *ds, double *dz,
                                int64_t outer, int64_t C, int64_t inner,
                                const float *scale, const float *zero_point,
                                int zp_learned, int qmin, int qmax,
                                const double *grad_scale, int mask_mode) {
    const float lo = (float)qmin, hi = (float)qmax;
    for (int64_t c = 0; c < C; ++c) {
        if (ds) ds[c] = 0.0;
        if (dz) dz[c] = 0.0;
    }
    for (int64_t c = 0; c < C; ++c) {
        const float s = scale[c];
        const float zf = zero_point[c];
        const float z = effective_zp(zf, zp_learned, lo, hi);
        double acc_mul = 0.0, acc_div = 0.0, acc_gv = 0.0, acc_gd = 0.0;
#pragma omp parallel for schedule(static) reduction(+:acc_mul,acc_div,acc_gv,acc_gd)
        for (int64_t o = 0; o < outer; ++o) {
            const int64_t base = (o * C + c) * inner;
            for (int64_t i = 0; i < inner; ++i) {
                const float xv = x[base + i], gv = g[base + i];
                if (mask_mode == 1) {
                    /* FunLSQ: quantizers/uniform.py:144-150 */
                    float v = xv / s;
                    float small = (v < lo) ? 1.0f : 0.0f;
                    float big = (v > hi) ? 1.0f : 0.0f;
                    float mid = 1.0f - small - big;
                    float term = small * lo + big * hi + mid * (-v + rintf(v));
                    acc_mul += (double)(term * gv);
                    dx[base + i] = mid * gv;
                    continue;
                }
                float v = xv / s;
                float t = v + z;
                float r = rintf(t);
                float q = clamp_torch(r, lo, hi);
                float d = q - z;
                int m = (r >= lo) && (r <= hi);
                float grad_d = gv * s;
                float grad_v = m ? grad_d : 0.0f;
                dx[base + i] = grad_v / s;
                acc_mul += (double)(gv * d);
                acc_div += (double)(-grad_v * (v / s));
                acc_gv += (double)grad_v;
                acc_gd += (double)grad_d;
            }
        }
        const double gs = grad_scale ? grad_scale[c] : 1.0;
        if (ds) ds[c] = gs * (acc_mul + acc_div);
        if (dz) {
            float zr = rintf(zf);
            int cz = zp_learned ? ((zr >= lo) && (zr <= hi)) : 1;
            dz[c] = cz ? gs * (acc_gv - acc_gd) : 0.0;
        }
    }
}

/* ------------------------------------------------------------------------
 * MinMax observation of ONE call + the three statistics the manager collects
 * on the same tensor.
 * Reference: MinMaxObserver.observe, observers/minmax.py:42-43
 *   (x.min().item(), x.max().item(): NaN anywhere => NaN result), and
 *   QuantizationManager.collect_qparameter,
 *   quantizers/quantization_manager.py:66-68 (mean|x|, mean x, unbiased std).
 * Per channel (C > 1) this is the [outer, C, inner] reduction over outer and
 * inner that torch's PerChannelMinMaxObserver performs for
 * quantizers/lsq_module.py:115-123.
 * out[c] = {min, max, sum|x|, sum x, sum x^2} in double (min/max are exact
 * fp32 values widened).
 * ---------------------------------------------------------------------- */
void vsiq_oracle_minmax_stats(const float *x, int64_t outer, int64_t C, int64_t inner,
                              double *out /* [C][5] */) {
    for (int64_t c = 0; c < C; ++c) {
        float mn = INFINITY, mx = -INFINITY;
        int has_nan = 0;
        double sa = 0.0, s1 = 0.0, s2 = 0.0;
        for (int64_t o = 0; o < outer; ++o) {
            const float *p = x + (o * C + c) * inner;
            for (int64_t i = 0; i < inner; ++i) {
                float v = p[i];
                if (v != v) has_nan = 1;
                if (v < mn) mn = v;
                if (v > mx) mx = v;
                sa += fabs((double)v);
                s1 += (double)v;
                s2 += (double)v * (double)v;
            }
        }
        out[c * 5 + 0] = has_nan ? (double)NAN : (double)mn;
        out[c * 5 + 1] = has_nan ? (double)NAN : (double)mx;
        out[c * 5 + 2] = sa;
        out[c * 5 + 3] = s1;
        out[c * 5 + 4] = s2;
    }
}

/* Running-state update of the observer, observers/minmax.py:44-47:
 *   if min_x < self.min_val: self.min_val = min_x     (NaN never updates)
 *   if max_x > self.max_val: self.max_val = max_x
 * The running state starts at 0 (observers/minmax.py:28-29), never None. */
void vsiq_oracle_minmax_update(double *run_min, double *run_max, double call_min, double call_max) {
    if (call_min < *run_min) *run_min = call_min;
    if (call_max > *run_max) *run_max = call_max;
}

/* ------------------------------------------------------------------------
 * Scale / zero-point from min/max, in double exactly as the Python floats of
 * MinMaxObserver.get_scale_zero_point, observers/minmax.py:67-74:
 *   symmetric : scale = max(|min|,|max|) / (2^(b-1) - 1 + eps) ; zp = 0
 *   asymmetric: scale = (max - min) / (2^b - 1 + eps)
 *               zp    = round(-min / (scale + eps))   (Python round = half-even; NOT clamped)
 * zp is returned as a double holding an integer value.
 * ---------------------------------------------------------------------- */
void vsiq_oracle_qparams(double mn, double mx, int bits, int symmetric, double eps,
                         double *scale, double *zp) {
    if (symmetric) {
        double a = fabs(mn), b = fabs(mx);
        /* Python max(a, b): returns b only if b > a */
        double max_abs = (b > a) ? b : a;
        double levels = (double)((1 << (bits - 1)) - 1) + eps;
        *scale = max_abs / levels;
        *zp = 0.0;
    } else {
        double levels = (double)((1 << bits) - 1) + eps;
        double s = (mx - mn) / levels;
        *scale = s;
        *zp = rint(-mn / (s + eps));
    }
}

/* LSQ step-size initialisation,
 * QuantizationManager.init_scaling_factor_for_learning,
 * quantizers/quantization_manager.py:112:
 *   scale = 2 * mean(mean_abs_x) / sqrt(2^(bits-1) - 1)
 * mean_abs[k] is the k-th calibration call's mean|x|. */
double vsiq_oracle_lsq_init_scale(const double *mean_abs, int64_t n_calls, int bits) {
    double acc = 0.0;
    for (int64_t k = 0; k < n_calls; ++k) acc += mean_abs[k];
    double m = acc / (double)n_calls;
    return 2.0 * m / sqrt((double)((1 << (bits - 1)) - 1));
}

/* Gradient scale of the learnable path.
 * Per tensor : UniformQuantizer.calculate_grad_scale, quantizers/uniform.py:69-71
 *              gs = (qmax * numel) ** -0.5            [* calib_grad_scale, :48]
 * Per channel: LSQFakeQuantize.calculate_grad_scale, quantizers/lsq_module.py:327-340
 *              gs = (quant_max * numel / C) ** -0.5   [* 5000 for config_act, :151-152] */
double vsiq_oracle_grad_scale(int qmax, int64_t numel, int64_t C_or_1) {
    double n = (double)numel / (double)C_or_1;
    return pow((double)qmax * n, -0.5);
}

/* ------------------------------------------------------------------------
 * Conv/Linear + BatchNorm fold.
 * Reference: ConvBnReLU.__init__, modules/fused.py:98-108 and
 *            LinearBnReLU.__init__, modules/fused.py:292-300:
 *   std = sqrt(running_var + eps)
 *   t   = gamma / std
 *   W'  = W * t[c]                    (broadcast over [C,1,1,1] / [C,1])
 *   b'  = beta + (b - running_mean) * t     (b = 0 when the layer has no bias)
 * All fp32, that operation order.  bias may be NULL.
 * ---------------------------------------------------------------------- */
void vsiq_oracle_bn_fold(const float *W, const float *bias,
                         const float *gamma, const float *beta,
                         const float *mean, const float *var, float eps,
                         int64_t C, int64_t inner, float *W_out, float *b_out) {
#pragma omp parallel for schedule(static)
    for (int64_t c = 0; c < C; ++c) {
        float sd = sqrtf(var[c] + eps);
        float t = gamma[c] / sd;
        for (int64_t i = 0; i < inner; ++i) W_out[c * inner + i] = W[c * inner + i] * t;
        if (b_out) {
            float b = bias ? bias[c] : 0.0f;
            float d = b - mean[c];
            float e = d * t;
            b_out[c] = beta[c] + e;
        }
    }
}

/* ------------------------------------------------------------------------
 * BN statistics re-estimation, utils/estimate_bn.py:56-99.
 * With momentum = 1 and bn.training = True each forward leaves
 *   running_mean = batch mean over (N,H,W)
 *   running_var  = UNBIASED batch variance over (N,H,W)
 * in the BN layer (:60-65, :82); the function sums them over batches (:86-87)
 * and divides by the batch count (:96-97).
 * moments(): one batch -> per-channel {mean, unbiased var, biased var}.
 * ---------------------------------------------------------------------- */
void vsiq_oracle_bn_moments(const float *x, int64_t N, int64_t C, int64_t HW,
                            double *mean_out, double *var_unbiased_out, double *var_biased_out) {
#pragma omp parallel for schedule(static)
    for (int64_t c = 0; c < C; ++c) {
        double s1 = 0.0;
        for (int64_t n = 0; n < N; ++n) {
            const float *p = x + (n * C + c) * HW;
            for (int64_t i = 0; i < HW; ++i) s1 += (double)p[i];
        }
        const double cnt = (double)(N * HW);
        const double m = s1 / cnt;
        double s2 = 0.0;
        for (int64_t n = 0; n < N; ++n) {
            const float *p = x + (n * C + c) * HW;
            for (int64_t i = 0; i < HW; ++i) {
                double d = (double)p[i] - m;
                s2 += d * d;
            }
        }
        mean_out[c] = m;
        if (var_unbiased_out) var_unbiased_out[c] = s2 / (cnt - 1.0);
        if (var_biased_out) var_biased_out[c] = s2 / cnt;
    }
}

/* Accumulate-and-finalise step of reestimate_BN_stats
 * (utils/estimate_bn.py:86-87 and :96-97), fp32 like the reference buffers. */
void vsiq_oracle_bn_reestimate_accumulate(float *mean_sum, float *var_sum,
                                          const float *batch_mean, const float *batch_var_unbiased,
                                          int64_t C) {
    for (int64_t c = 0; c < C; ++c) {
        mean_sum[c] = mean_sum[c] + batch_mean[c];
        var_sum[c] = var_sum[c] + batch_var_unbiased[c];
    }
}

void vsiq_oracle_bn_reestimate_finish(const float *mean_sum, const float *var_sum,
                                      int64_t batch_count, float *running_mean, float *running_var,
                                      int64_t C) {
    const float k = (float)batch_count;
    for (int64_t c = 0; c < C; ++c) {
        running_mean[c] = mean_sum[c] / k;
        running_var[c] = var_sum[c] / k;
    }
}

/* ------------------------------------------------------------------------
 * Fused forward + backward in one sweep -- the CPU baseline's "best case"
 * single-pass port of a5 + a6 used by bench.py's cpu_baseline (kind "port").
 * Same arithmetic as the two functions above (per tensor, constant zp).
 * ---------------------------------------------------------------------- */
void vsiq_oracle_fake_quant_fwd_bwd(const float *x, const float *g, float *y, float *dx,
                                    int64_t n, float s, float z, int qmin, int qmax) {
    const float lo = (float)qmin, hi = (float)qmax;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        float v = x[i] / s;
        float t = v + z;
        float r = rintf(t);
        float q = clamp_torch(r, lo, hi);
        float d = q - z;
        y[i] = d * s;
        int m = (r >= lo) && (r <= hi);
        float grad_d = g[i] * s;
        float grad_v = m ? grad_d : 0.0f;
        dx[i] = grad_v / s;
    }
}
