/* fastpath_proof.c -- CPU emulation of the kernels' division-free element arithmetic (TEST INFRASTRUCTURE).
 *
 * The CUDA kernels never divide per element: they hoist r = RN(1/s) and recover the correctly rounded quotient with
 * exact-residual FMAs (csrc/common.cuh: div_fast / dx_fast), falling back to the IEEE sequence for a whole vector when an
 * input lies outside 2^-60 <= |x| < 2^61 (or the scale outside [2^-40, 2^40]).  The GPU self-test
 * (vsiq_selftest_division) compares both paths over all 2^32 inputs on the device; this file does the same on the host,
 * where fmaf / float division are the same IEEE-754 operations, so the claim "bit-identical to x / s, i.e. to the
 * reference's torch.div" (quantizers/uniform.py:54,95) is checked independently of any GPU.
 *
 *   mode 0: x / s                 vs  copysign(q2, q0),  q0 = x*r, q1 = q0 + (x - q0*s)*r, q2 = q1 + (x - q1*s)*r
 *   mode 1: RN(RN(g*s) / s)       vs  copysign(g + (RN(g*s) - g*s)*r, g)          (the dx path of the STE backward)
 *
 * vsiq_proof_division(s, mode, first, stride) walks bit patterns first, first+stride, ... < 2^32 and returns the number
 * of in-range inputs whose fast result differs from the IEEE one (NaNs compare equal); *covered receives how many inputs
 * took the fast path.  Build: gcc -O2 -mfma -ffp-contract=off -fopenmp (see Makefile).
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

static inline float bits_to_float(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline uint32_t float_to_bits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }

static inline int same(float a, float b) {
    if (a != a && b != b) return 1;
    return float_to_bits(a) == float_to_bits(b);
}

#define FAST_LO 8.673617379884035e-19f /* 2^-60 */
#define FAST_HI 2.305843009213694e18f  /* 2^61  */

unsigned long long vsiq_proof_division(float s, int mode, uint32_t first, uint32_t stride,
                                       unsigned long long *covered) {
    const float as = fabsf(s);
    unsigned long long wrong = 0, fast = 0;
    if (!(as >= 9.094947017729282e-13f && as <= 1.099511627776e12f)) { /* scale outside [2^-40, 2^40]: always slow */
        if (covered) *covered = 0;
        return 0;
    }
    const float r = 1.0f / s; /* RN(1/s), what __frcp_rn returns */
    if (stride == 0) stride = 1;
    const uint64_t n = ((uint64_t)0x100000000ull - first + stride - 1) / stride;
#pragma omp parallel for reduction(+ : wrong, fast) schedule(static)
    for (uint64_t i = 0; i < n; ++i) {
        const float x = bits_to_float((uint32_t)(first + i * stride));
        const float ax = fabsf(x);
        const int in = (ax >= FAST_LO) && (ax < FAST_HI);
        if (!in && x != 0.0f) continue; /* the kernels recompute such vectors with the IEEE sequence */
        ++fast;
        float got, want;
        if (mode == 0) {
            const float q0 = x * r;
            const float e0 = fmaf(-q0, s, x);
            const float q1 = fmaf(e0, r, q0);
            const float e1 = fmaf(-q1, s, x);
            const float q2 = fmaf(e1, r, q1);
            got = copysignf(q2, q0);
            want = x / s;
        } else {
            const float gd = x * s;
            const float rho = fmaf(-x, s, gd);
            got = copysignf(fmaf(rho, r, x), x);
            want = gd / s;
        }
        if (!same(got, want)) ++wrong;
    }
    if (covered) *covered = fast;
    return wrong;
}

/* Integer-code export (csrc/fake_quant.cu: codes_vec).  On the fast path the pre-rounding clamp leaves t in
 * [qmin - 0.5, qmax + 0.5]; the kernel then takes the code from the low mantissa bits of RN(t + 1.5 * 2^23) instead of
 * converting rint(t) (the reference's torch.round + clamp, quantizers/uniform.py:54,95).  vsiq_proof_code_magic walks
 * EVERY float in [lo, hi] (both signs of zero and all denormals included) and counts the values whose two's-complement
 * code of `bits` bits differs from (int)rintf(t); *covered receives how many floats were visited. */
unsigned long long vsiq_proof_code_magic(float lo, float hi, int bits, unsigned long long *covered) {
    const float magic = 12582912.0f; /* 0x4B400000 */
    const uint32_t mask = bits >= 32 ? 0xffffffffu : ((1u << bits) - 1u);
    unsigned long long wrong = 0, seen = 0;
#pragma omp parallel for reduction(+ : wrong, seen) schedule(static)
    for (uint64_t u = 0; u < 0x100000000ull; ++u) {
        const float t = bits_to_float((uint32_t)u);
        if (!(t >= lo && t <= hi)) continue; /* NaN fails both compares */
        ++seen;
        const uint32_t got = float_to_bits(t + magic) & mask;
        const uint32_t want = (uint32_t)(int32_t)rintf(t) & mask;
        if (got != want) ++wrong;
    }
    if (covered) *covered = seen;
    return wrong;
}
