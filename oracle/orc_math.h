/*
 * oracle/orc_math.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Scalar arithmetic conventions of the CPU oracle.  The reference delegates
 * every vector op to Intel MKL through JNI
 * (scalann/src/main/scala/com/mass/scalann/tensor/TensorNumeric.scala:217-465 Float,
 * :524-777 Double); MKL's internal summation order and its vsExp/vdExp are a
 * black box (third-party jar com.intel.analytics.bigdl.core.native.mkl:
 * mkl-java-x86_64-linux:2.0.0, not vendored).  The oracle therefore FIXES:
 *
 *   dot/gemm/gemv : acc = 0; for k = 0..K-1: acc = fma(a[k], b[k], acc)
 *                   (one rounding per step, ascending k)
 *   exp           : the range-reduced polynomial below (<= 2 ulp), so host and
 *                   device can agree bit for bit
 *   everything else (scal, add, sub, inv, max) : single IEEE operations in the
 *                   order the Scala code issues them.
 *
 * Compile with -ffp-contract=off: only the explicit fma() calls may fuse.
 */
#ifndef ORC_MATH_H
#define ORC_MATH_H

#include <math.h>
#include <stdint.h>
#include <string.h>

static inline float orc_f32_from_bits(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline uint32_t orc_f32_bits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline double orc_f64_from_bits(uint64_t u) { double f; memcpy(&f, &u, 8); return f; }
static inline uint64_t orc_f64_bits(double f) { uint64_t u; memcpy(&u, &f, 8); return u; }

/* exp, fp32: n = rint(x*log2e) by the 1.5*2^23 trick, Cody-Waite reduction
 * with ln2 = 0.693359375 - 2.12194440e-4, Cephes degree-5 polynomial,
 * two-step power-of-two scaling so subnormal results round once. */
static inline float orc_expf(float x)
{
    if (x != x) return x;
    if (x > 88.72283172607421875f) return INFINITY;
    if (x < -103.97208404541015625f) return 0.0f;
    const float magic = 12582912.0f;
    float t = fmaf(x, 1.44269502162933349609375f, magic);
    float n = t - magic;
    float r = fmaf(n, -0.693359375f, x);
    r = fmaf(n, 2.12194440e-4f, r);
    float p = 1.9875691500e-4f;
    p = fmaf(p, r, 1.3981999507e-3f);
    p = fmaf(p, r, 8.3334519073e-3f);
    p = fmaf(p, r, 4.1665795894e-2f);
    p = fmaf(p, r, 1.6666665459e-1f);
    p = fmaf(p, r, 5.0000001201e-1f);
    float r2 = r * r;
    float y = fmaf(p, r2, r);
    y = y + 1.0f;
    int ni = (int)n;
    int n1 = ni / 2, n2 = ni - n1;
    y = y * orc_f32_from_bits((uint32_t)(n1 + 127) << 23);
    y = y * orc_f32_from_bits((uint32_t)(n2 + 127) << 23);
    return y;
}

/* exp, fp64: same scheme, degree-13 Taylor polynomial in Horner form. */
static inline double orc_exp(double x)
{
    if (x != x) return x;
    if (x > 709.782712893384) return INFINITY;
    if (x < -745.1332191019412) return 0.0;
    const double magic = 6755399441055744.0;
    double t = fma(x, 1.4426950408889634074, magic);
    double n = t - magic;
    double r = fma(n, -6.93147180369123816490e-01, x);
    r = fma(n, -1.90821492927058770002e-10, r);
    double p = 1.6059043836821613e-10;            /* 1/13! */
    p = fma(p, r, 2.08767569878681e-09);          /* 1/12! */
    p = fma(p, r, 2.505210838544172e-08);         /* 1/11! */
    p = fma(p, r, 2.755731922398589e-07);         /* 1/10! */
    p = fma(p, r, 2.7557319223985893e-06);        /* 1/9!  */
    p = fma(p, r, 2.48015873015873e-05);          /* 1/8!  */
    p = fma(p, r, 1.984126984126984e-04);         /* 1/7!  */
    p = fma(p, r, 1.388888888888889e-03);         /* 1/6!  */
    p = fma(p, r, 8.333333333333333e-03);         /* 1/5!  */
    p = fma(p, r, 4.1666666666666664e-02);        /* 1/4!  */
    p = fma(p, r, 1.6666666666666666e-01);        /* 1/3!  */
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    double y = fma(p, r, 1.0);
    int ni = (int)n;
    int n1 = ni / 2, n2 = ni - n1;
    y = y * orc_f64_from_bits((uint64_t)(n1 + 1023) << 52);
    y = y * orc_f64_from_bits((uint64_t)(n2 + 1023) << 52);
    return y;
}

/* java.lang.Float.compare / Double.compare total order as an unsigned key:
 * -0.0 < +0.0, every NaN canonical and greatest.  Used by every stable sort of
 * the reference (Recommender.scala:77-84, otm CandidateSearcher.scala:33). */
static inline uint32_t orc_key_f32(float f)
{
    uint32_t u = (f != f) ? 0x7fc00000u : orc_f32_bits(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
static inline uint64_t orc_key_f64(double f)
{
    uint64_t u = (f != f) ? 0x7ff8000000000000ull : orc_f64_bits(f);
    return (u & 0x8000000000000000ull) ? ~u : (u | 0x8000000000000000ull);
}

#endif
