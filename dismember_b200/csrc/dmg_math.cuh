// dmg_math.cuh -- scalar arithmetic spec of the engine (device side).
//
// The reference's arithmetic goes through Intel MKL (JNI) whose operation order is
// not observable; DESIGN.md "Arithmetic spec" fixes one order and this header is its
// device statement:
//   * every dot product / GEMM element is ONE sequential chain over ascending k:
//       acc = 0; acc = fma(a[k], b[k], acc)
//   * exp is the range-reduced polynomial below (fp32: Cephes degree 5, fp64: Taylor
//     degree 13), scaled by two exact powers of two
//   * all other steps are single IEEE operations in the order the Scala code issues them
//     (Mask.scala:17-33 scal then overwrite; SoftMax.scala:27-42 max, sub, exp, sum, *inv;
//     Linear.scala:41-47 product then bias).
// Only explicit __fmaf_rn/__fma_rn may fuse: the translation units are compiled with
// -fmad=false so that a*b+c written as two operations stays two roundings.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace dmg {

__device__ __forceinline__ float fma_(float a, float b, float c) { return __fmaf_rn(a, b, c); }
__device__ __forceinline__ double fma_(double a, double b, double c) { return __fma_rn(a, b, c); }
__device__ __forceinline__ float mul_(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ double mul_(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float add_(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ double add_(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ float sub_(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ double sub_(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ float inv_(float a) { return __fdiv_rn(1.0f, a); }
__device__ __forceinline__ double inv_(double a) { return __ddiv_rn(1.0, a); }

__device__ __forceinline__ float exp_(float x)
{
    if (x != x) return x;
    if (x > 88.72283172607421875f) return __int_as_float(0x7f800000);
    if (x < -103.97208404541015625f) return 0.0f;
    const float magic = 12582912.0f;
    float t = __fmaf_rn(x, 1.44269502162933349609375f, magic);
    float n = __fsub_rn(t, magic);
    float r = __fmaf_rn(n, -0.693359375f, x);
    r = __fmaf_rn(n, 2.12194440e-4f, r);
    float p = 1.9875691500e-4f;
    p = __fmaf_rn(p, r, 1.3981999507e-3f);
    p = __fmaf_rn(p, r, 8.3334519073e-3f);
    p = __fmaf_rn(p, r, 4.1665795894e-2f);
    p = __fmaf_rn(p, r, 1.6666665459e-1f);
    p = __fmaf_rn(p, r, 5.0000001201e-1f);
    float r2 = __fmul_rn(r, r);
    float y = __fmaf_rn(p, r2, r);
    y = __fadd_rn(y, 1.0f);
    int ni = (int)n;
    int n1 = ni / 2, n2 = ni - n1;
    y = __fmul_rn(y, __int_as_float((n1 + 127) << 23));
    y = __fmul_rn(y, __int_as_float((n2 + 127) << 23));
    return y;
}

__device__ __forceinline__ double exp_(double x)
{
    if (x != x) return x;
    if (x > 709.782712893384) return __longlong_as_double(0x7ff0000000000000LL);
    if (x < -745.1332191019412) return 0.0;
    const double magic = 6755399441055744.0;
    double t = __fma_rn(x, 1.4426950408889634074, magic);
    double n = __dsub_rn(t, magic);
    double r = __fma_rn(n, -6.93147180369123816490e-01, x);
    r = __fma_rn(n, -1.90821492927058770002e-10, r);
    double p = 1.6059043836821613e-10;
    p = __fma_rn(p, r, 2.08767569878681e-09);
    p = __fma_rn(p, r, 2.505210838544172e-08);
    p = __fma_rn(p, r, 2.755731922398589e-07);
    p = __fma_rn(p, r, 2.7557319223985893e-06);
    p = __fma_rn(p, r, 2.48015873015873e-05);
    p = __fma_rn(p, r, 1.984126984126984e-04);
    p = __fma_rn(p, r, 1.388888888888889e-03);
    p = __fma_rn(p, r, 8.333333333333333e-03);
    p = __fma_rn(p, r, 4.1666666666666664e-02);
    p = __fma_rn(p, r, 1.6666666666666666e-01);
    p = __fma_rn(p, r, 0.5);
    p = __fma_rn(p, r, 1.0);
    double y = __fma_rn(p, r, 1.0);
    int ni = (int)n;
    int n1 = ni / 2, n2 = ni - n1;
    y = __dmul_rn(y, __longlong_as_double((long long)(n1 + 1023) << 52));
    y = __dmul_rn(y, __longlong_as_double((long long)(n2 + 1023) << 52));
    return y;
}

// java.lang.Float.compare / Double.compare total order as an unsigned key
// (-0.0 < +0.0, NaN canonical and greatest) -- every stable sort of the reference:
// Recommender.scala:77-84, otm CandidateSearcher.scala:33, dr CandidateSearcher.scala:49.
__device__ __forceinline__ uint32_t order_key(float f)
{
    uint32_t u = (f != f) ? 0x7fc00000u : __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ uint64_t order_key(double f)
{
    uint64_t u = (f != f) ? 0x7ff8000000000000ull : (uint64_t)__double_as_longlong(f);
    return (u & 0x8000000000000000ull) ? ~u : (u | 0x8000000000000000ull);
}

// math.max(x, 0) of ReLU.scala:30-44 (NaN propagates, -0 -> +0)
__device__ __forceinline__ float relu_(float x) { return x > 0.0f ? x : (x != x ? x : 0.0f); }
__device__ __forceinline__ double relu_(double x) { return x > 0.0 ? x : (x != x ? x : 0.0); }

template <typename real> struct mask_value;
template <> struct mask_value<float> { static __device__ __forceinline__ float get() { return -3.4028234663852886e+38f; } };
template <> struct mask_value<double> { static __device__ __forceinline__ double get() { return -3.4028234663852886e+38; } };

}  // namespace dmg
