// lf_math.cuh -- float64 power function for the hot kernels:  pw(x, y) = 2^(y * log2 x).
//
// CUDA's pow() is ~280 SASS instructions and exp(y*log(x)) ~170 (both carry special-case paths the hot
// path never takes).  The Newton solver and the van Genuchten conductivity spend most of their
// instructions there, so this header provides a lean evaluation for x >= 0:
//   log2(x): x = 2^e * m, m in [sqrt(1/2), sqrt(2)); s = (m-1)/(m+1); ln(m) = 2 atanh(s) as an odd
//            polynomial in s (degree 21, |s| <= 0.1716) evaluated with FMAs; the division uses a float
//            reciprocal seed and two Newton steps (no slow-path call);
//   2^t    : t = n + r, |r| <= 1/2; 2^r by its degree-13 Taylor polynomial; scaled by 2^n through the
//            exponent bits.
// Relative error of pw: about |y log2 x| * 2^-52 + 3e-16, < 2e-14 over the model's range (checked on the host
// against libm pow by tests/test_lf_math.py; the parity tolerance of the model is 1e-6).
// Special values follow pow() for x >= 0: pw(0, y>0) = 0, pw(0, y<0) = inf, pw(x, 0) = 1, pw(inf, y>0) = inf,
// pw(inf, y<0) = 0; NaN propagates; x < 0 gives NaN (never produced by the model); results below 2^-1021
// and arguments below 1e-300 flush (to 0 resp. to 1e-300).
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#include "lf_math_tables.h"
#ifndef LF_HD
#define LF_HD __host__ __device__ __forceinline__
#endif

namespace lfm {

// polynomial coefficients: constant bank on the device (an FMA can read c[bank][offset] directly; literal
// doubles would each cost two UMOV issue slots), plain statics on the host
#ifdef __CUDA_ARCH__
#define LF_COEF_QUAL __constant__
#else
#define LF_COEF_QUAL static const
#endif
LF_COEF_QUAL double c_atanh[10] = {2.0 / 21.0, 2.0 / 19.0, 2.0 / 17.0, 2.0 / 15.0, 2.0 / 13.0,
                                   2.0 / 11.0, 2.0 / 9.0,  2.0 / 7.0,  2.0 / 5.0,  2.0 / 3.0};
LF_COEF_QUAL double c_exp2[14] = {1.3691488853904128e-12, 2.5678435993488206e-11, 4.4455382718708116e-10,
                                  7.054911620801123e-09,  1.01780860092397e-07,   1.321548679014431e-06,
                                  1.5252733804059841e-05, 0.0001540353039338161,  0.0013333558146428443,
                                  0.009618129107628477,   0.05550410866482158,    0.24022650695910072,
                                  0.6931471805599453,     1.0};
LF_COEF_QUAL double c_misc[4] = {1.4142135623730951, 1.4426950408889634, 2.0355273740931033e-17, 0.0};

LF_HD double bits_to_double(uint64_t u)
{
#ifdef __CUDA_ARCH__
    return __longlong_as_double((long long)u);
#else
    double d;
    memcpy(&d, &u, 8);
    return d;
#endif
}
LF_HD uint64_t double_to_bits(double d)
{
#ifdef __CUDA_ARCH__
    return (uint64_t)__double_as_longlong(d);
#else
    uint64_t u;
    memcpy(&u, &d, 8);
    return u;
#endif
}
LF_HD double fma_(double a, double b, double c)
{
#ifdef __CUDA_ARCH__
    return __fma_rn(a, b, c);
#else
    return fma(a, b, c);
#endif
}
// NaN-agnostic minimum / maximum: 3 instructions (DSETP + 2 FSEL); fmin()/fmax() cost 7-8 on sm_100 (NaN fix-up
// plus register moves).  A NaN in `a` selects `b`, like fmin/fmax; a NaN in `b` is returned.
LF_HD double dmin(double a, double b)
{
#ifdef __CUDA_ARCH__
    double r;
    asm("{.reg .pred p; setp.lt.f64 p, %1, %2; selp.f64 %0, %1, %2, p;}" : "=d"(r) : "d"(a), "d"(b));
    return r;
#else
    return a < b ? a : b;
#endif
}
LF_HD double dmax(double a, double b)
{
#ifdef __CUDA_ARCH__
    double r;
    asm("{.reg .pred p; setp.gt.f64 p, %1, %2; selp.f64 %0, %1, %2, p;}" : "=d"(r) : "d"(a), "d"(b));
    return r;
#else
    return a > b ? a : b;
#endif
}

// a / d for a finite and d a normal, non-zero double, to ~1 ulp: hardware reciprocal seed (MUFU.RCP64H, relative
// error e0 ~ 2^-20), one Newton step on the reciprocal (e0^2), one residual correction of the quotient (e0^4 plus the
// final rounding).  Branch-free, 6 instructions (the compiler's division is 12 plus a slow-path test and call).
// On the host: the plain division.
LF_HD double div_nr(double a, double d)
{
#ifdef __CUDA_ARCH__
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
    r = fma_(fma_(-d, r, 1.0), r, r);
    const double q = a * r;
    return fma_(fma_(-d, q, a), r, q);
#else
    return a / d;
#endif
}
LF_HD double div_small(double f, double d) { return div_nr(f, d); }  // d in [1.7, 2.42]

// log2 of a positive, finite, normal double
LF_HD double log2_fast(double x)
{
    const uint64_t u = double_to_bits(x);
    int e = (int)(u >> 52) - 1023;
    double m = bits_to_double((u & 0x000fffffffffffffull) | 0x3ff0000000000000ull);
    if (m > c_misc[0]) {
        m *= 0.5;
        e += 1;
    }
    const double s = div_small(m - 1.0, m + 1.0);
    const double z = s * s;
    double p = c_atanh[0];  // 2 atanh(s) = s * (2 + z * (2/3 + z * (2/5 + ...)))
#pragma unroll
    for (int k = 1; k < 10; ++k) p = fma_(p, z, c_atanh[k]);
    const double ln_hi = 2.0 * s, ln_lo = (s * z) * p;  // ln(m) = ln_hi + ln_lo, |ln_lo| << |ln_hi|
    const double LOG2E_HI = c_misc[1], LOG2E_LO = c_misc[2];
    const double hi = ln_hi * LOG2E_HI;
    const double lo = fma_(ln_hi, LOG2E_HI, -hi) + fma_(ln_lo, LOG2E_HI, ln_hi * LOG2E_LO);
    return ((double)e + hi) + lo;
}

// 2^t for -1021 <= t <= 1023.5
LF_HD double exp2_fast(double t)
{
    const double n = rint(t);
    const double r = t - n;  // |r| <= 1/2, exact
    double p = c_exp2[0];  // ln2^k / k!, k = 13 .. 0
#pragma unroll
    for (int k = 1; k < 14; ++k) p = fma_(p, r, c_exp2[k]);
    const int64_t ni = (int64_t)n;
    return p * bits_to_double((uint64_t)(ni + 1023) << 52);
}

// Branch-free: the fast path is evaluated unconditionally on a clamped argument and the special values are
// patched in with selects, so the compiler can interleave several independent pw() chains (the three soil layers,
// the three overland routers) instead of serialising them at basic-block boundaries.
LF_HD double pw(double x, double y)
{
    const double INF = bits_to_double(0x7ff0000000000000ull);
    const double xs = dmin(dmax(x, 1e-300), 1.7976931348623157e308);  // zero / denormal / inf are patched below
    const double t = y * log2_fast(xs);
    double r = exp2_fast(dmin(dmax(t, -1021.0), 1023.0));
    r = t < -1021.0 ? 0.0 : r;
    r = t > 1023.0 ? INF : r;
    const double at_zero = y > 0.0 ? 0.0 : (y < 0.0 ? INF : 1.0);
    const double at_inf = y > 0.0 ? INF : (y < 0.0 ? 0.0 : 1.0);
    r = x == 0.0 ? at_zero : r;
    r = x > 1.7976931348623157e308 ? at_inf : r;
    r = (x >= 0.0 && y == y) ? r : (y == 0.0 ? 1.0 : bits_to_double(0x7ff8000000000000ull));  // NaN / negative base
    return r;
}


// ---------------------------------------------------------------------------------------------------------------
// Table-driven variants for the soil column kernel (lf_soil_kernel.cuh), ~40 instructions per power instead of ~75.
// The column kernel evaluates ~9 powers x^y with 0 <= x <= 1 and y > 0 per column (van Genuchten conductivity,
// Xinanjiang infiltration, preferential flow: soilloop.py:360-383, :178-216) and is bound by instruction issue.
//   log2_tab(x): bits(x) - OFF puts z = x / 2^e in [0.707, 1.414); the top 7 fraction bits pick a table interval
//                with centre c; r = z * (1/c) - 1 by one FMA (|r| <= 2^-8); log2 z = logc + r * P6(r).
//                ABSOLUTE error <= ~3e-16 + |result| * 2^-52 (what a power needs; not relative accuracy near x = 1).
//   exp2_tab(t): t = (128 k + j)/128 + r, |r| <= 2^-8: 2^t = 2^k * T[j] * (1 + r * P5(r)); relative error <= 3e-16.
//   pw_tab(x,y): 2^(y * log2 x) for 0 <= x, y > 0, y * log2 x <= 1020: relative error <= ~(1 + |y log2 x|) * 3e-16
//                (host-tested against libm in tests/test_lf_math.py).  pw_tab(0, y) = 0; results below 2^-1021 are 0.
// The tables (lf_math_tables.h, generated by tools/gen_math_tables.py) are 3 KB: kernels copy them to shared
// memory once per block (tab_to_shared) -- lanes index them independently, which constant memory would serialise.
// ---------------------------------------------------------------------------------------------------------------
struct MathTab {
    double log2_tab[2 * LF_LOG2_TAB_N];  // {invc, logc}
    double exp2_tab[LF_EXP2_TAB_N];
};
#ifdef __CUDACC__
// one 3 KB object, 16-byte aligned: k_soil_staged fetches it with one bulk async copy; rare paths read it in place
__device__ __align__(16) const MathTab g_mathtab = {{LF_LOG2_TAB_VALUES}, {LF_EXP2_TAB_VALUES}};
LF_COEF_QUAL double c_log2_poly[6] = {LF_LOG2_POLY_VALUES};
LF_COEF_QUAL double c_exp2_poly[5] = {LF_EXP2_POLY_VALUES};
#else
static const MathTab g_mathtab = {{LF_LOG2_TAB_VALUES}, {LF_EXP2_TAB_VALUES}};
static const double c_log2_poly[6] = {LF_LOG2_POLY_VALUES};
static const double c_exp2_poly[5] = {LF_EXP2_POLY_VALUES};
#endif

#ifdef __CUDACC__
// called by every thread of the block, followed by __syncthreads()
__device__ __forceinline__ void tab_to_shared(MathTab *s, int tid, int nthreads)
{
    for (int k = tid; k < 2 * LF_LOG2_TAB_N; k += nthreads) s->log2_tab[k] = g_mathtab.log2_tab[k];
    for (int k = tid; k < LF_EXP2_TAB_N; k += nthreads) s->exp2_tab[k] = g_mathtab.exp2_tab[k];
}
#endif

// sqrt(x) for 0 <= x < 1e300 to ~1 ulp: reciprocal-square-root seed (MUFU.RSQ64H, e0 ~ 2^-20), one coupled Newton
// step (g ~ sqrt x, h ~ 1/(2 sqrt x), error 1.5 e0^2) and a final residual correction (e0^4); branch-free, 8
// instructions (the compiler's sqrt carries a slow-path call).  Host: sqrt().
LF_HD double sqrt_nr(double x)
{
#ifdef __CUDA_ARCH__
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double g = x * y, h = 0.5 * y;
    double r = fma_(-g, h, 0.5);
    g = fma_(g, r, g);
    h = fma_(h, r, h);
    g = fma_(fma_(-g, g, x), h, g);
    return x > 0.0 ? g : 0.0;
#else
    return sqrt(x);
#endif
}

LF_HD double log2_tab(double x, const double *T)
{
    const uint64_t u = double_to_bits(x);
    const uint32_t hi = (uint32_t)(u >> 32), lo = (uint32_t)u;
    const uint32_t t = hi - LF_LOG2_OFF_HI;  // the low word of OFF is zero
    const int e = (int)t >> 20;
    const int j = (int)((t >> 13) & (LF_LOG2_TAB_N - 1));
    const double z = bits_to_double(((uint64_t)(hi - (t & 0xfff00000u)) << 32) | lo);
    const double invc = T[2 * j], logc = T[2 * j + 1];
    const double r = fma_(z, invc, -1.0);
    double p = c_log2_poly[5];
#pragma unroll
    for (int k = 4; k >= 0; --k) p = fma_(p, r, c_log2_poly[k]);
    return fma_(r, p, (double)e + logc);
}

// 2^t for -1021 <= t <= 1020
LF_HD double exp2_tab(double t, const double *T)
{
    const double SHIFT = 6755399441055744.0;  // 1.5 * 2^52
    const double kd = fma_(t, (double)LF_EXP2_TAB_N, SHIFT);
    const int32_t n = (int32_t)(uint32_t)double_to_bits(kd);  // rint(128 t)
    const double r = fma_(kd - SHIFT, -1.0 / LF_EXP2_TAB_N, t);  // exact
    const uint64_t sb = double_to_bits(T[n & (LF_EXP2_TAB_N - 1)]) + ((uint64_t)(int64_t)(n >> 7) << 52);
    const double s = bits_to_double(sb);
    double p = c_exp2_poly[4];
#pragma unroll
    for (int k = 3; k >= 0; --k) p = fma_(p, r, c_exp2_poly[k]);
    return fma_(s, r * p, s);
}

// x^y for x >= 0 (finite), y > 0, y*log2(x) <= 1020.  EXACT0: results below 2^-1021 (and x == 0) return exactly 0;
// otherwise they return ~2^-1021 (callers that subtract the power from 1 do not care).
template <bool EXACT0 = true>
LF_HD double pw_tab(double x, double y, const MathTab *M)
{
    const double t = y * log2_tab(x, M->log2_tab);
    const double r = exp2_tab(dmax(t, -1021.0), M->exp2_tab);
    if (EXACT0) return (t < -1021.0 || x == 0.0) ? 0.0 : r;
    return r;
}
// e^x for x <= 0 (same scheme: 2^(x log2 e)); relative error <= (1 + |x|) * 3e-16
LF_HD double exp_neg_tab(double x, const MathTab *M)
{
    const double t = x * 1.4426950408889634;
    const double r = exp2_tab(dmax(t, -1021.0), M->exp2_tab);
    return t < -1021.0 ? 0.0 : r;
}

}  // namespace lfm
