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
// a / d for a finite and d a normal, non-zero double, to ~1 ulp: hardware reciprocal seed (MUFU.RCP64H, ~20 bits),
// two Newton steps on the reciprocal, one correction of the quotient.  Branch-free, 9 instructions (the compiler's
// division carries a slow-path test and call).  On the host: the plain division.
LF_HD double div_nr(double a, double d)
{
#ifdef __CUDA_ARCH__
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
    r = fma_(fma_(-d, r, 1.0), r, r);
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
    const double xs = fmin(fmax(x, 1e-300), 1.7976931348623157e308);  // zero / denormal / inf are patched below
    const double t = y * log2_fast(xs);
    double r = exp2_fast(fmin(fmax(t, -1021.0), 1023.0));
    r = t < -1021.0 ? 0.0 : r;
    r = t > 1023.0 ? INF : r;
    const double at_zero = y > 0.0 ? 0.0 : (y < 0.0 ? INF : 1.0);
    const double at_inf = y > 0.0 ? INF : (y < 0.0 ? 0.0 : 1.0);
    r = x == 0.0 ? at_zero : r;
    r = x > 1.7976931348623157e308 ? at_inf : r;
    r = (x >= 0.0 && y == y) ? r : (y == 0.0 ? 1.0 : bits_to_double(0x7ff8000000000000ull));  // NaN / negative base
    return r;
}

}  // namespace lfm
