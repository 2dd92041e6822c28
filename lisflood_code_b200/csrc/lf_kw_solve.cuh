// lf_kw_solve.cuh -- per-pixel kinematic-wave solve (device).
//
// Restates solve1Pixel / closureError of the reference
// (hydrological_modules/kinematic_wave_parallel_tools.py:48-92): find Q >= 0 with
//     f(Q) = Q + a*Q^beta - C = 0,   C = upstream inflow + a*Qold^beta + lateral inflow,
// Q clamped at 1e-12 and mapped to 0 there, C <= 1e-12 -> 0.
//
// Two evaluations of the same root:
//
//  * general beta: the reference's bracketed initial guess and Newton iteration in Q, with
//    x^y from lf_math.cuh, one power per iteration (Q^beta = Q * Q^(beta-1)), and an extra exit when
//    the relative Newton step is below 1e-8 (then converged to < 1e-16: the error constant of this f is
//    <= 0.2/Q), which removes the reference's confirmation iteration and its rare 2-value limit cycles.
//
//  * beta == 0.6 (Manning's 3/5 -- the value LISFLOOD ships, settings `beta`): with z = Q^(1/5) every power the
//    iteration needs is a product: Q^beta = z^3, Q^(beta-1) = 1/z^2.  The SAME Newton iterates in Q are formed
//    (same update, clamp and stopping tests), and z is refreshed after each update by a fifth root (float seed +
//    two Newton steps, ~20 instructions) -- no log/exp at all.  The state carried between routing steps is z, so
//    a*Qold^beta = a*z^3 is free and the discharge handed downstream is z^5.  The bracketed initial guess is the
//    reference's too (C^(-2/5) from a fifth root, the two 5/3 powers from cube roots), in float64: the
//    reference stops as soon as |f| <= 1e-12, which for near-dry pixels happens after 0-2 iterations, so the
//    returned value depends on the guess and on every iterate -- they must be reproduced, not just the root.
//    (A pure Newton on the quintic z^5 + a z^3 - C, or a float-precision guess, is cheaper but returns values that
//    differ from the reference's by up to its own 1e-12 absolute tolerance, i.e. 1e-4 relative on near-dry
//    pixels and 100 % on the derived volumes there -- measured and rejected.)
//    ~250 instructions per solve instead of ~1400: the routing kernels move from FP64-bound towards HBM-bound.
#pragma once
#include "lf_math.cuh"

namespace lfkw {

constexpr double NEWTON_TOL = 1e-12;  // tools:26
constexpr int MAX_ITERS = 3000;       // tools:27
constexpr double Z_MIN = 0.0039810717055349725;  // (1e-12)^(1/5)

struct Params {
    double beta, inv_beta, b_minus_1;
    int quintic;  // beta == 0.6: state is z = Q^(1/5)
    int pad_;     // explicit: kernel-argument structs are compared byte-wise (CUDA-graph cache)
};

__host__ __device__ inline Params make_params(double beta)
{
    Params P;
    P.beta = beta;
    P.inv_beta = 1 / beta;   // kinematic_wave_parallel.py:124
    P.b_minus_1 = beta - 1;  // :125
    P.quintic = (beta == 0.6) ? 1 : 0;
    P.pad_ = 0;
    return P;
}

__device__ __forceinline__ double pw(double x, double y) { return lfm::pw(x, y); }
__device__ __forceinline__ double pow5(double z)
{
    const double z2 = z * z;
    return z2 * z2 * z;
}

// ---- general beta: state is Q --------------------------------------------------------------------
// U: sum of the (new) discharge of the upstream pixels, in slot order.
__device__ __forceinline__ double solve(double U, double q_old, double lateral, double a, const Params &P)
{
    // constant = a_dx_div_dt * Qold**b + lateral_inflow   (kinematic_wave_parallel.py:174-175)
    double c = U + (a * pw(q_old, P.beta) + lateral);
    if (c <= NEWTON_TOL) return 0.0;  // tools:60-62
    double ba = P.beta * a;           // b_a_dx_div_dt, kinematic_wave_parallel.py:127
    double t = ba * pw(c, P.b_minus_1);
    double secant = (t <= 1.0) ? c / (1.0 + t) : c / (1.0 + pw(t, P.inv_beta));
    double other = pw((c - secant) / a, P.inv_beta);
    double q = (secant + other) / 2.0;
    double p = pw(q, P.b_minus_1);
    double err = q + a * (q * p) - c;
    int count = 0;
    while (fabs(err) > NEWTON_TOL && count < MAX_ITERS) {
        double prev = q;
        q -= err / (1.0 + ba * p);
        q = lfm::dmax(q, NEWTON_TOL);
        if (fabs(q - prev) <= 1e-8 * q) break;  // includes q == prev
        p = pw(q, P.b_minus_1);
        err = q + a * (q * p) - c;
        ++count;
    }
    if (q == NEWTON_TOL) q = 0.0;  // tools:79-80
    return q;
}

// ---- beta == 3/5: state is z = Q^(1/5) -----------------------------------------------------------
// hardware approximations used only as SEEDS (every seed is refined in float64 below)
__device__ __forceinline__ float lg2_approx(float x)
{
    float r;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float ex2_approx(float x)
{
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ double rcp64_approx(double x)  // MUFU.RCP64H: ~20 good bits, no conversions
{
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    return r;
}
__device__ __forceinline__ double div_nr(double a, double d) { return lfm::div_nr(a, d); }
// fifth root of q in [1e-30, 1e30]: approximate seed (rel. error ~1e-6) + two Newton steps whose divisions use the
// approximate reciprocal (Newton is self-correcting): error ~1e-16
__device__ __forceinline__ double root5(double q)
{
    double w = (double)ex2_approx(0.2f * lg2_approx((float)q));
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const double w2 = w * w, w4 = w2 * w2;
        w = lfm::fma_(-lfm::fma_(w4, w, -q), rcp64_approx(5.0 * w4), w);
    }
    return w;
}
// cube root of x in [1e-30, 1e30], same scheme
__device__ __forceinline__ double root3(double x)
{
    double w = (double)ex2_approx(0.33333334f * lg2_approx((float)x));
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const double w2 = w * w;
        w = lfm::fma_(-lfm::fma_(w2, w, -x), rcp64_approx(3.0 * w2), w);
    }
    return w;
}
__device__ __forceinline__ bool in_float_range(double x) { return x > 1e-30 && x < 1e30; }
// z of a discharge handed in by the caller (any magnitude, 0 -> 0, negative / NaN -> NaN)
__device__ __forceinline__ double z_of_q(double q)
{
    if (in_float_range(q)) return root5(q);
    if (q == 0.0) return 0.0;
    return lfm::pw(q, 0.2);
}
// x^(5/3) for 0 <= x < 1e30 (below 1e-30 the result, < 1e-50, is taken as 0)
__device__ __forceinline__ double pow_5_3(double x) { return x > 1e-30 ? pow5(root3(x)) : 0.0; }

// U: sum of z_k^5 over the upstream pixels; returns z_new (0 when the discharge is 0).
// Same initial guess and Newton iterates as the reference (tools:64-80), in float64; Q^beta = z^3 and
// Q^(beta-1) = 1/z^2 come from z = Q^(1/5), refreshed after every update by root5().
__device__ __forceinline__ double solve_z(double U, double z_old, double lateral, double a)
{
    const double c = U + (a * (z_old * z_old * z_old) + lateral);
    if (c <= NEWTON_TOL) return 0.0;  // tools:60-62
    if (!(c < 1.0e30)) return c < 1.0e300 ? lfm::pw(c, 0.2) : c;  // absurd / NaN / Inf inflow: propagate
    // bracketed initial guess, tools:64-70
    const double ba = 0.6 * a;
    const double zc = root5(c);
    const double t = div_nr(ba, zc * zc);  // b*a * C^(b-1)
    const double secant = div_nr(c, 1.0 + ((t <= 1.0) ? t : pow_5_3(lfm::dmin(t, 1e30))));
    const double other = pow_5_3(lfm::dmin(div_nr(c - secant, a), 1e29));
    double q = (secant + other) / 2.0;
    double z = root5(lfm::dmin(lfm::dmax(q, 1e-30), 1e30));  // q >= C / (2 (1 + t^(5/3))) > 0
    double err = q + a * (z * z * z) - c;
    int count = 0;
    while (fabs(err) > NEWTON_TOL && count < MAX_ITERS) {
        const double z2 = z * z;
        double qn = q - div_nr(err * z2, z2 + ba);  // q - err / (1 + b*a*q^(b-1))
        qn = lfm::dmax(qn, NEWTON_TOL);
        const bool small = fabs(qn - q) <= 1e-8 * qn;  // includes q == prev
        q = qn;
        z = root5(q);
        if (small) break;
        err = q + a * (z * z * z) - c;
        ++count;
    }
    if (q == NEWTON_TOL) return 0.0;  // tools:79-80
    return z;
}

}  // namespace lfkw
