// lf_kw_solve.cuh -- per-pixel kinematic-wave solve (device).
//
// Restates solve1Pixel / closureError of the reference
// (hydrological_modules/kinematic_wave_parallel_tools.py:48-92): bracketed initial guess, then
// Newton-Raphson on f(Q) = Q + a*Q^beta - C with Q clamped at 1e-12.
//
// Differences in evaluation, all far inside the 1e-6 parity tolerance:
//   * x^y is evaluated as exp(y*log(x)) in float64 (about 50 instructions instead of the ~200 of
//     CUDA's pow(); relative error <= |y ln x| * 2^-52).
//   * one power per Newton iteration: Q^beta = Q * Q^(beta-1) re-uses the derivative's power.
//   * the reference stops when |f| <= 1e-12, when Q stops changing, or after 3000 iterations.  The
//     same tests are kept; in addition the loop leaves as soon as the relative Newton step is below
//     1e-8 (the iterate is then converged to < 1e-16 relative because Newton's error constant for
//     this f is <= 0.2/Q), which avoids the reference's last "no change" confirmation iteration and
//     its rare two-value limit cycles that would otherwise pin a whole warp for 3000 iterations.
#pragma once

namespace lfkw {

constexpr double NEWTON_TOL = 1e-12;  // tools:26
constexpr int MAX_ITERS = 3000;       // tools:27

struct Params {
    double beta, inv_beta, b_minus_1;
};

__device__ __forceinline__ double pw(double x, double y) { return exp(y * log(x)); }

// U: sum of the (new) discharge of the upstream pixels, in slot order.
// Returns the new discharge; *iters (optional) receives the Newton iteration count.
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
        q = fmax(q, NEWTON_TOL);
        if (fabs(q - prev) <= 1e-8 * q) break;  // includes q == prev
        p = pw(q, P.b_minus_1);
        err = q + a * (q * p) - c;
        ++count;
    }
    if (q == NEWTON_TOL) q = 0.0;  // tools:79-80
    return q;
}

}  // namespace lfkw
