// host-side accuracy test of the pow replacement (compiled as plain C++)
#define LF_HD
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <random>
#include "lf_math.cuh"
int main() {
    std::mt19937_64 rng(1);
    std::uniform_real_distribution<double> ue(-40, 40), uy(-3, 6), um(1, 2);
    double worst = 0, worstx = 0, worsty = 0;
    long bad = 0;
    for (long i = 0; i < 20000000; ++i) {
        double x = std::ldexp(um(rng), (int)ue(rng));
        double y = uy(rng);
        if (i % 7 == 0) y = 0.6; if (i % 7 == 1) y = -0.4; if (i % 7 == 2) y = 1.0 / 0.6;
        if (i % 11 == 0) x = um(rng) - 1.0;  // (0,1)
        double ref = std::pow(x, y), got = lfm::pw(x, y);
        if (!(ref > 1e-300 && ref < 1e300)) continue;
        double e = std::fabs(got - ref) / ref;
        if (e > worst) { worst = e; worstx = x; worsty = y; }
        if (e > 1e-13) ++bad;
    }
    printf("max rel err %.3e at x=%.17g y=%.17g ; >1e-13: %ld\n", worst, worstx, worsty, bad);
    double sp[] = {0.0, 1.0, INFINITY, -1.0, NAN, 5e-324, 1e-310, 1e308};
    for (double x : sp) for (double y : {0.6, -0.4, 1.6666666666666667, 0.0}) printf("pw(%g,%g)=%g pow=%g\n", x, y, lfm::pw(x, y), std::pow(x, y));
    // log/exp separately
    double wl = 0, we = 0;
    for (long i = 0; i < 5000000; ++i) {
        double x = std::ldexp(um(rng), (int)ue(rng));
        double e1 = std::fabs(lfm::log2_fast(x) - std::log2(x)) / std::fmax(std::fabs(std::log2(x)), 1e-300);
        if (std::fabs(std::log2(x)) > 1e-3 && e1 > wl) wl = e1;
        double t = ue(rng) * 5;
        double e2 = std::fabs(lfm::exp2_fast(t) - std::exp2(t)) / std::exp2(t);
        if (e2 > we) we = e2;
    }
    printf("log2 max rel %.3e exp2 max rel %.3e\n", wl, we);
}
