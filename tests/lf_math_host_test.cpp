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
    // table-driven variants (soil column kernel): x in [0,1], y > 0
    {
        lfm::MathTab M;
        memcpy(M.log2_tab, lfm::g_mathtab.log2_tab, sizeof(M.log2_tab));
        memcpy(M.exp2_tab, lfm::g_mathtab.exp2_tab, sizeof(M.exp2_tab));
        std::uniform_real_distribution<double> u01(0, 1), uyy(0.04, 25), uex(-60, 0);
        double wt = 0, wtn = 0, wx = 0, wy = 0, wlabs = 0, wexp = 0;
        long badt = 0;
        for (long i = 0; i < 20000000; ++i) {
            double x = u01(rng);
            if (i % 5 == 0) x = std::ldexp(um(rng), (int)uex(rng));   // small arguments
            if (i % 5 == 1) x = 1.0 - std::ldexp(um(rng), (int)uex(rng) - 1);  // close to 1
            double y = uyy(rng);
            if (i % 3 == 0) y = u01(rng) * 0.5 + 0.05;
            double ref = std::pow(x, y), got = lfm::pw_tab<true>(x, y, &M);
            double tl = std::fabs(y * std::log2(x));
            if (ref > 1e-290) {
                double e = std::fabs(got - ref) / ref;
                if (e > wt) { wt = e; wx = x; wy = y; }
                double en = e / (1.0 + tl);
                if (en > wtn) wtn = en;
                if (e > 1e-13) ++badt;
            }
            double la = std::fabs(lfm::log2_tab(x, M.log2_tab) - std::log2(x)) / (1.0 + std::fabs(std::log2(x)));
            if (la > wlabs) wlabs = la;
            double xe = uex(rng) * 10;  // [-600, 0]
            double ee = std::fabs(lfm::exp_neg_tab(xe, &M) - std::exp(xe)) / std::exp(xe) / (1.0 + std::fabs(xe));
            if (ee > wexp) wexp = ee;
        }
        printf("pw_tab max rel err %.3e at x=%.17g y=%.17g ; normalised by (1+|y log2 x|) %.3e ; >1e-13: %ld\n", wt, wx, wy, wtn, badt);
        printf("log2_tab max err/(1+|log2 x|) %.3e ; exp_neg_tab max rel/(1+|x|) %.3e\n", wlabs, wexp);
        printf("pw_tab(0,0.3)=%g pw_tab(1,7.5)=%.17g pw_tab(1e-200,9)=%g pw_tab<false>(0,0.3)=%g exp_neg_tab(0)=%.17g exp_neg_tab(-1000)=%g\n",
               lfm::pw_tab<true>(0.0, 0.3, &M), lfm::pw_tab<true>(1.0, 7.5, &M), lfm::pw_tab<true>(1e-200, 9.0, &M),
               lfm::pw_tab<false>(0.0, 0.3, &M), lfm::exp_neg_tab(0.0, &M), lfm::exp_neg_tab(-1000.0, &M));
    }
}
