// lf_soil_kernel.cuh -- fused per-cell kernel of the soil / canopy / groundwater stack (device).
//
// ONE fused kernel (k_soil_fused) replaces the stages the reference executes as separate NumPy/Numba passes:
//   soilloop.dynamic_canopy     hydrological_modules/soilloop.py:519-627 (+ kernel :27-70)
//   soilloop.dynamic_soil       hydrological_modules/soilloop.py:630-665 (+ kernel :78-355)
//   opensealed.dynamic          hydrological_modules/opensealed.py:41-71
//   soil.dynamic_perpixel       hydrological_modules/soil.py:471-514 (deffraction: Lisflood_initial.py:393-396)
//   groundwater.dynamic         hydrological_modules/groundwater.py:134-180
//   surface_routing.dynamic     hydrological_modules/surface_routing.py:122-149 (runoff components only)
// There is no neighbour access anywhere in these stages: the kernel is a pure stream over
// SoA float64 maps stored in the overland-flow router's position order.
//
// Work decomposition: a block owns a tile of TILE consecutive pixels and has 3*TILE threads; warp-uniform
// v = thread / TILE is the vegetation fraction, so a thread integrates one (fraction, pixel) soil column.  The three
// columns of a pixel run in the same block at the same time: the per-pixel maps (forcing, Xinanjiang b, ...) and the
// land-use parameter rows shared between fractions are fetched from DRAM once (the 2nd and 3rd reader hit L1/L2), the
// fraction-weighted column results are exchanged through shared memory, and the v == 0 thread of the pixel finishes
// the per-pixel part (open water / sealed soil, sums over fractions, groundwater, runoff components).  DRAM traffic
// is the algorithmic 889 B per cell; the first version (separate column and pixel kernels over a
// (fraction, pixel) grid) moved twice that.
//
// Divergence control (the adaptive Darcy sub-stepping, soilloop.py:237-312): the number of sub-steps is
// per column (mean ~1.5, 99th percentile ~20, max ~100), so a warp that simply loops pays the maximum of its
// 32 lanes (~12 on average, measured).  k_soil_fused therefore completes only the columns that need ONE
// sub-step (~99 %) and appends the others, warp-aggregated, to one of six lists bucketed by sub-step count
// (2-3, 4-7, 8-15, 16-31, 32-63, 64+); k_soil_veg_deferred then integrates each list with one thread per
// column, so lanes of a warp differ by at most 2x in trip count.  Deferred columns recompute their (cheap)
// prologue instead of spilling ~30 doubles of state.  A pixel with a deferred column is flagged (pix_deferred);
// its finished columns park their contributions in the c* maps and k_soil_pixel_flagged completes it after the
// deferred lists have run.  Results do not depend on list order.
//
// Arithmetic: float64, unfused multiply-add like the reference (--fmad=false) except inside the library functions
// of lf_math.cuh (table-driven x^y and e^x, Newton division and square root, 3-instruction min/max): the kernel is
// bound by instruction issue, not by HBM, so every function call in the column is a hand-counted sequence.
#pragma once
#include <stdint.h>

#include "lf_math.cuh"

namespace lfsoil {

struct Ptrs {
    int64_t n;
    // forcing of this step
    const double *Rain, *SnowMelt, *ETRef, *EWRef, *ESRef, *LAI, *LAITerm;  // LAI, LAITerm: (V,N)
    const uint8_t *frozen;
    // per-pixel parameters
    const double *bX, *PowPref, *UZK, *GwPercStep, *LZK, *LZThreshold, *GwLossStep;
    const double *SoilFraction;  // (V,N)
    const double *DirectRunoffFraction, *WaterFraction;
    // land-use parameters: one pointer per land use (rows may alias, e.g. Irrigated == Rainfed)
    const double *KSat1a[3], *KSat1b[3], *KSat2[3], *InvM1a[3], *InvM1b[3], *InvM2[3];
    const double *WRes1a[3], *WRes1b[3], *WRes2[3], *WS1a[3], *WS1b[3], *WS2[3];
    const double *WWP1a[3], *WWP1b[3], *WFC1a[3], *WFC1b[3], *CropCoef[3], *CropGroup[3];
    // diagnostics-only parameters
    const double *WWP2[3], *WFC2[3], *Depth1a[3], *Depth1b[3], *Depth2[3];
    // state, updated in place
    double *CumInterception, *W1a, *W1b, *W2, *UZ, *DSLR;  // (V,N)
    double *LZ, *CumInterSealed, *LZInflowCUM, *TaCUM, *TaInterceptionCUM, *ESActCUM, *GwLossCUM;
    // outputs consumed by the routing stages
    double *DirectRunoff, *SurfOther, *SurfForest, *GwToChan;
    // fraction-weighted per-column contributions of the pixels that have a deferred column, (V,N); written sparsely.
    // cPref / cInf only exist (and are only touched) with diagnostics.
    double *cTaInt, *cTa, *cES, *cPref, *cInf, *cUZout, *cGwPerc, *cSurf;
    uint8_t *pix_deferred;  // (N): 1 = the per-pixel part is left to k_soil_pixel_flagged
    // deferred columns: six lists of column indices (k = veg*N + pixel), bucketed by sub-step count
    int32_t *list;      // [6 * list_cap]
    int32_t *list_cnt;  // [6]
    int32_t list_cap;
    // scalars
    double DtDay, InvDtDay, AvWaterThreshold, CourantCrit, DrainedFraction, LeafDrainageK, SMaxSealed, TimeSinceStart;
};

// Optional diagnostics (every flux array the reference keeps on self.var); NULL when not requested.
struct Diag {
    double *Interception, *TaInterception, *LeafDrainage, *potential_transpiration, *Ta, *ESAct, *PrefFlow,
        *Infiltration, *AvailableWaterForInfiltration, *SeepTopToSubA, *SeepTopToSubB, *SeepSubToGW, *Theta1a, *Theta1b,
        *Theta2, *Sat1a, *Sat1b, *Sat1, *Sat2, *UZOutflow, *GwPercUZLZ, *RWS, *Theta, *SurfaceRunSoil, *W1;  // (V,N)
    double *RainSnowmelt, *EWaterAct, *InterSealed, *TASealed, *TaInterceptionAll, *TaPixel, *ESActPixel,
        *PrefFlowPixel, *InfiltrationPixel, *ThetaAll, *SeepTopToSubPixelA, *SeepTopToSubPixelB, *SeepSubToGWPixel,
        *Theta1aPixel, *Theta1bPixel, *Theta2Pixel, *UZOutflowPixel, *GwPercUZLZPixel, *GwLossLZ, *LZOutflow, *LZAvInflow,
        *SurfaceRunoff, *TotalRunoff;  // (N)
    int32_t *NoSubS;  // (V,N)
};

using lfm::dmax;
using lfm::dmin;
using lfm::div_nr;
using lfm::MathTab;

// fraction-weighted results of one column that the per-pixel part sums over the fractions
struct Contrib {
    double taint, ta, es, pref, inf, uzout, gwperc, surf;
};

// saturationDegree + unsaturatedConductivity, soilloop.py:360-383
__device__ __forceinline__ double unsat_k(double w, bool pore, double wres, double ws, double ksat, double invm, double m,
                                          const MathTab *MT)
{
    double sat = dmax(dmin(div_nr(w - wres, ws - wres), 1.), 0.);
    sat = pore ? sat : 0.;
    const double t = 1. - lfm::pw_tab<false>(1. - lfm::pw_tab<false>(sat, invm, MT), m, MT);
    return ksat * lfm::sqrt_nr(sat) * (t * t);
}

constexpr int NBUCKET = 6;
__device__ __forceinline__ int bucket_of(int nsub)
{
    // 2-3 -> 0, 4-7 -> 1, 8-15 -> 2, 16-31 -> 3, 32-63 -> 4, 64+ -> 5
    int b = 30 - __clz(nsub);
    return b > 5 ? 5 : b;
}

enum ColumnResult { COL_DONE = 0, COL_QUEUED = 1 };

template <bool DIAG>
__device__ __noinline__ void soil_column_overflow(const Ptrs &P, const Diag &D, const MathTab *MT, int v, int i);

// One soil column (vegetation fraction v of pixel i, k = v*N + i).
// FIRST = true  (k_soil_fused): a column needing more than one Darcy sub-step is queued (COL_QUEUED; nothing written);
//                otherwise the state is written and the contributions are returned in C (COL_DONE).
// FIRST = false (k_soil_veg_deferred): integrates any number of sub-steps, writes state and the c* maps.
template <bool DIAG, bool FIRST>
__device__ __forceinline__ ColumnResult soil_column(const Ptrs &P, const Diag &D, const MathTab *MT, int v, int i, Contrib &C)
{
    // 32-bit pixel index + one 64-bit row offset: the ~50 map accesses of a column then cost one IMAD.WIDE each
    const int64_t N = P.n;
    const int64_t k = (int64_t)v * N + i;
    const double rain = P.Rain[i], etref = P.ETRef[i], ewref = P.EWRef[i];
    const bool frozen = P.frozen[i] != 0;
    const double bX = P.bX[i];
    const double rain_snow = rain + P.SnowMelt[i];
    const double frac = P.SoilFraction[k];
    const double lai = P.LAI[k], laiterm = P.LAITerm[k];
    const double wres1a = P.WRes1a[v][i], wres1b = P.WRes1b[v][i], wres2 = P.WRes2[v][i];
    const double ws1a = P.WS1a[v][i], ws1b = P.WS1b[v][i], ws2 = P.WS2[v][i];
    const double wwp1a = P.WWP1a[v][i], wwp1b = P.WWP1b[v][i], wfc1a = P.WFC1a[v][i], wfc1b = P.WFC1b[v][i];
    double w1a = P.W1a[k], w1b = P.W1b[k], w2 = P.W2[k];
    // ---------------- canopy: interception (soilloop.py:27-70) ----------------
    const double one_minus = 1. - laiterm;
    const double ta_int_max = ewref * one_minus;  // :531-532
    double cum = P.CumInterception[k];
    double smax;
    if (lai <= .1) smax = 0.;
    else if (lai <= 43.3) smax = 0.935 + 0.498 * lai - 0.00575 * (lai * lai);
    else smax = 11.718;
    double interception = 0., ta_int, leafdr;
    if (smax > 0) {
        const double wet = 1. - lfm::exp_neg_tab(div_nr(-0.046 * lai * rain, smax), MT);
        interception = dmin(dmin(smax - cum, smax * wet), rain);
        cum += interception;
    }
    if (cum > 0.) {
        ta_int = dmax(dmin(cum, ta_int_max), 0.);
        cum = dmax(cum - ta_int, 0.);
        leafdr = P.LeafDrainageK * cum;
        cum = dmax(cum - leafdr, 0.);
    } else {
        ta_int = 0.;
        leafdr = 0.;
    }
    // ---------------- canopy: transpiration and soil water stress (:549-627) ----------------
    const double transpir_max = P.CropCoef[v][i] * etref * one_minus;
    const double pot_t = dmax(transpir_max - ta_int, 0.);
    const double cgn = P.CropGroup[v][i];
    const double e_dep = dmin(0.1 * etref * P.InvDtDay, 1.0);
    double p = div_nr(1., 0.76 + 1.5 * e_dep) - 0.10 * (5 - cgn);
    if (cgn <= 2.5) p = p + div_nr(e_dep - 0.6, cgn * (cgn + 3));
    p = dmax(dmin(p, 1.0), 0.);
    const double wfc1 = wfc1a + wfc1b, wwp1 = wwp1a + wwp1b;
    const double wc1 = ((1 - p) * (wfc1 - wwp1)) + wwp1;
    const double wc1a = ((1 - p) * (wfc1a - wwp1a)) + wwp1a;
    const double wc1b = ((1 - p) * (wfc1b - wwp1b)) + wwp1b;
    double w1 = w1a + w1b;
    double rws = (wc1 - wwp1) > 0 ? div_nr(w1 - wwp1, wc1 - wwp1) : 1.;
    rws = dmax(dmin(rws, 1.), 0.);
    double ta = dmin(rws * pot_t, dmax(w1 - wwp1, 0.));
    if (frozen) ta = 0.;
    {
        const double a_free = dmax(w1a - wc1a, 0.), b_free = dmax(w1b - wc1b, 0.);
        double ta1a = dmin(ta, a_free);
        double rest = dmax(ta - ta1a, 0.);
        double ta1b = dmin(rest, b_free);
        rest = dmax(rest - ta1b, 0.);
        const double sa = dmax(w1a - ta1a - wwp1a, 0.), sb = dmax(w1b - ta1b - wwp1b, 0.);
        const double tot = sa + sb;
        const double fa = tot > 0 ? div_nr(sa, tot) : 0., fb = tot > 0 ? div_nr(sb, tot) : 0.;
        ta1a += fa * rest;
        ta1b += fb * rest;
        w1a -= ta1a;
        w1b -= ta1b;
        w1 = w1a + w1b;
    }
    // ---------------- soil column (soilloop.py:105-355) ----------------
    double avail = dmax(rain_snow + leafdr - interception, 0.);  // :131
    double dslr = P.DSLR[k];
    if (avail > P.AvWaterThreshold) dslr = 1;
    else dslr += P.DtDay;  // :137-140
    double esact;
    if (frozen) {
        esact = 0.;
    } else {
        const double esmax = P.ESRef[i] * laiterm;  // :638
        esact = esmax * (lfm::sqrt_nr(dslr) - lfm::sqrt_nr(dslr - 1));
        esact = dmax(dmin(esact, w1 - (wres1a + wres1b)), 0.);
        const double supply1a = w1a - wres1a;
        const double es1a = dmin(esact, supply1a), es1b = dmax(esact - supply1a, 0.);
        w1a = dmax(w1a - es1a, wres1a);
        w1b = dmax(w1b - es1b, wres1b);
    }
    w1 = w1a + w1b;
    const bool pore1a = ws1a != 0, pore1b = ws1b != 0, pore2 = ws2 != 0;  // PoreSpaceNotZero (depth != 0 && WS != 0)
    const double ws1 = ws1a + ws1b;
    const double relsat1 = pore1a ? dmin(div_nr(w1, ws1), 1.0) : 0.0;
    const double satfrac = 1.0 - lfm::pw_tab<true>(1.0 - relsat1, bX, MT);
    const double bX1 = bX + 1;
    const double store_max = div_nr(ws1, bX1);   // StoreMaxPervious, soil.py:363
    const double powinf = div_nr(bX1, bX);       // PowerInfPot, soil.py:361
    const double infpot = frozen ? 0.0 : store_max * lfm::pw_tab<true>(1. - satfrac, powinf, MT) * P.DtDay;
    const double prefflow = lfm::pw_tab<true>(relsat1, P.PowPref[i], MT) * avail;
    avail -= prefflow;
    double infil = dmax(dmin(avail, infpot), 0.);
    {
        const double test = w1a + infil;
        w1a = dmin(ws1a, test);
        w1b += dmax(test - ws1a, 0.);
    }
    const double ks1a = P.KSat1a[v][i], ks1b = P.KSat1b[v][i], ks2 = P.KSat2[v][i];
    const double im1a = P.InvM1a[v][i], im1b = P.InvM1b[v][i], im2 = P.InvM2[v][i];
    const double m1a = div_nr(1.0, im1a), m1b = div_nr(1.0, im1b), m2 = div_nr(1.0, im2);  // GenuM
    double k1a = unsat_k(w1a, pore1a, wres1a, ws1a, ks1a, im1a, m1a, MT);
    double k1b = unsat_k(w1b, pore1b, wres1b, ws1b, ks1b, im1b, m1b, MT);
    double k2 = unsat_k(w2, pore2, wres2, ws2, ks2, im2, m2, MT);
    double av1a = w1a - wres1a, av1b = w1b - wres1b, av2 = w2 - wres2;
    double cap1 = ws1b - w1b, cap2 = ws2 - w2;
    const double cA = av1a == 0 ? 0. : div_nr(k1a * P.DtDay, av1a);
    const double cB = av1b == 0 ? 0. : div_nr(k1b * P.DtDay, av1b);
    const double cG = av2 == 0 ? 0. : div_nr(k2 * P.DtDay, av2);
    const double courant = dmax(dmax(cA, cB), cG);
    const int nsub = (int)dmin(dmax(1., ceil(div_nr(courant, P.CourantCrit))), 2.0e9);
    double seepA, seepB, seepG;
    if (FIRST) {
        // ---- columns that need several sub-steps go to the bucket lists ----
        const unsigned act = __activemask();
        const bool defer = nsub > 1;
        if (__ballot_sync(act, defer)) {
            const int b = defer ? bucket_of(nsub) : -1;
            bool queued = false;
#pragma unroll
            for (int bb = 0; bb < NBUCKET; ++bb) {
                const unsigned mk = __ballot_sync(act, b == bb);
                if (mk == 0) continue;
                const int lane = threadIdx.x & 31;
                const int leader = __ffs(mk) - 1;
                int base = 0;
                if (lane == leader) base = atomicAdd(P.list_cnt + bb, __popc(mk));
                base = __shfl_sync(act, base, leader);
                if (b == bb) {
                    const int slot = base + __popc(mk & ((1u << lane) - 1));
                    if (slot < P.list_cap) P.list[(int64_t)bb * P.list_cap + slot] = (int32_t)k;
                    else soil_column_overflow<DIAG>(P, D, MT, v, i);  // list full: integrate here (never seen in practice)
                    queued = true;
                }
            }
            if (queued) return COL_QUEUED;
        }
        // single sub-step (:237-312 with NoSubS == 1)
        seepA = dmin(k1a * P.DtDay, cap1);
        seepB = dmin(k1b * P.DtDay, cap2);
        seepG = dmin(k2 * P.DtDay, av2);
    } else {
        const double dtsub = div_nr(P.DtDay, (double)nsub);
        seepA = seepB = seepG = 0.;
        double wt1a = w1a, wt1b = w1b, wt2 = w2;
        for (int s = 0; s < nsub; ++s) {
            if (s > 0) {
                k1a = unsat_k(wt1a, pore1a, wres1a, ws1a, ks1a, im1a, m1a, MT);
                k1b = unsat_k(wt1b, pore1b, wres1b, ws1b, ks1b, im1b, m1b, MT);
                k2 = unsat_k(wt2, pore2, wres2, ws2, ks2, im2, m2, MT);
            }
            const double sA = dmin(k1a * dtsub, cap1), sB = dmin(k1b * dtsub, cap2), sG = dmin(k2 * dtsub, av2);
            av1a -= sA;
            av1b += sA - sB;
            av2 += sB - sG;
            wt1a = av1a + wres1a;
            wt1b = av1b + wres1b;
            wt2 = av2 + wres2;
            cap1 = ws1b - wt1b;
            cap2 = ws2 - wt2;
            seepA += sA;
            seepB += sB;
            seepG += sG;
        }
    }
    if (frozen) seepA = seepB = seepG = 0.;
    w1a -= seepA;
    w1b = w1b + seepA - seepB;
    w2 = w2 + seepB - seepG;
    w1 = w1a + w1b;
    infil -= dmax(w1a - ws1a, 0.);
    w1a = dmin(w1a, ws1a);
    // upper zone (:340-354)
    double uz = P.UZ[k];
    double uzout = dmin(P.UZK[i] * uz, uz);
    uz = dmax(uz - uzout, 0.);
    if (v == 2 && P.DrainedFraction > 0) {  // is_irrigated[v] and DrainedFraction > 0 (:115)
        uzout += P.DrainedFraction * seepG;
        uz += (1 - P.DrainedFraction) * seepG + prefflow;
    } else {
        uz += seepG + prefflow;
    }
    const double gwp = dmin(P.GwPercStep[i], uz);
    uz = dmax(uz - gwp, 0.);
    // ---- state ----
    P.CumInterception[k] = cum;
    P.DSLR[k] = dslr;
    P.W1a[k] = w1a;
    P.W1b[k] = w1b;
    P.W2[k] = w2;
    P.UZ[k] = uz;
    // ---- fraction-weighted contributions to the pixel sums (deffraction, Lisflood_initial.py:393-396) ----
    C.taint = frac * ta_int;
    C.ta = frac * ta;
    C.es = frac * esact;
    C.uzout = frac * uzout;
    C.gwperc = frac * gwp;
    C.surf = frac * dmax(avail - infil, 0.);  // SurfaceRunSoil, surface_routing.py:122-126
    if (DIAG) {
        C.pref = frac * prefflow;
        C.inf = frac * infil;
    }
    if (!FIRST) {
        P.cTaInt[k] = C.taint;
        P.cTa[k] = C.ta;
        P.cES[k] = C.es;
        P.cUZout[k] = C.uzout;
        P.cGwPerc[k] = C.gwperc;
        P.cSurf[k] = C.surf;
        if (DIAG) {
            P.cPref[k] = C.pref;
            P.cInf[k] = C.inf;
        }
    }
    if (DIAG) {
        D.Interception[k] = interception;
        D.TaInterception[k] = ta_int;
        D.LeafDrainage[k] = leafdr;
        D.potential_transpiration[k] = pot_t;
        D.Ta[k] = ta;
        D.ESAct[k] = esact;
        D.PrefFlow[k] = prefflow;
        D.Infiltration[k] = infil;
        D.AvailableWaterForInfiltration[k] = avail;
        D.SeepTopToSubA[k] = seepA;
        D.SeepTopToSubB[k] = seepB;
        D.SeepSubToGW[k] = seepG;
        const double d1a = P.Depth1a[v][i], d1b = P.Depth1b[v][i], d2 = P.Depth2[v][i];
        D.Theta1a[k] = (pore1a && d1a != 0) ? w1a / d1a : 0.;
        D.Theta1b[k] = (pore1b && d1b != 0) ? w1b / d1b : 0.;
        D.Theta2[k] = (pore2 && d2 != 0) ? w2 / d2 : 0.;
        D.Sat1a[k] = (w1a - wwp1a) / (wfc1a - wwp1a);
        D.Sat1b[k] = (w1b - wwp1b) / (wfc1b - wwp1b);
        D.Sat1[k] = (w1 - wwp1) / (wfc1 - wwp1);
        D.Sat2[k] = (w2 - P.WWP2[v][i]) / (P.WFC2[v][i] - P.WWP2[v][i]);
        D.UZOutflow[k] = uzout;
        D.GwPercUZLZ[k] = gwp;
        D.RWS[k] = rws;
        D.W1[k] = w1;
        D.SurfaceRunSoil[k] = C.surf;
        D.NoSubS[k] = nsub;
        D.Theta[k] = frac * ((w1a + w1b) + w2) / ((d1a + d1b) + d2);  // soil.py:496-499
    }
    return COL_DONE;
}

template <bool DIAG>
__device__ __noinline__ void soil_column_overflow(const Ptrs &P, const Diag &D, const MathTab *MT, int v, int i)
{
    Contrib C;
    soil_column<DIAG, false>(P, D, MT, v, i, C);
}

// per pixel: open water / sealed soil, totals over the fractions, groundwater, runoff components.
// s*: sums over the three fractions in the reference's order, (c0 + c1) + c2.
template <bool DIAG>
__device__ __forceinline__ void soil_pixel(const Ptrs &P, const Diag &D, int64_t i, double sTaInt, double sTa, double sES,
                                           double sUZout, double sGwPerc, double surfOther, double surfForest,
                                           double sPref, double sInf)
{
    const int64_t N = P.n;
    const double ewref = P.EWRef[i];
    const double rain_snow = P.Rain[i] + P.SnowMelt[i];
    // ---------------- open water and sealed soil (opensealed.py:41-71) ----------------
    const double rsm = dmax(rain_snow, 0.);
    const double ewater = dmax(dmin(ewref, rsm) * 1.0, 0.);
    double cums = P.CumInterSealed[i];
    const double intersealed = dmin(dmax(P.SMaxSealed - cums, 0.), rsm);
    cums += intersealed;
    const double tasealed = dmax(dmin(cums, ewref), 0.);
    cums = dmax(cums - tasealed, 0.);
    P.CumInterSealed[i] = cums;
    const double drf = P.DirectRunoffFraction[i], wf = P.WaterFraction[i];
    const double direct = drf * (rsm - intersealed) + wf * (rsm - ewater);
    // ---------------- per-pixel totals (soil.py:475-486) ----------------
    const double taintall = sTaInt + drf * tasealed;
    const double esactpix = sES + wf * ewater;
    P.TaInterceptionCUM[i] += taintall;
    P.TaCUM[i] += sTa;
    P.ESActCUM[i] += esactpix;
    // ---------------- groundwater (groundwater.py:134-180) ----------------
    double lz = P.LZ[i];
    const double lzout = dmax(dmin(P.LZK[i] * lz, lz - P.LZThreshold[i]), 0.);
    lz -= lzout;
    lz += sGwPerc;
    const double gwloss = dmax(dmin(P.GwLossStep[i], lz), 0.0);
    lz = lz - gwloss;
    P.LZ[i] = lz;
    const double lzcum = dmax(P.LZInflowCUM[i] + (sGwPerc - gwloss), 0.0);
    P.LZInflowCUM[i] = lzcum;
    P.GwLossCUM[i] += gwloss;
    // ---------------- runoff components handed to the routers ----------------
    P.DirectRunoff[i] = direct;
    P.SurfOther[i] = surfOther;
    P.SurfForest[i] = surfForest;
    P.GwToChan[i] = sUZout + lzout;  // UZOutflowPixel + LZOutflowToChannelPixel (surface_routing.py:211)
    if (DIAG) {
        const double f0 = P.SoilFraction[i], f1 = P.SoilFraction[N + i], f2 = P.SoilFraction[2 * N + i];
#define LF_SUM3(arr) ((arr[i] + arr[N + i]) + arr[2 * N + i])
#define LF_WSUM3(arr) ((f0 * arr[i] + f1 * arr[N + i]) + f2 * arr[2 * N + i])
        D.RainSnowmelt[i] = rsm;
        D.EWaterAct[i] = ewater;
        D.InterSealed[i] = intersealed;
        D.TASealed[i] = tasealed;
        D.TaInterceptionAll[i] = taintall;
        D.TaPixel[i] = sTa;
        D.ESActPixel[i] = esactpix;
        D.PrefFlowPixel[i] = sPref;
        D.InfiltrationPixel[i] = sInf;
        const double fsum = (f0 + f1) + f2;
        D.ThetaAll[i] = fsum > 0 ? LF_SUM3(D.Theta) / fsum : 0.;
        D.SeepTopToSubPixelA[i] = LF_WSUM3(D.SeepTopToSubA);
        D.SeepTopToSubPixelB[i] = LF_WSUM3(D.SeepTopToSubB);
        D.SeepSubToGWPixel[i] = LF_WSUM3(D.SeepSubToGW);
        D.Theta1aPixel[i] = LF_WSUM3(D.Theta1a);
        D.Theta1bPixel[i] = LF_WSUM3(D.Theta1b);
        D.Theta2Pixel[i] = LF_WSUM3(D.Theta2);
        D.UZOutflowPixel[i] = sUZout;
        D.GwPercUZLZPixel[i] = sGwPerc;
        D.GwLossLZ[i] = gwloss;
        D.LZOutflow[i] = lzout;
        D.LZAvInflow[i] = (lzcum * P.InvDtDay) / P.TimeSinceStart;
        // SurfaceRunoff = DirectRunoff + sum over land uses (surface_routing.py:128)
        D.SurfaceRunoff[i] = direct + (surfOther + surfForest);
        D.TotalRunoff[i] = (direct + (surfOther + surfForest)) + sUZout + lzout;
#undef LF_WSUM3
#undef LF_SUM3
    }
}

// ---- kernel 1: every (fraction, pixel) column + the per-pixel part of the pixels without deferred columns ----
// With diagnostics every pixel is flagged: the per-pixel diagnostics read per-column maps of all three fractions.
template <bool DIAG, int TILE, int MINB>
__global__ void __launch_bounds__(3 * TILE, MINB) k_soil_fused(const __grid_constant__ Ptrs P, const __grid_constant__ Diag D)
{
    constexpr int NC = DIAG ? 8 : 6;
    __shared__ MathTab s_tab;
    __shared__ double s_c[NC][3][TILE];
    __shared__ int s_def[TILE];
    const int tid = threadIdx.x;
    const int v = tid / TILE, pl = tid - v * TILE;  // warp-uniform v (TILE is a multiple of 32)
    const int64_t i = (int64_t)blockIdx.x * TILE + pl;
    const bool inside = i < P.n;
    lfm::tab_to_shared(&s_tab, tid, 3 * TILE);
    if (tid < TILE) s_def[tid] = DIAG ? 1 : 0;
    __syncthreads();
    Contrib C;
    bool done = false;
    if (inside) {
        done = soil_column<DIAG, true>(P, D, &s_tab, v, (int)i, C) == COL_DONE;
        if (!done) s_def[pl] = 1;
    }
    if (done) {
        s_c[0][v][pl] = C.taint;
        s_c[1][v][pl] = C.ta;
        s_c[2][v][pl] = C.es;
        s_c[3][v][pl] = C.uzout;
        s_c[4][v][pl] = C.gwperc;
        s_c[5][v][pl] = C.surf;
        if (DIAG) {
            s_c[6][v][pl] = C.pref;
            s_c[7][v][pl] = C.inf;
        }
    }
    __syncthreads();
    if (!inside) return;
    const bool pdef = s_def[pl] != 0;
    if (pdef) {
        if (done) {  // park the finished column's contributions for k_soil_pixel_flagged
            const int64_t k = (int64_t)v * P.n + i;
            P.cTaInt[k] = C.taint;
            P.cTa[k] = C.ta;
            P.cES[k] = C.es;
            P.cUZout[k] = C.uzout;
            P.cGwPerc[k] = C.gwperc;
            P.cSurf[k] = C.surf;
            if (DIAG) {
                P.cPref[k] = C.pref;
                P.cInf[k] = C.inf;
            }
        }
        if (v == 0) P.pix_deferred[i] = 1;
        return;
    }
    if (v != 0) return;
    P.pix_deferred[i] = 0;
#define LF_S3(c) ((s_c[c][0][pl] + s_c[c][1][pl]) + s_c[c][2][pl])
    soil_pixel<DIAG>(P, D, i, LF_S3(0), LF_S3(1), LF_S3(2), LF_S3(3), LF_S3(4),
                     s_c[5][0][pl] + s_c[5][2][pl],  // Rainfed + Irrigated (surface_routing.py:145)
                     s_c[5][1][pl], DIAG ? LF_S3(NC - 2) : 0., DIAG ? LF_S3(NC - 1) : 0.);
#undef LF_S3
}

constexpr int SOIL_THREADS = 128;

// ---- kernel 2: the columns of one bucket list ----
template <bool DIAG, int MINB>
__global__ void __launch_bounds__(SOIL_THREADS, MINB) k_soil_veg_deferred(const __grid_constant__ Ptrs P, const __grid_constant__ Diag D, int bucket)
{
    __shared__ MathTab s_tab;
    const int cnt = min(P.list_cnt[bucket], P.list_cap);
    if ((int)(blockIdx.x * SOIL_THREADS) >= cnt) return;  // whole block beyond the list
    lfm::tab_to_shared(&s_tab, threadIdx.x, SOIL_THREADS);
    __syncthreads();
    const int j = blockIdx.x * SOIL_THREADS + threadIdx.x;
    if (j >= cnt) return;
    const int64_t k = P.list[(int64_t)bucket * P.list_cap + j];
    const int v = k >= 2 * P.n ? 2 : (k >= P.n ? 1 : 0);
    Contrib C;
    soil_column<DIAG, false>(P, D, &s_tab, v, (int)(k - (int64_t)v * P.n), C);
}

// ---- kernel 3: per-pixel part of the flagged pixels (those with a deferred column; all pixels with diagnostics) ----
template <bool DIAG>
__global__ void __launch_bounds__(256) k_soil_pixel_flagged(Ptrs P, Diag D)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t N = P.n;
    if (i >= N) return;
    if (!P.pix_deferred[i]) return;
#define LF_SUM3(arr) ((arr[i] + arr[N + i]) + arr[2 * N + i])
    soil_pixel<DIAG>(P, D, i, LF_SUM3(P.cTaInt), LF_SUM3(P.cTa), LF_SUM3(P.cES), LF_SUM3(P.cUZout), LF_SUM3(P.cGwPerc),
                     P.cSurf[i] + P.cSurf[2 * N + i], P.cSurf[N + i], DIAG ? LF_SUM3(P.cPref) : 0.,
                     DIAG ? LF_SUM3(P.cInf) : 0.);
#undef LF_SUM3
}

}  // namespace lfsoil
