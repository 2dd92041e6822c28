// lf_soil_kernel.cuh -- fused per-cell kernel of the soil / canopy / groundwater stack (device).
//
// Two fused kernels replace the stages the reference executes as separate NumPy/Numba passes
// (k_soil_veg: one thread per (vegetation fraction, pixel); k_soil_pixel: one thread per pixel):
//   soilloop.dynamic_canopy     hydrological_modules/soilloop.py:519-627 (+ kernel :27-70)
//   soilloop.dynamic_soil       hydrological_modules/soilloop.py:630-665 (+ kernel :78-355)
//   opensealed.dynamic          hydrological_modules/opensealed.py:41-71
//   soil.dynamic_perpixel       hydrological_modules/soil.py:471-514 (deffraction: Lisflood_initial.py:393-396)
//   groundwater.dynamic         hydrological_modules/groundwater.py:134-180
//   surface_routing.dynamic     hydrological_modules/surface_routing.py:122-149 (runoff components only)
// There is no neighbour access anywhere in these stages: the kernels are pure streams over
// SoA float64 maps stored in the overland-flow router's position order.
//
// Divergence control (the adaptive Darcy sub-stepping, soilloop.py:237-312): the number of sub-steps is
// per pixel (mean ~1.5, 99th percentile ~20, max ~100), so a warp that simply loops pays the maximum of its
// 32 lanes (~12 on average, measured).  k_soil_veg therefore completes only the columns that need ONE
// sub-step (94 %) and appends the others, warp-aggregated, to one of six lists bucketed by sub-step count
// (2-3, 4-7, 8-15, 16-31, 32-63, 64+); k_soil_veg_deferred then integrates each list with one thread per
// column, so lanes of a warp differ by at most 2x in trip count.  Deferred columns recompute their (cheap)
// prologue instead of spilling ~30 doubles of state.  Results do not depend on list order.
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
    // fraction-weighted per-column contributions handed from k_soil_veg to k_soil_pixel, (V,N)
    double *cTaInt, *cTa, *cES, *cPref, *cInf, *cUZout, *cGwPerc, *cSurf;
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

__device__ __forceinline__ double pw(double x, double y) { return lfm::pw(x, y); }

// saturationDegree + unsaturatedConductivity, soilloop.py:360-383
__device__ __forceinline__ double unsat_k(double w, bool pore, double wres, double ws, double ksat, double invm, double m)
{
    double sat = pore ? fmax(fmin(lfm::div_nr(w - wres, ws - wres), 1.), 0.) : 0.;
    double t = 1. - pw(1. - pw(sat, invm), m);
    return ksat * sqrt(sat) * (t * t);
}

constexpr int NBUCKET = 6;
__device__ __forceinline__ int bucket_of(int nsub)
{
    // 2-3 -> 0, 4-7 -> 1, 8-15 -> 2, 16-31 -> 3, 32-63 -> 4, 64+ -> 5
    int b = 30 - __clz(nsub);
    return b > 5 ? 5 : b;
}

// One soil column (vegetation fraction v of pixel i, k = v*N + i).  `may_defer`: columns needing more than one
// Darcy sub-step are queued instead of integrated (first pass).
template <bool DIAG>
__device__ __forceinline__ void soil_column(const Ptrs &P, const Diag &D, int v, int i, bool may_defer)
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
        interception = fmin(fmin(smax - cum, smax * (1. - exp(-0.046 * lai * rain / smax))), rain);
        cum += interception;
    }
    if (cum > 0.) {
        ta_int = fmax(fmin(cum, ta_int_max), 0.);
        cum = fmax(cum - ta_int, 0.);
        leafdr = P.LeafDrainageK * cum;
        cum = fmax(cum - leafdr, 0.);
    } else {
        ta_int = 0.;
        leafdr = 0.;
    }
    // ---------------- canopy: transpiration and soil water stress (:549-627) ----------------
    const double transpir_max = P.CropCoef[v][i] * etref * one_minus;
    const double pot_t = fmax(transpir_max - ta_int, 0.);
    const double cgn = P.CropGroup[v][i];
    const double e_dep = fmin(0.1 * etref * P.InvDtDay, 1.0);
    double p = 1 / (0.76 + 1.5 * e_dep) - 0.10 * (5 - cgn);
    if (cgn <= 2.5) p = p + (e_dep - 0.6) / (cgn * (cgn + 3));
    p = fmax(fmin(p, 1.0), 0.);
    const double wfc1 = wfc1a + wfc1b, wwp1 = wwp1a + wwp1b;
    const double wc1 = ((1 - p) * (wfc1 - wwp1)) + wwp1;
    const double wc1a = ((1 - p) * (wfc1a - wwp1a)) + wwp1a;
    const double wc1b = ((1 - p) * (wfc1b - wwp1b)) + wwp1b;
    double w1 = w1a + w1b;
    double rws = (wc1 - wwp1) > 0 ? lfm::div_nr(w1 - wwp1, wc1 - wwp1) : 1.;
    rws = fmax(fmin(rws, 1.), 0.);
    double ta = fmin(rws * pot_t, fmax(w1 - wwp1, 0.));
    if (frozen) ta = 0.;
    {
        const double a_free = fmax(w1a - wc1a, 0.), b_free = fmax(w1b - wc1b, 0.);
        double ta1a = fmin(ta, a_free);
        double rest = fmax(ta - ta1a, 0.);
        double ta1b = fmin(rest, b_free);
        rest = fmax(rest - ta1b, 0.);
        const double sa = fmax(w1a - ta1a - wwp1a, 0.), sb = fmax(w1b - ta1b - wwp1b, 0.);
        const double tot = sa + sb;
        const double fa = tot > 0 ? lfm::div_nr(sa, tot) : 0., fb = tot > 0 ? lfm::div_nr(sb, tot) : 0.;
        ta1a += fa * rest;
        ta1b += fb * rest;
        w1a -= ta1a;
        w1b -= ta1b;
        w1 = w1a + w1b;
    }
    // ---------------- soil column (soilloop.py:105-355) ----------------
    double avail = fmax(rain_snow + leafdr - interception, 0.);  // :131
    double dslr = P.DSLR[k];
    if (avail > P.AvWaterThreshold) dslr = 1;
    else dslr += P.DtDay;  // :137-140
    double esact;
    if (frozen) {
        esact = 0.;
    } else {
        const double esmax = P.ESRef[i] * laiterm;  // :638
        esact = esmax * (sqrt(dslr) - sqrt(dslr - 1));
        esact = fmax(fmin(esact, w1 - (wres1a + wres1b)), 0.);
        const double supply1a = w1a - wres1a;
        const double es1a = fmin(esact, supply1a), es1b = fmax(esact - supply1a, 0.);
        w1a = fmax(w1a - es1a, wres1a);
        w1b = fmax(w1b - es1b, wres1b);
    }
    w1 = w1a + w1b;
    const bool pore1a = ws1a != 0, pore1b = ws1b != 0, pore2 = ws2 != 0;  // PoreSpaceNotZero (depth != 0 && WS != 0)
    const double ws1 = ws1a + ws1b;
    const double relsat1 = pore1a ? fmin(lfm::div_nr(w1, ws1), 1.0) : 0.0;
    const double satfrac = 1.0 - pw(1.0 - relsat1, bX);
    const double store_max = ws1 / (bX + 1);      // StoreMaxPervious, soil.py:363
    const double powinf = (bX + 1) / bX;           // PowerInfPot, soil.py:361
    const double infpot = frozen ? 0.0 : store_max * pw(1. - satfrac, powinf) * P.DtDay;
    const double prefflow = pw(relsat1, P.PowPref[i]) * avail;
    avail -= prefflow;
    double infil = fmax(fmin(avail, infpot), 0.);
    {
        const double test = w1a + infil;
        w1a = fmin(ws1a, test);
        w1b += fmax(test - ws1a, 0.);
    }
    const double ks1a = P.KSat1a[v][i], ks1b = P.KSat1b[v][i], ks2 = P.KSat2[v][i];
    const double im1a = P.InvM1a[v][i], im1b = P.InvM1b[v][i], im2 = P.InvM2[v][i];
    const double m1a = lfm::div_nr(1.0, im1a), m1b = lfm::div_nr(1.0, im1b), m2 = lfm::div_nr(1.0, im2);  // GenuM
    double k1a = unsat_k(w1a, pore1a, wres1a, ws1a, ks1a, im1a, m1a);
    double k1b = unsat_k(w1b, pore1b, wres1b, ws1b, ks1b, im1b, m1b);
    double k2 = unsat_k(w2, pore2, wres2, ws2, ks2, im2, m2);
    double av1a = w1a - wres1a, av1b = w1b - wres1b, av2 = w2 - wres2;
    double cap1 = ws1b - w1b, cap2 = ws2 - w2;
    const double cA = av1a == 0 ? 0. : lfm::div_nr(k1a * P.DtDay, av1a);
    const double cB = av1b == 0 ? 0. : lfm::div_nr(k1b * P.DtDay, av1b);
    const double cG = av2 == 0 ? 0. : lfm::div_nr(k2 * P.DtDay, av2);
    const double courant = fmax(fmax(cA, cB), cG);
    const int nsub = (int)fmin(fmax(1., ceil(courant / P.CourantCrit)), 2.0e9);
    // ---- columns that need several sub-steps go to the bucket lists (first pass only) ----
    if (may_defer) {
        const unsigned act = __activemask();
        const bool defer = nsub > 1;
        const int b = defer ? bucket_of(nsub) : -1;
        bool queued = false;
#pragma unroll
        for (int bb = 0; bb < NBUCKET; ++bb) {
            const unsigned m = __ballot_sync(act, b == bb);
            if (m == 0) continue;
            const int lane = threadIdx.x & 31;
            const int leader = __ffs(m) - 1;
            int base = 0;
            if (lane == leader) base = atomicAdd(P.list_cnt + bb, __popc(m));
            base = __shfl_sync(act, base, leader);
            if (b == bb) {
                const int slot = base + __popc(m & ((1u << lane) - 1));
                if (slot < P.list_cap) {
                    P.list[(int64_t)bb * P.list_cap + slot] = (int32_t)k;
                    queued = true;
                }  // list full: integrate here
            }
        }
        if (queued) return;
    }
    const double dtsub = P.DtDay / (double)nsub;
    double seepA = 0., seepB = 0., seepG = 0.;
    {
        double wt1a = w1a, wt1b = w1b, wt2 = w2;
        for (int s = 0; s < nsub; ++s) {
            if (s > 0) {
                k1a = unsat_k(wt1a, pore1a, wres1a, ws1a, ks1a, im1a, m1a);
                k1b = unsat_k(wt1b, pore1b, wres1b, ws1b, ks1b, im1b, m1b);
                k2 = unsat_k(wt2, pore2, wres2, ws2, ks2, im2, m2);
            }
            const double sA = fmin(k1a * dtsub, cap1), sB = fmin(k1b * dtsub, cap2), sG = fmin(k2 * dtsub, av2);
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
    infil -= fmax(w1a - ws1a, 0.);
    w1a = fmin(w1a, ws1a);
    // upper zone (:340-354)
    double uz = P.UZ[k];
    double uzout = fmin(P.UZK[i] * uz, uz);
    uz = fmax(uz - uzout, 0.);
    if (v == 2 && P.DrainedFraction > 0) {  // is_irrigated[v] and DrainedFraction > 0 (:115)
        uzout += P.DrainedFraction * seepG;
        uz += (1 - P.DrainedFraction) * seepG + prefflow;
    } else {
        uz += seepG + prefflow;
    }
    const double gwp = fmin(P.GwPercStep[i], uz);
    uz = fmax(uz - gwp, 0.);
    // ---- state ----
    P.CumInterception[k] = cum;
    P.DSLR[k] = dslr;
    P.W1a[k] = w1a;
    P.W1b[k] = w1b;
    P.W2[k] = w2;
    P.UZ[k] = uz;
    // ---- fraction-weighted contributions to the pixel sums (deffraction, Lisflood_initial.py:393-396) ----
    P.cTaInt[k] = frac * ta_int;
    P.cTa[k] = frac * ta;
    P.cES[k] = frac * esact;
    P.cPref[k] = frac * prefflow;
    P.cInf[k] = frac * infil;
    P.cUZout[k] = frac * uzout;
    P.cGwPerc[k] = frac * gwp;
    P.cSurf[k] = frac * fmax(avail - infil, 0.);  // SurfaceRunSoil, surface_routing.py:122-126
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
        D.SurfaceRunSoil[k] = frac * fmax(avail - infil, 0.);
        D.NoSubS[k] = nsub;
        D.Theta[k] = frac * ((w1a + w1b) + w2) / ((d1a + d1b) + d2);  // soil.py:496-499
    }
}

constexpr int SOIL_THREADS = 128;

// first pass: every (vegetation fraction, pixel) column
template <bool DIAG, int MINB>
__global__ void __launch_bounds__(SOIL_THREADS, MINB) k_soil_veg(Ptrs P, Diag D)
{
    const int64_t i = (int64_t)blockIdx.x * SOIL_THREADS + threadIdx.x;   // blockIdx.y = vegetation fraction
    if (i >= P.n) return;
    soil_column<DIAG>(P, D, (int)blockIdx.y, (int)i, true);
}
// second pass: the columns of one bucket list
template <bool DIAG, int MINB>
__global__ void __launch_bounds__(SOIL_THREADS, MINB) k_soil_veg_deferred(Ptrs P, Diag D, int bucket)
{
    const int j = blockIdx.x * SOIL_THREADS + threadIdx.x;
    const int cnt = min(P.list_cnt[bucket], P.list_cap);
    if (j >= cnt) return;
    const int64_t k = P.list[(int64_t)bucket * P.list_cap + j];
    const int v = k >= 2 * P.n ? 2 : (k >= P.n ? 1 : 0);
    soil_column<DIAG>(P, D, v, (int)(k - (int64_t)v * P.n), false);
}

// per pixel: sums over the fractions, open water / sealed soil, groundwater, runoff components
template <bool DIAG>
__global__ void __launch_bounds__(256) k_soil_pixel(Ptrs P, Diag D)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t N = P.n;
    if (i >= N) return;
#define LF_SUM3(arr) ((arr[i] + arr[N + i]) + arr[2 * N + i])
    const double sTaInt = LF_SUM3(P.cTaInt), sTa = LF_SUM3(P.cTa), sES = LF_SUM3(P.cES), sUZout = LF_SUM3(P.cUZout),
                 sGwPerc = LF_SUM3(P.cGwPerc);
    const double surfOther = P.cSurf[i] + P.cSurf[2 * N + i];  // Rainfed + Irrigated (surface_routing.py:145)
    const double surfForest = P.cSurf[N + i];
    const double ewref = P.EWRef[i];
    const double rain_snow = P.Rain[i] + P.SnowMelt[i];
    // ---------------- open water and sealed soil (opensealed.py:41-71) ----------------
    const double rsm = fmax(rain_snow, 0.);
    const double ewater = fmax(fmin(ewref, rsm) * 1.0, 0.);
    double cums = P.CumInterSealed[i];
    const double intersealed = fmin(fmax(P.SMaxSealed - cums, 0.), rsm);
    cums += intersealed;
    const double tasealed = fmax(fmin(cums, ewref), 0.);
    cums = fmax(cums - tasealed, 0.);
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
    const double lzout = fmax(fmin(P.LZK[i] * lz, lz - P.LZThreshold[i]), 0.);
    lz -= lzout;
    lz += sGwPerc;
    const double gwloss = fmax(fmin(P.GwLossStep[i], lz), 0.0);
    lz = lz - gwloss;
    P.LZ[i] = lz;
    const double lzcum = fmax(P.LZInflowCUM[i] + (sGwPerc - gwloss), 0.0);
    P.LZInflowCUM[i] = lzcum;
    P.GwLossCUM[i] += gwloss;
    // ---------------- runoff components handed to the routers ----------------
    P.DirectRunoff[i] = direct;
    P.SurfOther[i] = surfOther;
    P.SurfForest[i] = surfForest;
    P.GwToChan[i] = sUZout + lzout;  // UZOutflowPixel + LZOutflowToChannelPixel (surface_routing.py:211)
    if (DIAG) {
        const double f0 = P.SoilFraction[i], f1 = P.SoilFraction[N + i], f2 = P.SoilFraction[2 * N + i];
#define LF_WSUM3(arr) ((f0 * arr[i] + f1 * arr[N + i]) + f2 * arr[2 * N + i])
        D.RainSnowmelt[i] = rsm;
        D.EWaterAct[i] = ewater;
        D.InterSealed[i] = intersealed;
        D.TASealed[i] = tasealed;
        D.TaInterceptionAll[i] = taintall;
        D.TaPixel[i] = sTa;
        D.ESActPixel[i] = esactpix;
        D.PrefFlowPixel[i] = LF_SUM3(P.cPref);
        D.InfiltrationPixel[i] = LF_SUM3(P.cInf);
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
    }
#undef LF_SUM3
}

}  // namespace lfsoil
