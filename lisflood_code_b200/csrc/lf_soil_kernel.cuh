// lf_soil_kernel.cuh -- fused per-cell kernels of the soil / canopy / groundwater stack (device).
//
// ONE fused stage replaces what the reference executes as separate NumPy/Numba passes:
//   soilloop.dynamic_canopy     hydrological_modules/soilloop.py:519-627 (+ kernel :27-70)
//   soilloop.dynamic_soil       hydrological_modules/soilloop.py:630-665 (+ kernel :78-355)
//   opensealed.dynamic          hydrological_modules/opensealed.py:41-71
//   soil.dynamic_perpixel       hydrological_modules/soil.py:471-514 (deffraction: Lisflood_initial.py:393-396)
//   groundwater.dynamic         hydrological_modules/groundwater.py:134-180
//   surface_routing.dynamic     hydrological_modules/surface_routing.py:122-149 (runoff components only)
// There is no neighbour access anywhere in these stages: the kernels are pure streams over
// SoA float64 maps stored in the overland-flow router's position order.
//
// Kernels:
//   k_soil_staged         first pass of the lean (production) build: the input rows of a 64-pixel tile are staged in
//                         shared memory by bulk async copies (TMA, one mbarrier), 3 x 64 threads integrate the three
//                         columns of each pixel, the v == 0 thread finishes the per-pixel part;
//   k_soil_fused          first pass of the diagnostics build (direct global loads, writes every flux map of self.var);
//   k_soil_veg_deferred   the columns that need several Darcy sub-steps, resumed from their mid-column records, one
//                         persistent launch over six bucket lists;
//   k_soil_pixel_flagged  the (late) per-pixel part of the pixels that had a queued column.
//
// Work decomposition of the first pass: a block owns a tile of TILE consecutive pixels and has 3*TILE threads;
// warp-uniform v = thread / TILE is the vegetation fraction, so a thread integrates one (fraction, pixel) soil column.
// The three columns of a pixel run in the same block at the same time: the per-pixel maps (forcing, Xinanjiang b, ...)
// and the land-use parameter rows shared between fractions are fetched from DRAM once, the fraction-weighted column
// results are exchanged through shared memory, and the v == 0 thread of the pixel finishes the per-pixel part (open
// water / sealed soil, sums over fractions, groundwater, runoff components).  DRAM traffic of the first pass is the
// algorithmic 889 B per cell; the first version (separate column and pixel kernels over a (fraction, pixel) grid) moved
// twice that.
//
// Divergence control (the adaptive Darcy sub-stepping, soilloop.py:237-312): the number of sub-steps is
// per column (mean ~1.5, 99th percentile ~20, max ~100), so a warp that simply loops pays the maximum of its
// 32 lanes (~12 on average, measured).  The first pass therefore completes only the columns that need ONE
// sub-step (~99 %) and appends the others, warp-aggregated, to one of six lists bucketed by sub-step count
// (2-3, 4-7, 8-15, 16-31, 32-63, 64+); k_soil_veg_deferred then integrates the lists with one thread per
// column, so lanes of a warp differ by at most 2x in trip count.  A queued column is complete up to the infiltration:
// it leaves a mid-column record in storage it owns (see soil_column) and k_soil_veg_deferred resumes from it -- 23
// gathered values instead of the column's 54, and none of the canopy / evaporation / infiltration arithmetic again.
// A pixel with a queued column is flagged (pix_deferred); its finished columns park their contributions in the
// pixel's contribution record and k_soil_pixel_flagged completes it after the lists have run.  Results do not depend
// on list order.
//
// Arithmetic: float64, unfused multiply-add like the reference (--fmad=false) except inside the library functions
// of lf_math.cuh (table-driven x^y and e^x, Newton division and square root, 3-instruction min/max): the first pass is
// bound by instruction issue (70 % of the issue slots busy), so every function call in the column is a hand-counted
// sequence.
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
    // fraction-weighted per-column contributions of the pixels that have a deferred column; written sparsely.  One record
    // of `cstride` doubles per column, the three columns of a pixel adjacent (record (i*3 + v)): a flagged pixel is summed
    // from 72 contiguous bytes instead of 9 sectors.  Slots: enum CSlot (3 without diagnostics, 8 with).
    double *cbuf;
    int32_t cstride;
    uint8_t *pix_deferred;  // (N): PIX_FLAGGED | overflow bits (see soil_column); 0 = pixel finished by the first pass
    // deferred columns: six lists of column indices (k = veg*N + pixel), bucketed by sub-step count
    int32_t *list;      // [6 * list_cap]
    int32_t *list_cnt;  // [6]
    int32_t list_cap;
    // Dense side records of the queued columns (lean build; nullptr = resume from the scattered mid-column record): the
    // first pass has every input of the resumed half in registers / shared memory and writes it, lane-consecutive, to
    // side[(bucket * NSIDE + field) * list_cap + slot]; k_soil_veg_deferred then reads coalesced records instead of
    // gathering 23 values per column from 23 maps (one 32-64 byte DRAM sector each for 8 useful bytes).
    double *side;
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

enum CSlot { CS_UZOUT, CS_GWPERC, CS_SURF, CS_TAINT, CS_TA, CS_ES, CS_PREF, CS_INF };  // the first 3 without diagnostics
__device__ __forceinline__ double *crec(const Ptrs &P, int v, int64_t i) { return P.cbuf + (i * 3 + v) * P.cstride; }

using lfm::dmax;
using lfm::dmin;
using lfm::div_nr;
using lfm::MathTab;

// fraction-weighted results of one column that the per-pixel part sums over the fractions
struct Contrib {
    double taint, ta, es, pref, inf, uzout, gwperc, surf;
};


// ---------------------------------------------------------------------------------------------------------------
// Input rows of one pixel tile.  The column / pixel code below reads every input through an accessor, so the same
// source serves two feeds: InGlobal (plain global loads: deferred columns, diagnostics build) and InStaged (the tile's
// rows staged in shared memory by bulk async copies, k_soil_staged).
// ---------------------------------------------------------------------------------------------------------------
enum PixRow {  // per-pixel maps (N); the first 9 are read by the columns, the rest by the per-pixel part
    R_Rain, R_SnowMelt, R_ETRef, R_EWRef, R_ESRef, R_bX, R_PowPref, R_UZK, R_GwPercStep,
    R_LZK, R_LZThreshold, R_GwLossStep, R_DRF, R_WF, R_LZ, R_CumInterSealed, R_LZInflowCUM, R_TaCUM, R_TaIntCUM,
    R_ESActCUM, R_GwLossCUM, NPIXROW
};
enum VegRow { V_SoilFraction, V_LAI, V_LAITerm, V_CumInt, V_W1a, V_W1b, V_W2, V_UZ, V_DSLR, NVEGROW };  // (V,N) maps
// land-use parameter rows in three groups that alias differently between land uses (Lisflood_initial.py:371-391):
// A = layers 1a/1b (separate forest maps), B = layer 2 (one map for all land uses), C = crop maps (three maps)
enum LuARow { A_KSat1a, A_KSat1b, A_InvM1a, A_InvM1b, A_WRes1a, A_WRes1b, A_WS1a, A_WS1b, A_WWP1a, A_WWP1b, A_WFC1a,
              A_WFC1b, NLUA };
enum LuBRow { B_KSat2, B_InvM2, B_WRes2, B_WS2, NLUB };
enum LuCRow { C_CropCoef, C_CropGroup, NLUC };
constexpr int MAX_STAGE_ROWS = NPIXROW + 3 * NVEGROW + 3 * (NLUA + NLUB + NLUC);

// Copy plan of k_soil_staged, built on the host per launch (lf_model.cu::soil_stage): shared-memory row r of a tile
// starting at pixel `base` is src[r][base .. base+TILE).  Rows 0..NPIXROW-1 are the per-pixel maps, then NVEGROW rows
// per vegetation fraction, then the land-use groups; a group whose row pointers all equal those of an earlier land use
// is stored once (offA/offB/offC give the first row of the group fraction v reads).
struct Stage {
    const double *src[MAX_STAGE_ROWS];
    int32_t nrows;
    int32_t offA[3], offB[3], offC[3];
    int32_t bulk_ok;  // every src pointer (and the frozen flags) 16-byte aligned for every full tile (needs n even)
};

struct InGlobal {
    const Ptrs &P;
    int v, i;
    int64_t k;
    __device__ __forceinline__ InGlobal(const Ptrs &P_, int v_, int i_) : P(P_), v(v_), i(i_), k((int64_t)v_ * P_.n + i_) {}
#define LF_IN_PIX(name, member) __device__ __forceinline__ double name() const { return P.member[i]; }
#define LF_IN_VEG(name, member) __device__ __forceinline__ double name() const { return P.member[k]; }
#define LF_IN_LU(name, member) __device__ __forceinline__ double name() const { return P.member[v][i]; }
    LF_IN_PIX(Rain, Rain) LF_IN_PIX(SnowMelt, SnowMelt) LF_IN_PIX(ETRef, ETRef) LF_IN_PIX(EWRef, EWRef)
    LF_IN_PIX(ESRef, ESRef) LF_IN_PIX(bX, bX) LF_IN_PIX(PowPref, PowPref) LF_IN_PIX(UZK, UZK)
    LF_IN_PIX(GwPercStep, GwPercStep) LF_IN_PIX(LZK, LZK) LF_IN_PIX(LZThreshold, LZThreshold)
    LF_IN_PIX(GwLossStep, GwLossStep) LF_IN_PIX(DRF, DirectRunoffFraction) LF_IN_PIX(WF, WaterFraction)
    LF_IN_PIX(LZ, LZ) LF_IN_PIX(CumInterSealed, CumInterSealed) LF_IN_PIX(LZInflowCUM, LZInflowCUM)
    LF_IN_PIX(TaCUM, TaCUM) LF_IN_PIX(TaIntCUM, TaInterceptionCUM) LF_IN_PIX(ESActCUM, ESActCUM)
    LF_IN_PIX(GwLossCUM, GwLossCUM)
    LF_IN_VEG(SoilFraction, SoilFraction) LF_IN_VEG(LAI, LAI) LF_IN_VEG(LAITerm, LAITerm)
    LF_IN_VEG(CumInt, CumInterception) LF_IN_VEG(W1a, W1a) LF_IN_VEG(W1b, W1b) LF_IN_VEG(W2, W2) LF_IN_VEG(UZ, UZ)
    LF_IN_VEG(DSLR, DSLR)
    LF_IN_LU(KSat1a, KSat1a) LF_IN_LU(KSat1b, KSat1b) LF_IN_LU(KSat2, KSat2) LF_IN_LU(InvM1a, InvM1a)
    LF_IN_LU(InvM1b, InvM1b) LF_IN_LU(InvM2, InvM2) LF_IN_LU(WRes1a, WRes1a) LF_IN_LU(WRes1b, WRes1b)
    LF_IN_LU(WRes2, WRes2) LF_IN_LU(WS1a, WS1a) LF_IN_LU(WS1b, WS1b) LF_IN_LU(WS2, WS2) LF_IN_LU(WWP1a, WWP1a)
    LF_IN_LU(WWP1b, WWP1b) LF_IN_LU(WFC1a, WFC1a) LF_IN_LU(WFC1b, WFC1b) LF_IN_LU(CropCoef, CropCoef)
    LF_IN_LU(CropGroup, CropGroup)
#undef LF_IN_PIX
#undef LF_IN_VEG
#undef LF_IN_LU
    __device__ __forceinline__ bool frozen() const { return P.frozen[i] != 0; }
};

// rows of the tile in shared memory: row r of local pixel pl is rows[r * TILE + pl]
template <int TILE>
struct InStaged {
    const double *pp, *pv, *pa, *pb, *pc;  // pixel rows, this fraction's (V,N) rows, its land-use groups (all + pl)
    bool fr;
    __device__ __forceinline__ InStaged(const double *rows, const Stage &G, const uint8_t *frozen_row, int v, int pl)
        : pp(rows + pl), pv(rows + (NPIXROW + NVEGROW * v) * TILE + pl), pa(rows + G.offA[v] * TILE + pl),
          pb(rows + G.offB[v] * TILE + pl), pc(rows + G.offC[v] * TILE + pl), fr(frozen_row[pl] != 0) {}
#define LF_IN_PIX(name, row) __device__ __forceinline__ double name() const { return pp[(row) * TILE]; }
#define LF_IN_VEG(name, row) __device__ __forceinline__ double name() const { return pv[(row) * TILE]; }
#define LF_IN_A(name, row) __device__ __forceinline__ double name() const { return pa[(row) * TILE]; }
#define LF_IN_B(name, row) __device__ __forceinline__ double name() const { return pb[(row) * TILE]; }
#define LF_IN_C(name, row) __device__ __forceinline__ double name() const { return pc[(row) * TILE]; }
    LF_IN_PIX(Rain, R_Rain) LF_IN_PIX(SnowMelt, R_SnowMelt) LF_IN_PIX(ETRef, R_ETRef) LF_IN_PIX(EWRef, R_EWRef)
    LF_IN_PIX(ESRef, R_ESRef) LF_IN_PIX(bX, R_bX) LF_IN_PIX(PowPref, R_PowPref) LF_IN_PIX(UZK, R_UZK)
    LF_IN_PIX(GwPercStep, R_GwPercStep) LF_IN_PIX(LZK, R_LZK) LF_IN_PIX(LZThreshold, R_LZThreshold)
    LF_IN_PIX(GwLossStep, R_GwLossStep) LF_IN_PIX(DRF, R_DRF) LF_IN_PIX(WF, R_WF) LF_IN_PIX(LZ, R_LZ)
    LF_IN_PIX(CumInterSealed, R_CumInterSealed) LF_IN_PIX(LZInflowCUM, R_LZInflowCUM) LF_IN_PIX(TaCUM, R_TaCUM)
    LF_IN_PIX(TaIntCUM, R_TaIntCUM) LF_IN_PIX(ESActCUM, R_ESActCUM) LF_IN_PIX(GwLossCUM, R_GwLossCUM)
    LF_IN_VEG(SoilFraction, V_SoilFraction) LF_IN_VEG(LAI, V_LAI) LF_IN_VEG(LAITerm, V_LAITerm)
    LF_IN_VEG(CumInt, V_CumInt) LF_IN_VEG(W1a, V_W1a) LF_IN_VEG(W1b, V_W1b) LF_IN_VEG(W2, V_W2) LF_IN_VEG(UZ, V_UZ)
    LF_IN_VEG(DSLR, V_DSLR)
    LF_IN_A(KSat1a, A_KSat1a) LF_IN_A(KSat1b, A_KSat1b) LF_IN_A(InvM1a, A_InvM1a) LF_IN_A(InvM1b, A_InvM1b)
    LF_IN_A(WRes1a, A_WRes1a) LF_IN_A(WRes1b, A_WRes1b) LF_IN_A(WS1a, A_WS1a) LF_IN_A(WS1b, A_WS1b)
    LF_IN_A(WWP1a, A_WWP1a) LF_IN_A(WWP1b, A_WWP1b) LF_IN_A(WFC1a, A_WFC1a) LF_IN_A(WFC1b, A_WFC1b)
    LF_IN_B(KSat2, B_KSat2) LF_IN_B(InvM2, B_InvM2) LF_IN_B(WRes2, B_WRes2) LF_IN_B(WS2, B_WS2)
    LF_IN_C(CropCoef, C_CropCoef) LF_IN_C(CropGroup, C_CropGroup)
#undef LF_IN_PIX
#undef LF_IN_VEG
#undef LF_IN_A
#undef LF_IN_B
#undef LF_IN_C
    __device__ __forceinline__ bool frozen() const { return fr; }
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
enum SideField { SF_W1A, SF_W1B, SF_W2, SF_AVAIL, SF_INFIL, SF_PREF, SF_WRES1A, SF_WRES1B, SF_WRES2, SF_WS1A, SF_WS1B, SF_WS2,
                 SF_KS1A, SF_KS1B, SF_KS2, SF_IM1A, SF_IM1B, SF_IM2, SF_FRAC, SF_UZ, SF_UZK, SF_GWPERC, SF_FROZEN, NSIDE };

// flags of a pixel in pix_deferred: bit 0 = the per-pixel part is left to k_soil_pixel_flagged (a column of the pixel was
// queued, or diagnostics are on); bits 1..3 = column v found its bucket list full (never seen in practice: each list holds
// a quarter of all columns) and is integrated by k_soil_pixel_flagged itself.  Nothing is called from the first pass: a
// call there costs ~260 bytes of register spills on the common path.
constexpr int PIX_FLAGGED = 1;
__device__ __forceinline__ int pix_overflow_bit(int v) { return 2 << v; }

// ---- second half of a column: apply the seepage, upper zone, state and contributions (soilloop.py:314-354) ----
template <bool DIAG>
__device__ __forceinline__ void column_finish(const Ptrs &P, const Diag &D, int v, int i, int64_t k, bool frozen, double frac,
                                              double w1a, double w1b, double w2, double seepA, double seepB, double seepG,
                                              double avail, double infil, double prefflow, double ws1a, double uz,
                                              double uzk, double gwpercstep, int nsub, Contrib &C)
{
    if (frozen) seepA = seepB = seepG = 0.;
    w1a -= seepA;
    w1b = w1b + seepA - seepB;
    w2 = w2 + seepB - seepG;
    const double w1 = w1a + w1b;
    infil -= dmax(w1a - ws1a, 0.);
    w1a = dmin(w1a, ws1a);
    // upper zone (:340-354)
    double uzout = dmin(uzk * uz, uz);
    uz = dmax(uz - uzout, 0.);
    if (v == 2 && P.DrainedFraction > 0) {  // is_irrigated[v] and DrainedFraction > 0 (:115)
        uzout += P.DrainedFraction * seepG;
        uz += (1 - P.DrainedFraction) * seepG + prefflow;
    } else {
        uz += seepG + prefflow;
    }
    const double gwp = dmin(gwpercstep, uz);
    uz = dmax(uz - gwp, 0.);
    P.W1a[k] = w1a;
    P.W1b[k] = w1b;
    P.W2[k] = w2;
    P.UZ[k] = uz;
    C.uzout = frac * uzout;
    C.gwperc = frac * gwp;
    C.surf = frac * dmax(avail - infil, 0.);  // SurfaceRunSoil, surface_routing.py:122-126
    if (DIAG) {
        C.inf = frac * infil;
        D.Infiltration[k] = infil;
        D.SeepTopToSubA[k] = seepA;
        D.SeepTopToSubB[k] = seepB;
        D.SeepSubToGW[k] = seepG;
        const double d1a = P.Depth1a[v][i], d1b = P.Depth1b[v][i], d2 = P.Depth2[v][i];
        const double wwp1a = P.WWP1a[v][i], wwp1b = P.WWP1b[v][i], wfc1a = P.WFC1a[v][i], wfc1b = P.WFC1b[v][i];
        const double wfc1 = wfc1a + wfc1b, wwp1 = wwp1a + wwp1b;
        D.Theta1a[k] = (P.WS1a[v][i] != 0 && d1a != 0) ? w1a / d1a : 0.;
        D.Theta1b[k] = (P.WS1b[v][i] != 0 && d1b != 0) ? w1b / d1b : 0.;
        D.Theta2[k] = (P.WS2[v][i] != 0 && d2 != 0) ? w2 / d2 : 0.;
        D.Sat1a[k] = (w1a - wwp1a) / (wfc1a - wwp1a);
        D.Sat1b[k] = (w1b - wwp1b) / (wfc1b - wwp1b);
        D.Sat1[k] = (w1 - wwp1) / (wfc1 - wwp1);
        D.Sat2[k] = (w2 - P.WWP2[v][i]) / (P.WFC2[v][i] - P.WWP2[v][i]);
        D.UZOutflow[k] = uzout;
        D.GwPercUZLZ[k] = gwp;
        D.W1[k] = w1;
        D.SurfaceRunSoil[k] = C.surf;
        D.NoSubS[k] = nsub;
        D.Theta[k] = frac * ((w1a + w1b) + w2) / ((d1a + d1b) + d2);  // soil.py:496-499
    }
}

// Courant number of the column -> number of Darcy sub-steps (soilloop.py:218-236)
__device__ __forceinline__ int substeps_of(const Ptrs &P, double k1a, double k1b, double k2, double av1a, double av1b,
                                           double av2)
{
    const double cA = av1a == 0 ? 0. : div_nr(k1a * P.DtDay, av1a);
    const double cB = av1b == 0 ? 0. : div_nr(k1b * P.DtDay, av1b);
    const double cG = av2 == 0 ? 0. : div_nr(k2 * P.DtDay, av2);
    const double courant = dmax(dmax(cA, cB), cG);
    return (int)dmin(dmax(1., ceil(div_nr(courant, P.CourantCrit))), 2.0e9);
}

// ---- first pass over one soil column (vegetation fraction v of pixel i, k = v*N + i) ----
// Canopy, transpiration, evaporation and infiltration (everything before the Darcy seepage) are final after this call
// for EVERY column: CumInterception, DSLR and the contributions taint / ta / es (pref) are written or returned.
//   one sub-step needed (~99 %): seepage, upper zone and state follow at once; contributions in C (COL_DONE);
//   several sub-steps needed  : the column is appended to the bucket list of its sub-step count and its mid-column
//     record is left in storage the column owns -- W1a/W1b/W2[k] = moisture after infiltration; in its contribution
//     record CS_SURF = available water, CS_UZOUT = infiltration, CS_GWPERC = preferential flow (with diagnostics also
//     the final CS_TAINT/TA/ES/PREF) --
//     for soil_column_resume (COL_QUEUED; *flags gets PIX_FLAGGED, plus the overflow bit if the list was full).
template <bool DIAG, class IN>
__device__ __forceinline__ ColumnResult soil_column(const Ptrs &P, const Diag &D, const MathTab *MT, int v, int i, const IN &in,
                                                    Contrib &C, int *flags)
{
    // 32-bit pixel index + one 64-bit row offset: a map access then costs one IMAD.WIDE
    const int64_t N = P.n;
    const int64_t k = (int64_t)v * N + i;
    const double rain = in.Rain(), etref = in.ETRef(), ewref = in.EWRef();
    const bool frozen = in.frozen();
    const double bX = in.bX();
    const double rain_snow = rain + in.SnowMelt();
    const double frac = in.SoilFraction();
    const double lai = in.LAI(), laiterm = in.LAITerm();
    const double wres1a = in.WRes1a(), wres1b = in.WRes1b(), wres2 = in.WRes2();
    const double ws1a = in.WS1a(), ws1b = in.WS1b(), ws2 = in.WS2();
    const double wwp1a = in.WWP1a(), wwp1b = in.WWP1b(), wfc1a = in.WFC1a(), wfc1b = in.WFC1b();
    double w1a = in.W1a(), w1b = in.W1b(), w2 = in.W2();
    // ---------------- canopy: interception (soilloop.py:27-70) ----------------
    const double one_minus = 1. - laiterm;
    const double ta_int_max = ewref * one_minus;  // :531-532
    double cum = in.CumInt();
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
    const double transpir_max = in.CropCoef() * etref * one_minus;
    const double pot_t = dmax(transpir_max - ta_int, 0.);
    const double cgn = in.CropGroup();
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
    double dslr = in.DSLR();
    if (avail > P.AvWaterThreshold) dslr = 1;
    else dslr += P.DtDay;  // :137-140
    double esact;
    if (frozen) {
        esact = 0.;
    } else {
        const double esmax = in.ESRef() * laiterm;  // :638
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
    const double prefflow = lfm::pw_tab<true>(relsat1, in.PowPref(), MT) * avail;
    avail -= prefflow;
    const double infil = dmax(dmin(avail, infpot), 0.);
    {
        const double test = w1a + infil;
        w1a = dmin(ws1a, test);
        w1b += dmax(test - ws1a, 0.);
    }
    // ---- everything up to here is final: canopy state and the contributions that do not depend on the seepage ----
    P.CumInterception[k] = cum;
    P.DSLR[k] = dslr;
    C.taint = frac * ta_int;
    C.ta = frac * ta;
    C.es = frac * esact;
    if (DIAG) {
        C.pref = frac * prefflow;
        D.Interception[k] = interception;
        D.TaInterception[k] = ta_int;
        D.LeafDrainage[k] = leafdr;
        D.potential_transpiration[k] = pot_t;
        D.Ta[k] = ta;
        D.ESAct[k] = esact;
        D.PrefFlow[k] = prefflow;
        D.AvailableWaterForInfiltration[k] = avail;
        D.RWS[k] = rws;
    }
    const double ks1a = in.KSat1a(), ks1b = in.KSat1b(), ks2 = in.KSat2();
    const double im1a = in.InvM1a(), im1b = in.InvM1b(), im2 = in.InvM2();
    const double m1a = div_nr(1.0, im1a), m1b = div_nr(1.0, im1b), m2 = div_nr(1.0, im2);  // GenuM
    const double k1a = unsat_k(w1a, pore1a, wres1a, ws1a, ks1a, im1a, m1a, MT);
    const double k1b = unsat_k(w1b, pore1b, wres1b, ws1b, ks1b, im1b, m1b, MT);
    const double k2 = unsat_k(w2, pore2, wres2, ws2, ks2, im2, m2, MT);
    const double av2 = w2 - wres2;
    const int nsub = substeps_of(P, k1a, k1b, k2, w1a - wres1a, w1b - wres1b, av2);
    // ---- columns that need several sub-steps go to the bucket lists ----
    const unsigned act = __activemask();
    const bool defer = nsub > 1;
    if (__ballot_sync(act, defer)) {
        const int b = defer ? bucket_of(nsub) : -1;
        int my_slot = -1;
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
                if (slot < P.list_cap) {
                    P.list[(int64_t)bb * P.list_cap + slot] = (int32_t)k;
                    atomicOr(flags, PIX_FLAGGED);
                    my_slot = slot;
                } else {
                    atomicOr(flags, PIX_FLAGGED | pix_overflow_bit(v));  // list full: left to k_soil_pixel_flagged
                }
            }
        }
        if (!DIAG && defer && my_slot >= 0 && P.side) {  // dense side record: everything the resumed half reads
            double *sr = P.side + ((int64_t)b * NSIDE) * P.list_cap + my_slot;
            const int64_t st = P.list_cap;
            sr[SF_W1A * st] = w1a;
            sr[SF_W1B * st] = w1b;
            sr[SF_W2 * st] = w2;
            sr[SF_AVAIL * st] = avail;
            sr[SF_INFIL * st] = infil;
            sr[SF_PREF * st] = prefflow;
            sr[SF_WRES1A * st] = wres1a;
            sr[SF_WRES1B * st] = wres1b;
            sr[SF_WRES2 * st] = wres2;
            sr[SF_WS1A * st] = ws1a;
            sr[SF_WS1B * st] = ws1b;
            sr[SF_WS2 * st] = ws2;
            sr[SF_KS1A * st] = ks1a;
            sr[SF_KS1B * st] = ks1b;
            sr[SF_KS2 * st] = ks2;
            sr[SF_IM1A * st] = im1a;
            sr[SF_IM1B * st] = im1b;
            sr[SF_IM2 * st] = im2;
            sr[SF_FRAC * st] = frac;
            sr[SF_UZ * st] = in.UZ();
            sr[SF_UZK * st] = in.UZK();
            sr[SF_GWPERC * st] = in.GwPercStep();
            sr[SF_FROZEN * st] = frozen ? 1.0 : 0.0;
            return COL_QUEUED;
        }
        if (defer) {  // mid-column record
            P.W1a[k] = w1a;
            P.W1b[k] = w1b;
            P.W2[k] = w2;
            double *rec = crec(P, v, i);
            rec[CS_SURF] = avail;
            rec[CS_UZOUT] = infil;
            rec[CS_GWPERC] = prefflow;
            if (DIAG) {  // (without diagnostics the first pass itself sums taint / ta / es of every pixel)
                rec[CS_TAINT] = C.taint;
                rec[CS_TA] = C.ta;
                rec[CS_ES] = C.es;
                rec[CS_PREF] = C.pref;
            }
            return COL_QUEUED;
        }
    }
    // single sub-step (:237-312 with NoSubS == 1)
    const double seepA = dmin(k1a * P.DtDay, ws1b - w1b);
    const double seepB = dmin(k1b * P.DtDay, ws2 - w2);
    const double seepG = dmin(k2 * P.DtDay, av2);
    column_finish<DIAG>(P, D, v, i, k, frozen, frac, w1a, w1b, w2, seepA, seepB, seepG, avail, infil, prefflow, ws1a, in.UZ(),
                        in.UZK(), in.GwPercStep(), 1, C);
    return COL_DONE;
}

// the Darcy sub-steps of a resumed column and its second half (shared by the two resume feeds below)
template <bool DIAG>
__device__ __forceinline__ void resume_core(const Ptrs &P, const Diag &D, const MathTab *MT, int v, int i, int64_t k, bool frozen,
                                            double w1a, double w1b, double w2, double avail, double infil, double prefflow,
                                            double wres1a, double wres1b, double wres2, double ws1a, double ws1b, double ws2,
                                            double ks1a, double ks1b, double ks2, double im1a, double im1b, double im2,
                                            double frac, double uz, double uzk, double gwpercstep)
{
    double *rec = crec(P, v, i);
    const bool pore1a = ws1a != 0, pore1b = ws1b != 0, pore2 = ws2 != 0;
    const double m1a = div_nr(1.0, im1a), m1b = div_nr(1.0, im1b), m2 = div_nr(1.0, im2);  // GenuM
    double k1a = unsat_k(w1a, pore1a, wres1a, ws1a, ks1a, im1a, m1a, MT);
    double k1b = unsat_k(w1b, pore1b, wres1b, ws1b, ks1b, im1b, m1b, MT);
    double k2 = unsat_k(w2, pore2, wres2, ws2, ks2, im2, m2, MT);
    double av1a = w1a - wres1a, av1b = w1b - wres1b, av2 = w2 - wres2;
    double cap1 = ws1b - w1b, cap2 = ws2 - w2;
    const int nsub = substeps_of(P, k1a, k1b, k2, av1a, av1b, av2);
    const double dtsub = div_nr(P.DtDay, (double)nsub);
    double seepA = 0., seepB = 0., seepG = 0.;
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
    Contrib C;
    column_finish<DIAG>(P, D, v, i, k, frozen, frac, w1a, w1b, w2, seepA, seepB, seepG, avail, infil, prefflow, ws1a, uz, uzk,
                        gwpercstep, nsub, C);
    rec[CS_UZOUT] = C.uzout;
    rec[CS_GWPERC] = C.gwperc;
    rec[CS_SURF] = C.surf;
    if (DIAG) rec[CS_INF] = C.inf;
}

// a queued column resumed from its dense side record (bucket b, slot): coalesced reads
template <bool DIAG>
__device__ __forceinline__ void soil_column_resume_side(const Ptrs &P, const Diag &D, const MathTab *MT, int b, int64_t slot,
                                                        int64_t k)
{
    const int v = k >= 2 * P.n ? 2 : (k >= P.n ? 1 : 0);
    const int i = (int)(k - (int64_t)v * P.n);
    const double *sr = P.side + ((int64_t)b * NSIDE) * P.list_cap + slot;
    const int64_t st = P.list_cap;
    resume_core<DIAG>(P, D, MT, v, i, k, sr[SF_FROZEN * st] != 0.0, sr[SF_W1A * st], sr[SF_W1B * st], sr[SF_W2 * st],
                      sr[SF_AVAIL * st], sr[SF_INFIL * st], sr[SF_PREF * st], sr[SF_WRES1A * st], sr[SF_WRES1B * st],
                      sr[SF_WRES2 * st], sr[SF_WS1A * st], sr[SF_WS1B * st], sr[SF_WS2 * st], sr[SF_KS1A * st], sr[SF_KS1B * st],
                      sr[SF_KS2 * st], sr[SF_IM1A * st], sr[SF_IM1B * st], sr[SF_IM2 * st], sr[SF_FRAC * st], sr[SF_UZ * st],
                      sr[SF_UZK * st], sr[SF_GWPERC * st]);
}

// ---- a queued column, resumed from its mid-column record: adaptive Darcy sub-steps (soilloop.py:237-312), then the
// second half.  Writes the state and the remaining contributions (CS_UZOUT, CS_GWPERC, CS_SURF(, CS_INF)) over the record. ----
template <bool DIAG>
__device__ __forceinline__ void soil_column_resume(const Ptrs &P, const Diag &D, const MathTab *MT, int v, int i)
{
    const int64_t k = (int64_t)v * P.n + i;
    const double *rec = crec(P, v, i);
    resume_core<DIAG>(P, D, MT, v, i, k, P.frozen[i] != 0, P.W1a[k], P.W1b[k], P.W2[k], rec[CS_SURF], rec[CS_UZOUT],
                      rec[CS_GWPERC], P.WRes1a[v][i], P.WRes1b[v][i], P.WRes2[v][i], P.WS1a[v][i], P.WS1b[v][i], P.WS2[v][i],
                      P.KSat1a[v][i], P.KSat1b[v][i], P.KSat2[v][i], P.InvM1a[v][i], P.InvM1b[v][i], P.InvM2[v][i],
                      P.SoilFraction[k], P.UZ[k], P.UZK[i], P.GwPercStep[i]);
}

template <bool DIAG>
__device__ __noinline__ void soil_column_overflow(const Ptrs &P, const Diag &D, int v, int i)
{
    soil_column_resume<DIAG>(P, D, &lfm::g_mathtab, v, i);  // tables read in place (rare path)
}

// per pixel: open water / sealed soil, totals over the fractions, groundwater, runoff components.
// s*: sums over the three fractions in the reference's order, (c0 + c1) + c2.
// PART: PIX_EARLY = what is final after the first pass for EVERY pixel (open water / sealed soil, evaporation totals,
// DirectRunoff: they need taint / ta / es only); PIX_LATE = what needs the seepage of all three columns (groundwater,
// SurfOther / SurfForest / GwToChan).  A pixel with a queued column gets PIX_EARLY from the first pass, where its inputs
// sit in shared memory, and only PIX_LATE (15 gathered values instead of 39) from k_soil_pixel_flagged.
enum PixelPart { PIX_ALL, PIX_EARLY, PIX_LATE };
template <bool DIAG, int PART = PIX_ALL, class IN>
__device__ __forceinline__ void soil_pixel(const Ptrs &P, const Diag &D, int64_t i, const IN &in, double sTaInt, double sTa,
                                           double sES, double sUZout, double sGwPerc, double surfOther, double surfForest,
                                           double sPref, double sInf)
{
    static_assert(!DIAG || PART == PIX_ALL, "diagnostics need the whole per-pixel part in one call");
    const int64_t N = P.n;
    double rsm = 0., ewater = 0., intersealed = 0., tasealed = 0., taintall = 0., esactpix = 0., direct = 0.;
    if (PART != PIX_LATE) {
        const double ewref = in.EWRef();
        const double rain_snow = in.Rain() + in.SnowMelt();
        // ---------------- open water and sealed soil (opensealed.py:41-71) ----------------
        rsm = dmax(rain_snow, 0.);
        ewater = dmax(dmin(ewref, rsm) * 1.0, 0.);
        double cums = in.CumInterSealed();
        intersealed = dmin(dmax(P.SMaxSealed - cums, 0.), rsm);
        cums += intersealed;
        tasealed = dmax(dmin(cums, ewref), 0.);
        cums = dmax(cums - tasealed, 0.);
        P.CumInterSealed[i] = cums;
        const double drf = in.DRF(), wf = in.WF();
        direct = drf * (rsm - intersealed) + wf * (rsm - ewater);
        // ---------------- per-pixel totals (soil.py:475-486) ----------------
        taintall = sTaInt + drf * tasealed;
        esactpix = sES + wf * ewater;
        P.TaInterceptionCUM[i] = in.TaIntCUM() + taintall;
        P.TaCUM[i] = in.TaCUM() + sTa;
        P.ESActCUM[i] = in.ESActCUM() + esactpix;
        P.DirectRunoff[i] = direct;
    }
    double lzout = 0., gwloss = 0., lzcum = 0.;
    if (PART != PIX_EARLY) {
        // ---------------- groundwater (groundwater.py:134-180) ----------------
        double lz = in.LZ();
        lzout = dmax(dmin(in.LZK() * lz, lz - in.LZThreshold()), 0.);
        lz -= lzout;
        lz += sGwPerc;
        gwloss = dmax(dmin(in.GwLossStep(), lz), 0.0);
        lz = lz - gwloss;
        P.LZ[i] = lz;
        lzcum = dmax(in.LZInflowCUM() + (sGwPerc - gwloss), 0.0);
        P.LZInflowCUM[i] = lzcum;
        P.GwLossCUM[i] = in.GwLossCUM() + gwloss;
        // ---------------- runoff components handed to the routers ----------------
        P.SurfOther[i] = surfOther;
        P.SurfForest[i] = surfForest;
        P.GwToChan[i] = sUZout + lzout;  // UZOutflowPixel + LZOutflowToChannelPixel (surface_routing.py:211)
    }
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
    static_assert(DIAG, "the lean build runs k_soil_staged (its first pass also does the early per-pixel part)");
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
        done = soil_column<DIAG>(P, D, &s_tab, v, (int)i, InGlobal(P, v, (int)i), C, &s_def[pl]) == COL_DONE;
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
    const int fl = s_def[pl];
    if (fl != 0) {
        if (done) {  // park the finished column's contributions for k_soil_pixel_flagged
            double *rec = crec(P, v, i);
            rec[CS_TAINT] = C.taint;
            rec[CS_TA] = C.ta;
            rec[CS_ES] = C.es;
            rec[CS_UZOUT] = C.uzout;
            rec[CS_GWPERC] = C.gwperc;
            rec[CS_SURF] = C.surf;
            if (DIAG) {
                rec[CS_PREF] = C.pref;
                rec[CS_INF] = C.inf;
            }
        }
        if (v == 0) P.pix_deferred[i] = (uint8_t)fl;
        return;
    }
    if (v != 0) return;
    P.pix_deferred[i] = 0;
#define LF_S3(c) ((s_c[c][0][pl] + s_c[c][1][pl]) + s_c[c][2][pl])
    soil_pixel<DIAG>(P, D, i, InGlobal(P, 0, (int)i), LF_S3(0), LF_S3(1), LF_S3(2), LF_S3(3), LF_S3(4),
                     s_c[5][0][pl] + s_c[5][2][pl],  // Rainfed + Irrigated (surface_routing.py:145)
                     s_c[5][1][pl], DIAG ? LF_S3(NC - 2) : 0., DIAG ? LF_S3(NC - 1) : 0.);
#undef LF_S3
}

constexpr int SOIL_THREADS = 128;

// ---- kernel 2: the queued columns of all bucket lists, one persistent launch ----
// The grid is a fixed number of blocks (lf_model.cu sizes it to the SMs); threads stride over the concatenation of the
// six lists, LONGEST sub-step counts first, so the few columns with 64+ sub-steps start at once and the many short ones
// fill in behind them, and lanes of a warp stay within one bucket (trip counts within 2x).  Replaces six launches whose
// grids had to cover the list capacity (millions of empty blocks) and whose small tails ran one after the other.
template <bool DIAG, int MINB, bool SIDE = false>
__global__ void __launch_bounds__(SOIL_THREADS, MINB) k_soil_veg_deferred(const __grid_constant__ Ptrs P, const __grid_constant__ Diag D)
{
    __shared__ MathTab s_tab;
    lfm::tab_to_shared(&s_tab, threadIdx.x, SOIL_THREADS);
    __syncthreads();
    int64_t end[NBUCKET];  // end[q]: items of the buckets NBUCKET-1 .. NBUCKET-1-q
    int64_t total = 0;
#pragma unroll
    for (int q = 0; q < NBUCKET; ++q) {
        total += min(P.list_cnt[NBUCKET - 1 - q], P.list_cap);
        end[q] = total;
    }
    for (int64_t j = (int64_t)blockIdx.x * SOIL_THREADS + threadIdx.x; j < total; j += (int64_t)gridDim.x * SOIL_THREADS) {
        int q = 0;
        int64_t first = 0;
#pragma unroll
        for (int t = 0; t < NBUCKET - 1; ++t)
            if (j >= end[t]) {
                q = t + 1;
                first = end[t];
            }
        const int64_t k = P.list[(int64_t)(NBUCKET - 1 - q) * P.list_cap + (j - first)];
        if (SIDE) {
            soil_column_resume_side<DIAG>(P, D, &s_tab, NBUCKET - 1 - q, j - first, k);
        } else {
            const int v = k >= 2 * P.n ? 2 : (k >= P.n ? 1 : 0);
            soil_column_resume<DIAG>(P, D, &s_tab, v, (int)(k - (int64_t)v * P.n));
        }
    }
}

// ---- kernel 3: per-pixel part of the flagged pixels (those with a deferred column; all pixels with diagnostics) ----
template <bool DIAG>
__global__ void __launch_bounds__(256, 4) k_soil_pixel_flagged(const __grid_constant__ Ptrs P, const __grid_constant__ Diag D)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t N = P.n;
    if (i >= N) return;
    const int fl = P.pix_deferred[i];
    if (fl == 0) return;
    if (fl & ~PIX_FLAGGED) {  // columns that found their bucket list full
        for (int v = 0; v < 3; ++v)
            if (fl & pix_overflow_bit(v)) soil_column_overflow<DIAG>(P, D, v, (int)i);
    }
    const double *r0 = crec(P, 0, i), *r1 = r0 + P.cstride, *r2 = r1 + P.cstride;
#define LF_SUM3(c) ((r0[c] + r1[c]) + r2[c])
    if (DIAG)
        soil_pixel<DIAG, DIAG ? PIX_ALL : PIX_LATE>(P, D, i, InGlobal(P, 0, (int)i), LF_SUM3(CS_TAINT), LF_SUM3(CS_TA),
                                                    LF_SUM3(CS_ES), LF_SUM3(CS_UZOUT), LF_SUM3(CS_GWPERC),
                                                    r0[CS_SURF] + r2[CS_SURF], r1[CS_SURF], LF_SUM3(CS_PREF), LF_SUM3(CS_INF));
    else
        soil_pixel<false, PIX_LATE>(P, D, i, InGlobal(P, 0, (int)i), 0., 0., 0., LF_SUM3(CS_UZOUT), LF_SUM3(CS_GWPERC),
                                    r0[CS_SURF] + r2[CS_SURF], r1[CS_SURF], 0., 0.);
#undef LF_SUM3
}

// ---------------------------------------------------------------------------------------------------------------
// k_soil_staged: the lean (no diagnostics) first pass with the tile's input rows staged in shared memory.
//
// k_soil_fused spends 61 % of its warp-stall samples waiting for global loads (profiles/r01_soil_fused_ncu.txt): the
// ~54 loads of a column cannot all be in flight with 56-64 registers per thread.  Here a block first pulls every
// input row of its TILE pixels into shared memory -- one `cp.async.bulk` (TMA, 1-D) of TILE*8 bytes per map, issued by
// one lane per warp, all completing on one mbarrier -- and the three columns of each pixel then read their inputs with
// LDS at immediate offsets (no pointer fetch, no 64-bit address arithmetic, ~30 cycles instead of a DRAM round trip).
// The math tables arrive by the same mechanism.  While a block waits for its rows the other resident blocks of the SM
// compute, so the copies overlap the arithmetic without a software pipeline.  The fraction-weighted column results are
// exchanged through the column's own (already consumed) state rows.  Tiles that cannot use bulk copies (the ragged
// last tile, odd N: the 16-byte alignment rule) are staged with ordinary loads by the whole block; results are
// identical.  State and outputs are written straight to global memory as before.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_addr(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(bar)
                 : "memory");
}

// one lane of a converged warp (the compiler then keeps the elected lane's operands in uniform registers)
__device__ __forceinline__ bool elect_one_sync()
{
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}

// dynamic shared memory of k_soil_staged<TILE>: mbarrier (16 B) | rows | math tables | frozen flags | deferred flags
template <int TILE>
__host__ __device__ constexpr size_t staged_smem_bytes(int nrows)
{
    return 16 + (size_t)nrows * TILE * 8 + sizeof(MathTab) + TILE + TILE * 4;
}

template <int TILE, int MINB>
__global__ void __launch_bounds__(3 * TILE, MINB) k_soil_staged(const __grid_constant__ Ptrs P, const __grid_constant__ Stage G, int force_plain)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw);
    double *rows = reinterpret_cast<double *>(smem_raw + 16);
    MathTab *tab = reinterpret_cast<MathTab *>(rows + (size_t)G.nrows * TILE);
    uint8_t *s_frozen = reinterpret_cast<uint8_t *>(tab + 1);
    int *s_def = reinterpret_cast<int *>(s_frozen + TILE);
    const int tid = threadIdx.x;
    const int v = tid / TILE, pl = tid - v * TILE;  // warp-uniform v (TILE is a multiple of 32)
    const int64_t base = (int64_t)blockIdx.x * TILE;
    const int64_t i = base + pl;
    const bool inside = i < P.n;
    const int cnt = (int)((P.n - base) < (int64_t)TILE ? (P.n - base) : (int64_t)TILE);
    const bool bulk = G.bulk_ok && cnt == TILE && !force_plain;  // block-uniform
    if (tid < TILE) s_def[tid] = 0;
    if (bulk) {
        const uint32_t bar_a = smem_addr(bar);
        if (tid == 0) {
            const uint32_t total = (uint32_t)(G.nrows * TILE * 8 + sizeof(MathTab) + TILE);
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a) : "memory");
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(total) : "memory");
        }
        __syncthreads();
        // one lane per warp issues its share of the row copies; the warp index is made provably warp-uniform so that
        // row index, addresses and the copy itself stay on the uniform datapath (4 instructions per copy instead of 11)
        const int w = __shfl_sync(0xffffffffu, tid >> 5, 0);
        if (elect_one_sync()) {
            constexpr int NW = 3 * TILE / 32;
            if (w == 0) {
                bulk_g2s(tab, &lfm::g_mathtab, sizeof(MathTab), bar_a);
                bulk_g2s(s_frozen, P.frozen + base, TILE, bar_a);
            }
            for (int r = w; r < G.nrows; r += NW) bulk_g2s(rows + (size_t)r * TILE, G.src[r] + base, TILE * 8, bar_a);
        }
        // every thread waits for phase 0 of the barrier (the copies' bytes); bounded so that a bug traps instead of hanging
        // (try_wait suspends the thread for up to the time hint, so the loop body runs a handful of times)
        uint32_t ok = 0;
        for (int spins = 0; !ok; ++spins) {
            asm volatile(
                "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0, 20000;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                : "=r"(ok)
                : "r"(bar_a)
                : "memory");
            if (spins > 400000) __trap();
        }
    } else {
        lfm::tab_to_shared(tab, tid, 3 * TILE);
        for (int idx = tid; idx < G.nrows * TILE; idx += 3 * TILE) {
            const int r = idx / TILE, p = idx - r * TILE;
            if (p < cnt) rows[idx] = G.src[r][base + p];
        }
        if (tid < cnt) s_frozen[tid] = P.frozen[base + tid];
        __syncthreads();
    }
    Contrib C;
    bool done = false;
    double *mine = rows + (size_t)(NPIXROW + NVEGROW * v) * TILE + pl;  // this column's (V,N) rows
    if (inside) {
        const InStaged<TILE> in(rows, G, s_frozen, v, pl);
        done = soil_column<false>(P, Diag(), tab, v, (int)i, in, C, &s_def[pl]) == COL_DONE;
    }
    // the column's own state rows are consumed: they carry its contributions to the per-pixel part
    if (inside) {  // final for every column, queued or not
        mine[0 * TILE] = C.taint;
        mine[1 * TILE] = C.ta;
        mine[2 * TILE] = C.es;
    }
    if (done) {
        mine[3 * TILE] = C.uzout;
        mine[4 * TILE] = C.gwperc;
        mine[5 * TILE] = C.surf;
    }
    __syncthreads();
    if (!inside) return;
    const int fl = s_def[pl];
    const double *c0 = rows + (size_t)NPIXROW * TILE + pl, *c1 = c0 + NVEGROW * TILE, *c2 = c1 + NVEGROW * TILE;
#define LF_S3(c) ((c0[(c) * TILE] + c1[(c) * TILE]) + c2[(c) * TILE])
    if (fl != 0) {
        if (done) {  // park what the late per-pixel part needs from a finished column
            double *rec = crec(P, v, i);
            rec[CS_UZOUT] = C.uzout;
            rec[CS_GWPERC] = C.gwperc;
            rec[CS_SURF] = C.surf;
        }
        if (v == 0) {  // the early per-pixel part does not wait for the queued columns
            P.pix_deferred[i] = (uint8_t)fl;
            const InStaged<TILE> in(rows, G, s_frozen, 0, pl);
            soil_pixel<false, PIX_EARLY>(P, Diag(), i, in, LF_S3(0), LF_S3(1), LF_S3(2), 0., 0., 0., 0., 0., 0.);
        }
        return;
    }
    if (v != 0) return;
    P.pix_deferred[i] = 0;
    const InStaged<TILE> in(rows, G, s_frozen, 0, pl);
    soil_pixel<false>(P, Diag(), i, in, LF_S3(0), LF_S3(1), LF_S3(2), LF_S3(3), LF_S3(4),
                      c0[5 * TILE] + c2[5 * TILE],  // Rainfed + Irrigated (surface_routing.py:145)
                      c1[5 * TILE], 0., 0.);
#undef LF_S3
}

}  // namespace lfsoil
