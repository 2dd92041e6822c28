// lf_model.cu -- the full hot-path time step on device-resident state (C ABI: lf_model_*).
//
// Stages per model step (reference call order, Lisflood_dynamic.py:114-229):
//   1. k_soil_staged (k_soil_fused with diagnostics) / k_soil_veg_deferred / k_soil_pixel_flagged   fused canopy +
//                         soil column per (fraction, pixel) with open/sealed + per-pixel sums + groundwater per pixel
//                         (lf_soil_kernel.cuh).
//   2. k_of_level/k_of_post  the three overland-flow routers (Other, Forest, Direct) on LddToChan,
//                         solved together in one level sweep (three independent Newton solves per thread),
//                         then OFToChanM3 / ToChanM3RunoffDt written in channel order.
//   3. k_chan_diagonal    NoRoutSteps channel sub-steps (routing.dynamic) as ONE space-time wavefront:
//                         side-flow assembly, kinematic solve(s), volume / discharge post-processing and
//                         sumDisDay accumulation fused per (pixel, sub-step);
//      k_chan_isolated_ws pixels without any link (the ~90 % non-channel pixels of LddKinematic) advance
//                         through all sub-steps in registers; a chunk queue drained by an early launch that
//                         shares the SMs with the soil stage and a late one in the channel stage;
//      k_chan_post        ChanM3, TotalCrossSectionArea, ChanQAvg, sumDis, DischargeM3Out
//                         (Lisflood_dynamic.py:194-229).
// Storage: soil and overland maps live in the overland router's position order, channel maps in the
// channel router's position order; lf_model_set/get translate from/to the reference's
// compressed order.
#include <stdlib.h>
#include <string.h>

#include <map>
#include <memory>
#include <string>
#include <vector>

#include "lf_common.cuh"
#include "lf_kw_solve.cuh"
#include "lf_soil_kernel.cuh"
#include "lf_xchg.cuh"

extern "C" int lf_ldd_build(const double *, const uint8_t *, int64_t, int64_t, lf_graph **);
extern "C" void lf_graph_destroy(lf_graph *);
extern "C" int lf_xchg_begin(lf_xchg *, int32_t *);
extern "C" int lf_xchg_end(lf_xchg *);
extern "C" int lf_xchg_peer_base(lf_xchg *, int32_t, uint64_t *);

namespace {

enum Order { SOIL = 0, CHAN = 1 };

struct Field {
    lf::DevBuf<double> buf;
    int rows = 1;
    Order order = SOIL;
    bool landuse = false;      // (landuse, pixel) parameter: rows may be shared
    int row_index[3] = {0, 1, 2};  // storage row of land use r
    bool diag = false;
    bool as_z = false;  // discharge map stored as z = Q^(1/5) (3/5 mode)
};

struct ChanPtrs {
    int32_t n;
    const int32_t *lev, *cfirst, *cend;   // cend: restricted graphs only (lf_graph_restrict), else children end at cfirst[i+1]
    lfx::View X;                           // LDD-cut exchange (multi-GPU), X.xslot == nullptr otherwise
    double *Qk, *Qr0, *Qr1, *M3, *sumDis, *ChanQ;
    const double *a, *L, *alpha, *sideDt;
    const uint8_t *isChan;
    // split routing
    double *Q2k, *Q2r0, *Q2r1, *M32, *CS2A, *S1, *sumNotLast;
    const double *a2, *alpha2, *QLimit, *M3Limit, *C2M3Start, *C2QStart, *z2floor;
    double InvDtRouting, DtRouting;
    int S, split;
    lfkw::Params P;
    // structures in the sub-step loop (reservoir.py:173-322, lakes.py:199-297): sid[i] = +j+1 reservoir j, -(j+1) lake j,
    // 0 none; feeds[i] != 0: the pixel drains into a structure (its ChanQ of every sub-step is kept in CQ[parity])
    const int32_t *sid;
    const uint8_t *feeds;
    double *CQ0, *CQ1;
    int gs0;                 // global index of sub-step 0 of this model step (parity of the CQ buffers)
    struct Res {
        double *Storage, *Fill, *OutM3;
        const double *Total, *Cons, *Norm, *NormFlood, *Flood, *MinOut, *NormOut, *NonDam, *DeltaO, *DeltaLN, *DeltaNFL;
    } res;
    struct Lake {
        double *Storage, *Outflow, *InflowOld, *Balance, *Level, *OutM3;
        const double *Area, *Factor, *FactorSqr;
    } lake;
};

// reservoir.dynamic_inloop, reservoir.py:173-322 (one reservoir, one routing sub-step): returns the outflow volume [m3]
__device__ __forceinline__ double reservoir_substep(const ChanPtrs::Res &R, int j, double inflow, double DtRouting)
{
    const double inv_day = 1 / 86400.0;
    const double total = R.Total[j], cons = R.Cons[j], flood = R.Flood[j], nflood = R.NormFlood[j];
    const double omin = R.MinOut[j], onorm = R.NormOut[j], ondam = R.NonDam[j];
    double storage = R.Storage[j] + inflow * DtRouting;                   // :196-200
    const double fill = storage / total;                                    // :202
    const double o1 = fmin(omin, storage * inv_day);                        // :205
    const double o2 = omin + R.DeltaO[j] * (fill - 2 * cons) / R.DeltaLN[j];   // :209
    const double o3a = onorm;
    const double o3b = onorm + ((fill - nflood) / R.DeltaNFL[j]) * (ondam - onorm);   // :216-218
    double temp = fmin(ondam, fmax(inflow * 1.2, onorm));                   // :222
    const double o4 = fmax((fill - flood - 0.01) * total * inv_day, temp);  // :223
    double out = o1;
    out = fill > 2 * cons ? o2 : out;                                       // :232-240
    out = fill > R.Norm[j] ? o3a : out;
    out = fill > nflood ? o3b : out;
    out = fill > flood ? o4 : out;
    temp = fmin(out, fmax(inflow, onorm));                                  // :243
    out = (out > 1.2 * inflow && out > onorm && fill < flood) ? temp : out; // :245-249
    double out_m3 = out * DtRouting;                                        // :252
    out_m3 = fmin(out_m3, storage);                                         // :258
    out_m3 = fmax(out_m3, storage - total);                                 // :260
    storage = storage - out_m3;                                             // :267
    double f = storage / total;                                             // :271
    f = (f != f || f < 0) ? 0.0 : f;                                        // :272-273
    R.Storage[j] = storage;
    R.Fill[j] = f;
    R.OutM3[j] = out_m3;
    return out_m3;
}
// lakes.dynamic_inloop, lakes.py:199-297 (Modified Puls): returns the outflow volume [m3]
__device__ __forceinline__ double lake_substep(const ChanPtrs::Lake &K, int j, double inflow, double DtRouting)
{
    const double lake_in = (inflow + K.InflowOld[j]) * 0.5;                 // :224
    K.InflowOld[j] = inflow;                                                // :227
    const double factor = K.Factor[j];
    const double indicator = K.Storage[j] / DtRouting - 0.5 * K.Outflow[j] + lake_in;   // :234
    const double r = -factor + sqrt(K.FactorSqr[j] + 2 * indicator);        // :243
    const double outflow = r * r;
    const double out_m3 = outflow * DtRouting;                              // :248
    double storage = (indicator - outflow * 0.5) * DtRouting;               // :252
    storage = (storage != storage || storage < 0) ? 0.0 : storage;          // :255-256 (NaN / negative -> 0)
    K.Balance[j] = K.Balance[j] + (lake_in * DtRouting - out_m3);           // :262
    K.Outflow[j] = outflow;
    K.Storage[j] = storage;
    K.Level[j] = storage / K.Area[j];                                       // :266
    K.OutM3[j] = out_m3;
    return out_m3;
}

__device__ __forceinline__ double pw(double x, double y) { return lfm::pw(x, y); }

// one routing.dynamic sub-step of one pixel (hydrological_modules/routing.py:462-604).
// U1/U2: upstream inflow of the main channel / floodplain for this sub-step.
// In 3/5 mode (QZ) qk / q2k / qr1 / qr2 hold z = Q^(1/5) (lf_kw_solve.cuh): Q^beta = z^3 makes the volume
// L*alpha*Q^beta free, and the reference's re-derivation ChanQKin = (ChanM3Kin/(L*alpha))^(1/beta) (:531) is the
// identity on z except where the floodplain volume is floored at Chan2M3Start (:586-593), where z becomes z2floor.
struct ChanLocal {
    double qk, m3, sum, q2k, m32, cs2a, s1, sumnl, chanq;
};

template <bool QZ>
__device__ __forceinline__ void chan_substep(const ChanPtrs &C, int i, double U1, double U2, ChanLocal &X,
                                             double L, double invL, double alpha, double a, double sideDt, bool isch,
                                             double &qr1, double &qr2, double alpha2, double a2, double qlimit,
                                             double m3limit, double c2start, double c2qstart, double z2floor)
{
    double side = isch ? sideDt * invL * C.InvDtRouting : 0.;  // :512
    if (!C.split) {
        if (isnan(side)) side = 0.;  // :523
        if (QZ) {
            const double zn = lfkw::solve_z(U1, X.qk, side * L, a);
            qr1 = zn;
            X.m3 = lfm::dmax(L * alpha * (zn * zn * zn), 0.0);  // ChanLength * ChannelAlpha * ChanQKin**Beta, :527-530
            X.qk = zn;                                     // :531 is the identity on z
            X.chanq = lfkw::pow5(zn);
        } else {
            double qn = lfkw::solve(U1, X.qk, side * L, a, C.P);
            qr1 = qn;
            X.m3 = lfm::dmax(L * alpha * pw(qn, C.P.beta), 0.0);          // :527-530
            X.qk = pw(X.m3 * invL * (1 / alpha), C.P.inv_beta);      // :531
            X.chanq = X.qk;
        }
        X.sum += X.chanq;                                         // :537
    } else if (QZ) {
        const double tot = X.m3 + X.m32;
        const double ratio = tot > 0 ? X.m3 / tot : 0.0;
        double s1 = (tot - c2start) > m3limit ? ratio * side : side;
        s1 = fabs(side) < 1e-7 ? side : s1;
        X.s1 = s1;
        double s2 = side - s1;
        s2 = s2 + c2qstart * invL;
        const double zn = lfkw::solve_z(U1, X.qk, s1 * L, a);
        qr1 = zn;
        X.m3 = lfm::dmax(L * alpha * (zn * zn * zn), 0.0);
        X.qk = zn;
        const double zn2 = lfkw::solve_z(U2, X.q2k, s2 * L, a2);
        qr2 = zn2;
        double m32 = L * alpha2 * (zn2 * zn2 * zn2);
        double z2 = zn2;
        if (m32 - c2start < 0.0) {  // :586-587: volume floored at Chan2M3Start, discharge re-derived from it (:593)
            m32 = c2start;
            z2 = z2floor;
        }
        X.m32 = m32;
        X.cs2a = (m32 - c2start) * invL;
        X.q2k = z2;
        X.chanq = lfm::dmax(lfkw::pow5(zn) + lfkw::pow5(z2) - qlimit, 0.0);
        X.sumnl = X.sum;
        X.sum += X.chanq;
    } else {
        const double tot = X.m3 + X.m32;
        const double ratio = tot > 0 ? X.m3 / tot : 0.0;          // :549
        double s1 = (tot - c2start) > m3limit ? ratio * side : side;  // :558-559
        s1 = fabs(side) < 1e-7 ? side : s1;                       // :564
        X.s1 = s1;
        double s2 = side - s1;
        s2 = s2 + c2qstart * invL;                                // :566-568
        double qn = lfkw::solve(U1, X.qk, s1 * L, a, C.P);        // :573
        qr1 = qn;
        X.m3 = lfm::dmax(L * alpha * pw(qn, C.P.beta), 0.0);
        X.qk = pw(X.m3 * invL * (1 / alpha), C.P.inv_beta);
        double qn2 = lfkw::solve(U2, X.q2k, s2 * L, a2, C.P);     // :583
        qr2 = qn2;
        double m32 = L * alpha2 * pw(qn2, C.P.beta);
        if (m32 - c2start < 0.0) m32 = c2start;                   // :586-587
        X.m32 = m32;
        X.cs2a = (m32 - c2start) * invL;                          // :590
        X.q2k = pw(m32 * invL * (1 / alpha2), C.P.inv_beta);      // :593
        X.chanq = lfm::dmax(X.qk + X.q2k - qlimit, 0.0);               // :597
        X.sumnl = X.sum;
        X.sum += X.chanq;
    }
}

constexpr int CH_THREADS = 128;

// one work item of the channel wavefront: position i on diagonal d, sub-step s = d - level
template <bool QZ, bool HASX, bool HASS>
__device__ __forceinline__ void chan_item(const ChanPtrs &C, int i, int d)
{
    int s = d - C.lev[i];
    double *Qr = (s & 1) ? C.Qr1 : C.Qr0;
    double *Q2r = (s & 1) ? C.Q2r1 : C.Q2r0;
    int xs = -1;
    if (HASX) {
        xs = C.X.xslot[i];
        if (xs <= -2) {   // ghost of a pixel another rank owns: its routed discharge of this sub-step, not solved here
            const bool inert = xs == lfx::INERT;
            Qr[i] = inert ? 0.0 : lfx::take(lfx::import_slot(C.X, -2 - xs, 0, s), C.X.abort_flag);
            if (C.split) Q2r[i] = inert ? 0.0 : lfx::take(lfx::import_slot(C.X, -2 - xs, 1, s), C.X.abort_flag);
            return;
        }
    }
    int c0 = C.cfirst[i], c1 = C.cend ? C.cend[i] : C.cfirst[i + 1];
    double U1 = 0., U2 = 0.;
    double sideDt = C.sideDt[i];
    int sid = 0;
    if (HASS) sid = C.sid[i];
    if (HASS && sid != 0) {
        // A structure pixel: structures.initial turns the pixels draining into it into pits (structures.py:51-59), so
        // nothing is routed into it (U = 0); its inflow is the ChanQ those pixels had after the PREVIOUS sub-step
        // (np.bincount(downstruct, ChanQ), reservoir.py:190 / lakes.py:215 -- summed in pixel order = slot order), and
        // the outflow volume of the sub-step joins its side flow (routing.py:472-476).
        const double *CQprev = ((C.gs0 + s) & 1) ? C.CQ0 : C.CQ1;
        double inflow = 0.;
        for (int k = c0; k < c1; ++k) inflow += CQprev[k];
        sideDt += sid > 0 ? reservoir_substep(C.res, sid - 1, inflow, C.DtRouting)
                          : lake_substep(C.lake, -sid - 1, inflow, C.DtRouting);
        c1 = c0;
    }
    for (int k = c0; k < c1; ++k) U1 += QZ ? lfkw::pow5(Qr[k]) : Qr[k];
    ChanLocal X;
    X.qk = C.Qk[i];
    X.sum = s == 0 ? 0. : C.sumDis[i];  // sumDisDay = 0 before the sub-step loop, Lisflood_dynamic.py:177
    const double L = C.L[i], alpha = C.alpha[i], a = C.a[i];
    const bool isch = C.isChan[i] != 0;
    double alpha2 = 0, a2 = 0, ql = 0, m3l = 0, c2s = 0, c2q = 0, z2f = 0;
    if (C.split) {
        for (int k = c0; k < c1; ++k) U2 += QZ ? lfkw::pow5(Q2r[k]) : Q2r[k];
        if (QZ) z2f = C.z2floor[i];
        X.m3 = C.M3[i];
        X.m32 = C.M32[i];
        X.q2k = C.Q2k[i];
        alpha2 = C.alpha2[i];
        a2 = C.a2[i];
        ql = C.QLimit[i];
        m3l = C.M3Limit[i];
        c2s = C.C2M3Start[i];
        c2q = C.C2QStart[i];
    }
    double qr1, qr2 = 0.;
    chan_substep<QZ>(C, i, U1, U2, X, L, 1 / L, alpha, a, sideDt, isch, qr1, qr2, alpha2, a2, ql, m3l, c2s, c2q, z2f);
    Qr[i] = qr1;
    if (HASX && xs >= 0) {
        lfx::push(lfx::export_slot(C.X, xs, 0, s), qr1);
        if (C.split) lfx::push(lfx::export_slot(C.X, xs, 1, s), qr2);
    }
    if (HASS && C.feeds[i]) (((C.gs0 + s) & 1) ? C.CQ1 : C.CQ0)[i] = X.chanq;
    C.Qk[i] = X.qk;
    C.sumDis[i] = X.sum;
    const bool last = s == C.S - 1;
    if (C.split) {
        Q2r[i] = qr2;
        C.Q2k[i] = X.q2k;
        C.M3[i] = X.m3;
        C.M32[i] = X.m32;
        if (last) {
            C.CS2A[i] = X.cs2a;
            C.S1[i] = X.s1;
            C.sumNotLast[i] = X.sumnl;
            C.ChanQ[i] = X.chanq;
        }
    } else if (last) {
        C.M3[i] = X.m3;
        C.ChanQ[i] = X.chanq;
    }
}

// wavefront diagonal over channel-network pixels: positions [lo, hi)
template <bool QZ, bool HASX, bool HASS>
__global__ void __launch_bounds__(CH_THREADS) k_chan_diagonal(ChanPtrs C, int lo, int hi, int d)
{
    int i = lo + blockIdx.x * CH_THREADS + threadIdx.x;
    if (i >= hi) return;
    chan_item<QZ, HASX, HASS>(C, i, d);
}

// A run of consecutive NARROW diagonals [d0, d1) -- at most CH_RUN_THREADS items each -- in ONE block: a __syncthreads()
// between two diagonals instead of a kernel boundary.  An OPTION ("narrow_runs"), off by default: in the reference's
// ordering the levels count the distance to the outlet downwards, so only the first few hundred diagonals of a deep basin
// are narrow (the far ends of its longest paths) -- nothing to gain there --, and on C3 the fat block cannot share an SM
// with the persistent isolated-pixel kernel, so the tail it merges runs AFTER that kernel instead of beside it (channel
// stage 48.7 ms instead of 44.1).  n_connected: end of the last level (the isolated pixels follow it in position order).
constexpr int CH_RUN_THREADS = 512;
template <bool QZ, bool HASX, bool HASS>
__global__ void __launch_bounds__(CH_RUN_THREADS) k_chan_narrow_run(ChanPtrs C, int d0, int d1, int nlev,
                                                                    const int32_t *__restrict__ level_start, int n_connected)
{
    for (int d = d0; d < d1; ++d) {
        const int lo_lev = d - C.S + 1 > 0 ? d - C.S + 1 : 0;
        const int hi_lev = d < nlev - 1 ? d : nlev - 1;
        const int i = level_start[lo_lev] + threadIdx.x;
        const int hi = hi_lev == nlev - 1 ? n_connected : level_start[hi_lev + 1];
        if (i < hi) chan_item<QZ, HASX, HASS>(C, i, d);
        __syncthreads();
    }
}

// Isolated pixels that are NOT channel pixels (the bulk of LddKinematic: ~89 % of C3) receive no side flow
// (routing.py:512), so their sub-steps do not depend on this step's runoff: they are queued in chunks of CH_THREADS
// positions behind one atomic counter and drained by TWO launches of this kernel -- an early one with a small
// persistent grid that shares the SMs with the soil stage (FP64 pipe and issue slots the HBM-bound stencil leaves
// idle), and a machine-filling one in the channel stage that takes whatever is left.  Channel pixels of a chunk are
// skipped here and handled by k_chan_isolated_list once their side flow exists.
template <bool QZ>
__global__ void __launch_bounds__(CH_THREADS) k_chan_isolated_ws(ChanPtrs C, int lo, int hi, int *__restrict__ next)
{
    __shared__ int s_chunk;
    const int nchunks = (hi - lo + CH_THREADS - 1) / CH_THREADS;
    for (;;) {
        if (threadIdx.x == 0) s_chunk = atomicAdd(next, 1);
        __syncthreads();
        const int chunk = s_chunk;
        __syncthreads();
        if (chunk >= nchunks) return;
        const int i = lo + chunk * CH_THREADS + threadIdx.x;
        if (i >= hi || C.isChan[i]) continue;
        ChanLocal X;
        X.qk = C.Qk[i];
        X.sum = 0.;
        X.m3 = C.M3[i];
        const double L = C.L[i], alpha = C.alpha[i], a = C.a[i];
        double alpha2 = 0, a2 = 0, ql = 0, m3l = 0, c2s = 0, c2q = 0, z2f = 0;
        if (C.split) {
            if (QZ) z2f = C.z2floor[i];
            X.m32 = C.M32[i];
            X.q2k = C.Q2k[i];
            alpha2 = C.alpha2[i];
            a2 = C.a2[i];
            ql = C.QLimit[i];
            m3l = C.M3Limit[i];
            c2s = C.C2M3Start[i];
            c2q = C.C2QStart[i];
        }
        double qr1 = 0., qr2 = 0.;
        const double invL = 1 / L;
#pragma unroll 1
        for (int s = 0; s < C.S; ++s)
            chan_substep<QZ>(C, i, 0., 0., X, L, invL, alpha, a, 0., false, qr1, qr2, alpha2, a2, ql, m3l, c2s, c2q, z2f);
        C.Qr0[i] = qr1;
        C.Qr1[i] = qr1;
        C.Qk[i] = X.qk;
        C.sumDis[i] = X.sum;
        C.M3[i] = X.m3;
        C.ChanQ[i] = X.chanq;
        if (C.split) {
            C.Q2r0[i] = qr2;
            C.Q2r1[i] = qr2;
            C.Q2k[i] = X.q2k;
            C.M32[i] = X.m32;
            C.CS2A[i] = X.cs2a;
            C.S1[i] = X.s1;
            C.sumNotLast[i] = X.sumnl;
        }
    }
}
// positions of the isolated pixels that ARE channel pixels (rare: one-pixel channels), any order
__global__ void k_chan_iso_collect(const uint8_t *__restrict__ isChan, int lo, int hi, int32_t *__restrict__ list,
                                   int *__restrict__ count)
{
    int i = lo + blockIdx.x * blockDim.x + threadIdx.x;
    if (i < hi && isChan[i]) list[atomicAdd(count, 1)] = i;
}
template <bool QZ>
__global__ void __launch_bounds__(CH_THREADS) k_chan_isolated_list(ChanPtrs C, const int32_t *__restrict__ list, int count)
{
    int k = blockIdx.x * CH_THREADS + threadIdx.x;
    if (k >= count) return;
    const int i = list[k];
    ChanLocal X;
    X.qk = C.Qk[i];
    X.sum = 0.;
    X.m3 = C.M3[i];
    const double L = C.L[i], alpha = C.alpha[i], a = C.a[i], sideDt = C.sideDt[i];
    double alpha2 = 0, a2 = 0, ql = 0, m3l = 0, c2s = 0, c2q = 0, z2f = 0;
    if (C.split) {
        if (QZ) z2f = C.z2floor[i];
        X.m32 = C.M32[i];
        X.q2k = C.Q2k[i];
        alpha2 = C.alpha2[i];
        a2 = C.a2[i];
        ql = C.QLimit[i];
        m3l = C.M3Limit[i];
        c2s = C.C2M3Start[i];
        c2q = C.C2QStart[i];
    }
    double qr1 = 0., qr2 = 0.;
    const double invL = 1 / L;
#pragma unroll 1
    for (int s = 0; s < C.S; ++s)
        chan_substep<QZ>(C, i, 0., 0., X, L, invL, alpha, a, sideDt, true, qr1, qr2, alpha2, a2, ql, m3l, c2s, c2q, z2f);
    C.Qr0[i] = qr1;
    C.Qr1[i] = qr1;
    C.Qk[i] = X.qk;
    C.sumDis[i] = X.sum;
    C.M3[i] = X.m3;
    C.ChanQ[i] = X.chanq;
    if (C.split) {
        C.Q2r0[i] = qr2;
        C.Q2r1[i] = qr2;
        C.Q2k[i] = X.q2k;
        C.M32[i] = X.m32;
        C.CS2A[i] = X.cs2a;
        C.S1[i] = X.s1;
        C.sumNotLast[i] = X.sumnl;
    }
}

// Lisflood_dynamic.py:194-229 and hydrological_modules/routing.py:695-703
__global__ void k_chan_post(int n, int split, int qz, int S, double DtSec, const double *__restrict__ M3, const double *__restrict__ M32,
                            const double *__restrict__ C2M3Start, const double *__restrict__ L,
                            const double *__restrict__ sumDisDay, const double *__restrict__ ChanQ,
                            const double *__restrict__ Qk, const uint8_t *__restrict__ atLast,
                            const double *__restrict__ PixelArea_ch, double *__restrict__ ChanM3,
                            double *__restrict__ TotalCS, double *__restrict__ sumDis, double *__restrict__ ChanQAvg,
                            double *__restrict__ DischargeM3Out, double *__restrict__ FlowVelocity,
                            double *__restrict__ TravelDistance, int *__restrict__ nonfinite, double *__restrict__ CumQ)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (CumQ) CumQ[i] += ChanQ[i];   // InitLisflood / repAverageDis: avgdis = CumQ / TimeSinceStart, Lisflood_dynamic.py:224-227
    // the -n option of the reference (flagnancheck, kinematic_wave_parallel.py:180-184): non-finite discharge is
    // reported once, it does not stop the run
    if (nonfinite && !isfinite(ChanQ[i])) *nonfinite = 1;
    const double invL = 1 / L[i];
    const double m3 = split ? M3[i] + M32[i] - C2M3Start[i] : M3[i];
    ChanM3[i] = m3;
    TotalCS[i] = m3 * invL;
    const double sd = sumDisDay[i];
    sumDis[i] += sd;
    ChanQAvg[i] = sd / S;
    DischargeM3Out[i] += atLast[i] ? ChanQ[i] * DtSec : 0.;
    if (FlowVelocity) {
        const double area = lfm::dmax(M3[i] * invL, 0.01);
        const double q = qz ? lfkw::pow5(Qk[i]) : Qk[i];
        double fv = lfm::dmin(q / area, 0.36 * pw(q, 0.24));
        fv *= lfm::dmin(sqrt(PixelArea_ch[i]) * invL, 1.);
        FlowVelocity[i] = fv;
        TravelDistance[i] = fv * DtSec;
    }
}

// ---- overland flow: three routers in one sweep (surface_routing.py:143-153) ----
struct OfPtrs {
    const int32_t *cfirst, *cend;
    lfx::View X;
    const double *DirectRunoff, *SurfOther, *SurfForest, *MMtoM3;
    const double *a[3];   // Other, Forest, Direct
    double *Qnew[3];
    const double *Qold[3];
    double PixelLength, InvPixelLength, InvDtSec;
    lfkw::Params P;
};
template <bool QZ, bool HASX>
__global__ void __launch_bounds__(128) k_of_level(OfPtrs O, int lo, int hi)
{
    int i = lo + blockIdx.x * 128 + threadIdx.x;
    if (i >= hi) return;
    int xs = -1;
    if (HASX) {
        xs = O.X.xslot[i];
        if (xs <= -2) {   // ghost: the owner's new discharge of the three routers
            for (int r = 0; r < 3; ++r)
                O.Qnew[r][i] = xs == lfx::INERT ? 0.0 : lfx::take(lfx::import_slot(O.X, -2 - xs, r, 0), O.X.abort_flag);
            return;
        }
    }
    int c0 = O.cfirst[i], c1 = O.cend ? O.cend[i] : O.cfirst[i + 1];
    const double mm2m3 = O.MMtoM3[i];
    const double runoff[3] = {O.SurfOther[i], O.SurfForest[i], O.DirectRunoff[i]};
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        double U = 0.;
        const double side = runoff[r] * mm2m3 * O.InvPixelLength * O.InvDtSec;  // [m3 s-1 m-1], :143-149
        if (QZ) {
            for (int k = c0; k < c1; ++k) U += lfkw::pow5(O.Qnew[r][k]);
            U = lfkw::solve_z(U, O.Qold[r][i], side * O.PixelLength, O.a[r][i]);
        } else {
            for (int k = c0; k < c1; ++k) U += O.Qnew[r][k];
            U = lfkw::solve(U, O.Qold[r][i], side * O.PixelLength, O.a[r][i], O.P);
        }
        O.Qnew[r][i] = U;
        if (HASX && xs >= 0) lfx::push(lfx::export_slot(O.X, xs, r, 0), U);
    }
}
// surface_routing.py:191-212 (+ scatter of ToChanM3RunoffDt into channel order)
__global__ void k_of_post(int n, const double *__restrict__ QO, const double *__restrict__ QF, const double *__restrict__ QD,
                          const double *__restrict__ GwToChan, const double *__restrict__ MMtoM3,
                          const uint8_t *__restrict__ isChannel, const int32_t *__restrict__ soil_to_chan, int qz,
                          double DtSec, double InvNoRoutSteps, double *__restrict__ sideDt_chan, double *__restrict__ ToChanM3Runoff,
                          double *__restrict__ OFToChanM3, double *__restrict__ Qall_out)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double qall = qz ? lfkw::pow5(QD[i]) + lfkw::pow5(QO[i]) + lfkw::pow5(QF[i])
                           : QD[i] + QO[i] + QF[i];  // :195
    const double oftochan = isChannel[i] ? qall * DtSec : 0.;       // :199
    const double tochan = GwToChan[i] * MMtoM3[i] + oftochan;      // :211
    sideDt_chan[soil_to_chan[i]] = tochan * InvNoRoutSteps;         // :212
    if (ToChanM3Runoff) {
        ToChanM3Runoff[i] = tochan;
        OFToChanM3[i] = oftochan;
        Qall_out[i] = qall;
    }
}


// ---- feeder modules of a step, fused: readmeteo scaling (readmeteo.py:61-81), snow (snow.py:95-187),
// frost (frost.py:61-78).  Raw meteo maps arrive in the reference's compressed order (float32 as read from NetCDF, or
// float64) and are gathered into the soil stage's storage order here, so no separate layout translation is needed.
struct ScalarOrMap {
    const double *map;   // storage order, or nullptr
    double value;
    __device__ __forceinline__ double at(int64_t i) const { return map ? map[i] : value; }
};
struct FeedPtrs {
    int64_t n;
    const int32_t *pix_of_pos;
    const void *prec, *tavg, *et0, *e0;
    ScalarOrMap PrScaling, CalEvaporation, DeltaTSnow, SnowSeason, TempSnow, SnowFactor, SnowMeltCoef, TempMelt, lat_rad, Kfrost,
        Afrost, FrostIndexThreshold, SnowWaterEquivalent;
    double DtDay, snowmelt_coeff, ice_n, ice_s;
    double scale[4], offset[4];                                  // CF packing of int16 input: value = raw * scale + offset
    int32_t decode_f32, pad_;                                    // unpack in float32 arithmetic (as a float32 decoder does)
    double *SnowCoverS, *FrostIndex, *TotalPrecipitation;       // state
    double *Rain, *SnowMelt, *ETRef, *EWRef, *ESRef;            // forcing of the soil stage
    uint8_t *frozen;
    double *Snow, *SnowCover, *Precipitation, *Tavg;             // outputs kept for reporting (may be nullptr)
};
// DT: element type of the raw maps -- 0 float64, 1 float32, 2 int16 packed by the CF convention (scale_factor / add_offset
// attributes of the NetCDF variable, which the reference's reader applies on the host: netcdf.py:231-232)
template <int DT>
__device__ __forceinline__ double raw_value(const void *src, int p, double scale, double offset, int decode_f32)
{
    if (DT == 0) return ((const double *)src)[p];
    if (DT == 1) return (double)((const float *)src)[p];
    const int16_t r = ((const int16_t *)src)[p];
    if (decode_f32) return (double)((float)r * (float)scale + (float)offset);
    return (double)r * scale + offset;
}
template <int DT>
__global__ void __launch_bounds__(256) k_feeder(FeedPtrs F)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= F.n) return;
    const int p = F.pix_of_pos[i];
    const double praw = raw_value<DT>(F.prec, p, F.scale[0], F.offset[0], F.decode_f32);
    const double tavg = raw_value<DT>(F.tavg, p, F.scale[1], F.offset[1], F.decode_f32);
    const double et0 = raw_value<DT>(F.et0, p, F.scale[2], F.offset[2], F.decode_f32);
    const double e0 = raw_value<DT>(F.e0, p, F.scale[3], F.offset[3], F.decode_f32);
    const double dt = F.DtDay;
    const double prec = praw * dt * F.PrScaling.at(i);               // readmeteo.py:66
    const double cal = F.CalEvaporation.at(i);
    const double etref = et0 * dt * cal, ewref = e0 * dt * cal;      // :68-69
    F.ETRef[i] = etref;
    F.EWRef[i] = ewref;
    F.ESRef[i] = (ewref + etref) / 2;                                // :78
    // snow.py:104-118
    const bool north = F.lat_rad.at(i) > 0;
    const double seas = F.SnowSeason.at(i) * (north ? F.snowmelt_coeff : -F.snowmelt_coeff) + F.SnowMeltCoef.at(i);
    const double summer = north ? F.ice_n : F.ice_s;
    const double dts = F.DeltaTSnow.at(i), tsnow = F.TempSnow.at(i), sfac = F.SnowFactor.at(i), tmelt = F.TempMelt.at(i);
    double snow = 0., rain = 0., melt = 0., cover = 0.;
#pragma unroll
    for (int z = 0; z < 3; ++z) {                                    // :150-178, zones A (highest), B, C
        const double tz = tavg + dts * (z - 1);
        const double snow_s = tz < tsnow ? sfac * prec : 0.;
        const double rain_s = tz >= tsnow ? prec : 0.;
        double melt_s = (tz - tmelt) * seas * (1 + 0.01 * rain_s) * dt;
        const double ice = (z < 2 ? tavg : tz) * 7.0 * dt * summer;
        double sc = F.SnowCoverS[(int64_t)z * F.n + i];
        melt_s = fmax(fmin(melt_s + ice, sc), 0.);
        sc = sc + snow_s - melt_s;
        F.SnowCoverS[(int64_t)z * F.n + i] = sc;
        snow += snow_s;
        rain += rain_s;
        melt += melt_s;
        cover += sc;
    }
    snow /= 3;
    rain /= 3;
    melt /= 3;
    cover /= 3;
    F.Rain[i] = rain;
    F.SnowMelt[i] = melt;
    F.TotalPrecipitation[i] = F.TotalPrecipitation[i] + (snow + rain);   // :186
    // frost.py:66-73
    double fi = F.FrostIndex[i];
    const double rate = -(1 - F.Afrost.at(i)) * fi - tavg * exp(-0.04 * F.Kfrost.at(i) * cover / F.SnowWaterEquivalent.at(i));
    fi = fmax(fi + rate * dt, 0.);
    fi = fi > 57.0 ? 57.0 : fi;
    F.FrostIndex[i] = fi;
    F.frozen[i] = fi > F.FrostIndexThreshold.at(i) ? 1 : 0;
    if (F.Snow) {
        F.Snow[i] = snow;
        F.SnowCover[i] = cover;
        F.Precipitation[i] = prec;
        F.Tavg[i] = tavg;
    }
}
// LAITerm = exp(-kgb * LAI), leafarea.py:90
__global__ void k_lai_term(const double *__restrict__ lai, ScalarOrMap kgb, double *__restrict__ out, int64_t n)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double k = kgb.at(i);
    for (int v = 0; v < 3; ++v) out[(int64_t)v * n + i] = exp(-k * lai[(int64_t)v * n + i]);
}

// ---- layout translation ----
// reference (compressed row-major) order -> position order: a gather.  Four positions per thread: the four index loads
// and then the four gathered values are in flight together (the kernel is bound by the latency of the scattered reads).
constexpr int PERM_U = 4;
__global__ void __launch_bounds__(256) k_rows_to_pos(const double *__restrict__ src, double *__restrict__ dst,
                                                     const int32_t *__restrict__ pix_of_pos, int64_t n, int rows)
{
    const int64_t i0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * PERM_U;
    if (i0 >= n) return;
    int p[PERM_U];
#pragma unroll
    for (int u = 0; u < PERM_U; ++u) p[u] = i0 + u < n ? pix_of_pos[i0 + u] : 0;
    for (int r = 0; r < rows; ++r) {
        const double *s = src + (int64_t)r * n;
        double *d = dst + (int64_t)r * n;
        double x[PERM_U];
#pragma unroll
        for (int u = 0; u < PERM_U; ++u) x[u] = s[p[u]];
#pragma unroll
        for (int u = 0; u < PERM_U; ++u)
            if (i0 + u < n) d[i0 + u] = x[u];
    }
}
__global__ void k_rows_to_pix(const double *__restrict__ src, double *__restrict__ dst, const int32_t *__restrict__ pos_of_pix,
                              int64_t n, int rows)
{
    int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    int i = pos_of_pix[p];
    for (int r = 0; r < rows; ++r) dst[(int64_t)r * n + p] = src[(int64_t)r * n + i];
}
// position order -> reference order, narrowed to float32 (OutputMapsDataType = float32, netcdf.py:478); maps stored as
// z = Q^(1/5) are converted on the way
__global__ void k_rows_to_pix_f32(const double *__restrict__ src, float *__restrict__ dst, const int32_t *__restrict__ pos_of_pix,
                                  int64_t n, int rows, int as_z)
{
    int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    int i = pos_of_pix[p];
    for (int r = 0; r < rows; ++r) {
        const double v = src[(int64_t)r * n + i];
        dst[(int64_t)r * n + p] = (float)(as_z ? lfkw::pow5(v) : v);
    }
}
__global__ void k_u8_to_pos(const uint8_t *__restrict__ src, uint8_t *__restrict__ dst, const int32_t *__restrict__ pix_of_pos,
                            int64_t n)
{
    const int64_t i0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * PERM_U;
    if (i0 >= n) return;
    uint8_t x[PERM_U];
#pragma unroll
    for (int u = 0; u < PERM_U; ++u) x[u] = src[i0 + u < n ? pix_of_pos[i0 + u] : 0];
#pragma unroll
    for (int u = 0; u < PERM_U; ++u)
        if (i0 + u < n) dst[i0 + u] = x[u];
}
__global__ void k_i32_rows_to_pos(const int32_t *__restrict__ src, int32_t *__restrict__ dst, const int32_t *__restrict__ pix_of_pos,
                                  int64_t n)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[pix_of_pos[i]];
}
__global__ void k_struct_feeds(const int32_t *__restrict__ sid, const int32_t *__restrict__ cfirst, const int32_t *__restrict__ cend,
                               uint8_t *__restrict__ feeds, int64_t n)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || sid[i] == 0) return;
    for (int k = cfirst[i]; k < (cend ? cend[i] : cfirst[i + 1]); ++k) feeds[k] = 1;
}
__global__ void k_struct_cq_init(const uint8_t *__restrict__ feeds, const double *__restrict__ ChanQ, double *__restrict__ CQ0,
                                 double *__restrict__ CQ1, int64_t n)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || !feeds[i]) return;
    CQ0[i] = ChanQ[i];
    CQ1[i] = ChanQ[i];
}
__global__ void k_rows_differ(const double *__restrict__ a, const double *__restrict__ b, int64_t n, int *__restrict__ flag)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && a[i] != b[i]) *flag = 1;
}
__global__ void k_soil_to_chan(const int32_t *__restrict__ pix_of_pos_soil, const int32_t *__restrict__ pos_of_pix_chan,
                               int32_t *__restrict__ soil_to_chan, int64_t n)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) soil_to_chan[i] = pos_of_pix_chan[pix_of_pos_soil[i]];
}
__global__ void k_make_a(const double *__restrict__ alpha, const double *__restrict__ dx, double dxs, double dt,
                         double *__restrict__ a, int64_t n)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] = alpha[i] * (dx ? dx[i] : dxs) / dt;  // a_dx_div_dt, kinematic_wave_parallel.py:126
}
__global__ void k_of_m3(const double *__restrict__ Q, const double *__restrict__ alpha, double PixelLength, double beta,
                        int qz, double *__restrict__ M3, int64_t n)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double z = Q[i];
    M3[i] = PixelLength * alpha[i] * (qz ? z * z * z : pw(z, beta));  // surface_routing.py:191-193
}
// discharge maps are stored as z = Q^(1/5) in 3/5 mode: conversions at the API boundary
__global__ void k_q_to_z(double *__restrict__ v, int64_t n)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) v[i] = lfkw::z_of_q(v[i]);
}
__global__ void k_z_to_q(double *__restrict__ v, int64_t n)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) v[i] = lfkw::pow5(v[i]);
}
// (Chan2M3Start / (L * alpha2))^(1/3): z of the discharge re-derived from the floored floodplain volume
__global__ void k_z2floor(const double *__restrict__ c2start, const double *__restrict__ L, const double *__restrict__ alpha2,
                          double *__restrict__ out, int64_t n)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = pw(c2start[i] * (1 / L[i]) * (1 / alpha2[i]), 1.0 / 3.0);
}

}  // namespace


struct lf_model {
    lf_model_config cfg;
    int64_t n = 0;
    double DtDay = 0, DtRouting = 0;
    lf_graph *g_of = nullptr, *g_ch = nullptr;
    std::map<std::string, std::unique_ptr<Field>> fields;
    std::map<std::string, std::unique_ptr<lf::DevBuf<uint8_t>>> flags;
    lf::DevBuf<double> stage;        // 3N staging (compressed order)
    lf::DevBuf<uint8_t> stage_u8;
    lf::DevBuf<int32_t> soil_to_chan;
    lf::DevBuf<int> flag;
    lf::DevBuf<int32_t> soil_list, soil_list_cnt;  // deferred soil columns (lf_soil_kernel.cuh)
    lf::DevBuf<double> soil_side;                  // their dense side records (lean build)
    int32_t soil_list_cap = 0;
    // asynchronous input path (lf_model_set_async): H2D on a copy stream into a per-map staging buffer, layout
    // translation on the compute stream once the copy has landed
    cudaStream_t copy_stream = nullptr;
    cudaStream_t side_stream = nullptr;   // the channel wavefront runs here, concurrently with k_chan_isolated
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    // non-channel isolated pixels: chunk queue drained by an early launch (low-priority stream, alongside the soil
    // stage) and by a late launch in the channel stage
    cudaStream_t early_stream = nullptr;
    cudaEvent_t ev_early_fork = nullptr, ev_early_join = nullptr;
    lf::DevBuf<int> iso_next;             // [0] next chunk, [1] length of iso_chan_list, [2] non-finite flag
    lf::DevBuf<int32_t> iso_chan_list;    // isolated pixels that are channel pixels
    int iso_chan_count = 0;
    bool iso_list_dirty = true, early_in_flight = false;
    // LDD-cut exchange (multi-GPU): one region, one import block per graph (overland: 3 routers x 1 step per edge;
    // channel: 1 or 2 sections x NoRoutSteps per edge)
    lf_xchg *xchg = nullptr;
    struct XSide {
        lf::DevBuf<int32_t> xslot;
        lf::DevBuf<double *> exp_dst;
        lf::DevBuf<long long> exp_stride;
        double *imp = nullptr;
        long long imp_parity_stride = 0;
        int32_t n_export = 0, n_import = 0;
    } xs_of, xs_ch;
    int32_t x_parity = 0;
    // asynchronous output path (lf_model_get_async): layout translation on the compute stream into a per-map staging
    // buffer, device-to-host copy on its own stream
    cudaStream_t out_stream = nullptr;
    std::map<std::string, std::unique_ptr<lf::DevBuf<double>>> out_stage;
    std::map<std::string, cudaEvent_t> out_ready, out_copied;
    // feeder modules: scalar parameters (a map of the same name, when set, takes precedence) and the raw-forcing staging
    std::map<std::string, double> scalars;
    lf::DevBuf<uint8_t> raw_stage[2];      // 4 maps each, double-buffered: the upload of step k+1 overlaps step k
    cudaEvent_t raw_copied[2] = {nullptr, nullptr}, raw_consumed[2] = {nullptr, nullptr};
    int raw_turn = 0;
    // structures in the routing sub-step loop (lf_model_set_structures)
    struct Structures {
        int32_t n_res = 0, n_lake = 0;
        lf::DevBuf<int32_t> sid;
        lf::DevBuf<uint8_t> feeds;
        lf::DevBuf<double> CQ0, CQ1;
        std::map<std::string, std::unique_ptr<lf::DevBuf<double>>> v;   // per-structure parameter / state arrays
        bool cq_dirty = true;
    } st;
    lf::GraphCache graphs_of, graphs_ch;
    int use_graphs = 1;                   // option "cuda_graphs"
    int narrow_runs = 0;                  // option "narrow_runs": runs of narrow wavefront diagonals in one single-block launch
                                          // (off: measured slower on C3, without effect on deep basins -- DESIGN.md 4.5)
    int accumulate_discharge = 0;         // option "accumulate_discharge" (InitLisflood / repAverageDis)
    int overlap_isolated = 0;             // option "overlap_isolated" (measured: no gain)
    int early_blocks_per_sm = 2;          // option "early_blocks_per_sm"
    int iso_blocks_per_sm = 6;            // option "isolated_blocks_per_sm" (0: one block per chunk)
    int nancheck = 0;                     // option "flagnancheck"
    std::map<std::string, std::unique_ptr<lf::DevBuf<double>>> async_stage;
    std::map<std::string, cudaEvent_t> async_copied, async_consumed;
    bool soil_profile = false;            // time the kernels of the soil stage individually (lf_model_soil_stats)
    std::vector<cudaEvent_t> soil_ev;     // 9 events
    // per-stage device timing (CUDA events on the library stream), accumulated on query
    std::vector<cudaEvent_t> ev_pool;
    std::vector<int> ev_used;  // 4 events per recorded step: start, after soil, after overland, after channel
    double t_soil = 0, t_of = 0, t_chan = 0;
    int64_t t_steps = 0;
    bool quintic = false;            // Beta == 0.6: discharge state held as z = Q^(1/5)
    int64_t steps = 0;               // model steps done (parity of the overland discharge buffers)
    bool params_dirty = true;        // a_dx_div_dt arrays need a rebuild
    int64_t bytes = 0;
    ~lf_model()
    {
        if (g_of) lf_graph_destroy(g_of);
        if (g_ch) lf_graph_destroy(g_ch);
        for (cudaEvent_t e : ev_pool) cudaEventDestroy(e);
        for (cudaEvent_t e : soil_ev) cudaEventDestroy(e);
        for (auto &kv : async_copied) cudaEventDestroy(kv.second);
        for (auto &kv : async_consumed) cudaEventDestroy(kv.second);
        if (copy_stream) cudaStreamDestroy(copy_stream);
        if (side_stream) cudaStreamDestroy(side_stream);
        if (early_stream) cudaStreamDestroy(early_stream);
        if (out_stream) cudaStreamDestroy(out_stream);
        for (auto &kv : out_ready) cudaEventDestroy(kv.second);
        for (auto &kv : out_copied) cudaEventDestroy(kv.second);
        for (int k = 0; k < 2; ++k) {
            if (raw_copied[k]) cudaEventDestroy(raw_copied[k]);
            if (raw_consumed[k]) cudaEventDestroy(raw_consumed[k]);
        }
        if (ev_early_fork) cudaEventDestroy(ev_early_fork);
        if (ev_early_join) cudaEventDestroy(ev_early_join);
        if (ev_fork) cudaEventDestroy(ev_fork);
        if (ev_join) cudaEventDestroy(ev_join);
    }
};

namespace {

struct FieldSpec {
    const char *name;
    int rows;
    Order order;
    bool landuse;
    bool diag;
};

// every named map of the model (reference attribute names)
const FieldSpec SPECS[] = {
    // forcing
    {"Rain", 1, SOIL, false, false}, {"SnowMelt", 1, SOIL, false, false}, {"ETRef", 1, SOIL, false, false},
    {"EWRef", 1, SOIL, false, false}, {"ESRef", 1, SOIL, false, false}, {"LAI", 3, SOIL, false, false},
    {"LAITerm", 3, SOIL, false, false},
    // feeder modules (readmeteo scaling, snow, frost: lf_model_feed): parameters, state, outputs
    {"PrScaling", 1, SOIL, false, false}, {"CalEvaporation", 1, SOIL, false, false}, {"DeltaTSnow", 1, SOIL, false, false},
    {"SnowSeason", 1, SOIL, false, false}, {"TempSnow", 1, SOIL, false, false}, {"SnowFactor", 1, SOIL, false, false},
    {"SnowMeltCoef", 1, SOIL, false, false}, {"TempMelt", 1, SOIL, false, false}, {"lat_rad", 1, SOIL, false, false},
    {"Kfrost", 1, SOIL, false, false}, {"Afrost", 1, SOIL, false, false}, {"FrostIndexThreshold", 1, SOIL, false, false},
    {"SnowWaterEquivalent", 1, SOIL, false, false}, {"kgb", 1, SOIL, false, false},
    {"SnowCoverS", 3, SOIL, false, false}, {"FrostIndex", 1, SOIL, false, false}, {"TotalPrecipitation", 1, SOIL, false, false},
    {"Snow", 1, SOIL, false, false}, {"SnowCover", 1, SOIL, false, false}, {"Precipitation", 1, SOIL, false, false},
    {"Tavg", 1, SOIL, false, false},
    // per-pixel parameters
    {"b_Xinanjiang", 1, SOIL, false, false}, {"PowerPrefFlow", 1, SOIL, false, false}, {"UpperZoneK", 1, SOIL, false, false},
    {"GwPercStep", 1, SOIL, false, false}, {"LowerZoneK", 1, SOIL, false, false}, {"LZThreshold", 1, SOIL, false, false},
    {"GwLossStep", 1, SOIL, false, false}, {"SoilFraction", 3, SOIL, false, false},
    {"DirectRunoffFraction", 1, SOIL, false, false}, {"WaterFraction", 1, SOIL, false, false},
    {"MMtoM3", 1, SOIL, false, false}, {"PixelArea", 1, CHAN, false, false},
    // land-use parameters
    {"KSat1a", 3, SOIL, true, false}, {"KSat1b", 3, SOIL, true, false}, {"KSat2", 3, SOIL, true, false},
    {"GenuInvM1a", 3, SOIL, true, false}, {"GenuInvM1b", 3, SOIL, true, false}, {"GenuInvM2", 3, SOIL, true, false},
    {"WRes1a", 3, SOIL, true, false}, {"WRes1b", 3, SOIL, true, false}, {"WRes2", 3, SOIL, true, false},
    {"WS1a", 3, SOIL, true, false}, {"WS1b", 3, SOIL, true, false}, {"WS2", 3, SOIL, true, false},
    {"WWP1a", 3, SOIL, true, false}, {"WWP1b", 3, SOIL, true, false}, {"WWP2", 3, SOIL, true, true},
    {"WFC1a", 3, SOIL, true, false}, {"WFC1b", 3, SOIL, true, false}, {"WFC2", 3, SOIL, true, true},
    {"SoilDepth1a", 3, SOIL, true, true}, {"SoilDepth1b", 3, SOIL, true, true}, {"SoilDepth2", 3, SOIL, true, true},
    {"CropCoef", 3, SOIL, true, false}, {"CropGroupNumber", 3, SOIL, true, false},
    // soil state
    {"CumInterception", 3, SOIL, false, false}, {"W1a", 3, SOIL, false, false}, {"W1b", 3, SOIL, false, false},
    {"W2", 3, SOIL, false, false}, {"UZ", 3, SOIL, false, false}, {"DSLR", 3, SOIL, false, false},
    {"LZ", 1, SOIL, false, false}, {"CumInterSealed", 1, SOIL, false, false}, {"LZInflowCUM", 1, SOIL, false, false},
    {"TaCUM", 1, SOIL, false, false}, {"TaInterceptionCUM", 1, SOIL, false, false}, {"ESActCUM", 1, SOIL, false, false},
    {"GwLossCUM", 1, SOIL, false, false},
    // runoff components
    {"DirectRunoff", 1, SOIL, false, false}, {"SurfOther", 1, SOIL, false, false}, {"SurfForest", 1, SOIL, false, false},
    {"GwToChan", 1, SOIL, false, false},
    // overland flow
    {"OFAlpha", 3, SOIL, false, false}, {"OFQOther", 1, SOIL, false, false}, {"OFQForest", 1, SOIL, false, false},
    {"OFQDirect", 1, SOIL, false, false},
    // channel
    {"ChanLength", 1, CHAN, false, false}, {"ChannelAlpha", 1, CHAN, false, false}, {"ChannelAlpha2", 1, CHAN, false, false},
    {"ChanQKin", 1, CHAN, false, false}, {"ChanM3Kin", 1, CHAN, false, false}, {"ChanQ", 1, CHAN, false, false},
    {"sumDisDay", 1, CHAN, false, false}, {"sumDis", 1, CHAN, false, false}, {"ChanQAvg", 1, CHAN, false, false},
    {"ChanM3", 1, CHAN, false, false}, {"TotalCrossSectionArea", 1, CHAN, false, false},
    {"DischargeM3Out", 1, CHAN, false, false}, {"ToChanM3RunoffDt", 1, CHAN, false, false}, {"CumQ", 1, CHAN, false, false},
    {"Chan2QKin", 1, CHAN, false, false}, {"Chan2M3Kin", 1, CHAN, false, false}, {"CrossSection2Area", 1, CHAN, false, false},
    {"Sideflow1Chan", 1, CHAN, false, false}, {"sumDisDay_NOTlast", 1, CHAN, false, false}, {"QLimit", 1, CHAN, false, false},
    {"M3Limit", 1, CHAN, false, false}, {"Chan2M3Start", 1, CHAN, false, false}, {"Chan2QStart", 1, CHAN, false, false},
    {"FlowVelocity", 1, CHAN, false, true}, {"TravelDistance", 1, CHAN, false, true},
    // diagnostics (vegetation, pixel)
    {"Interception", 3, SOIL, false, true}, {"TaInterception", 3, SOIL, false, true}, {"LeafDrainage", 3, SOIL, false, true},
    {"potential_transpiration", 3, SOIL, false, true}, {"Ta", 3, SOIL, false, true}, {"ESAct", 3, SOIL, false, true},
    {"PrefFlow", 3, SOIL, false, true}, {"Infiltration", 3, SOIL, false, true},
    {"AvailableWaterForInfiltration", 3, SOIL, false, true}, {"SeepTopToSubA", 3, SOIL, false, true},
    {"SeepTopToSubB", 3, SOIL, false, true}, {"SeepSubToGW", 3, SOIL, false, true}, {"Theta1a", 3, SOIL, false, true},
    {"Theta1b", 3, SOIL, false, true}, {"Theta2", 3, SOIL, false, true}, {"Sat1a", 3, SOIL, false, true},
    {"Sat1b", 3, SOIL, false, true}, {"Sat1", 3, SOIL, false, true}, {"Sat2", 3, SOIL, false, true},
    {"UZOutflow", 3, SOIL, false, true}, {"GwPercUZLZ", 3, SOIL, false, true}, {"RWS", 3, SOIL, false, true},
    {"Theta", 3, SOIL, false, true}, {"SurfaceRunSoil", 3, SOIL, false, true}, {"W1", 3, SOIL, false, true},
    // diagnostics (pixel)
    {"RainSnowmelt", 1, SOIL, false, true}, {"EWaterAct", 1, SOIL, false, true}, {"InterSealed", 1, SOIL, false, true},
    {"TASealed", 1, SOIL, false, true}, {"TaInterceptionAll", 1, SOIL, false, true}, {"TaPixel", 1, SOIL, false, true},
    {"ESActPixel", 1, SOIL, false, true}, {"PrefFlowPixel", 1, SOIL, false, true}, {"InfiltrationPixel", 1, SOIL, false, true},
    {"ThetaAll", 1, SOIL, false, true}, {"SeepTopToSubPixelA", 1, SOIL, false, true},
    {"SeepTopToSubPixelB", 1, SOIL, false, true}, {"SeepSubToGWPixel", 1, SOIL, false, true},
    {"Theta1aPixel", 1, SOIL, false, true}, {"Theta1bPixel", 1, SOIL, false, true}, {"Theta2Pixel", 1, SOIL, false, true},
    {"UZOutflowPixel", 1, SOIL, false, true}, {"GwPercUZLZPixel", 1, SOIL, false, true}, {"GwLossLZ", 1, SOIL, false, true},
    {"LZOutflow", 1, SOIL, false, true}, {"LZAvInflow", 1, SOIL, false, true}, {"SurfaceRunoff", 1, SOIL, false, true},
    {"TotalRunoff", 1, SOIL, false, true}, {"ToChanM3Runoff", 1, SOIL, false, true}, {"OFToChanM3", 1, SOIL, false, true},
    {"Qall", 1, SOIL, false, true}, {"OFM3Other", 1, SOIL, false, true}, {"OFM3Forest", 1, SOIL, false, true},
    {"OFM3Direct", 1, SOIL, false, true},
};

const FieldSpec *find_spec(const char *name)
{
    for (const FieldSpec &s : SPECS)
        if (strcmp(s.name, name) == 0) return &s;
    return nullptr;
}

// device buffer of a field, allocated (zero-filled) on first use
int field(lf_model *m, const char *name, Field **out)
{
    auto it = m->fields.find(name);
    if (it != m->fields.end()) {
        *out = it->second.get();
        return LF_OK;
    }
    const FieldSpec *s = find_spec(name);
    if (!s) {
        lf::set_error("unknown map name '%s'", name);
        return LF_ERR_INVALID;
    }
    std::unique_ptr<Field> f(new Field());
    f->rows = s->rows;
    f->order = s->order;
    f->landuse = s->landuse;
    f->diag = s->diag;
    f->as_z = m->quintic && (strcmp(name, "OFQOther") == 0 || strcmp(name, "OFQForest") == 0 ||
                             strcmp(name, "OFQDirect") == 0 || strcmp(name, "ChanQKin") == 0 || strcmp(name, "Chan2QKin") == 0);
    LF_CHECK(f->buf.alloc((size_t)s->rows * m->n));
    LF_CUDA(cudaMemsetAsync(f->buf.p, 0, (size_t)s->rows * m->n * sizeof(double), lf::stream()));
    m->bytes += (int64_t)s->rows * m->n * 8;
    *out = f.get();
    m->fields[name] = std::move(f);
    return LF_OK;
}

#define FIELD(var, name)              \
    Field *var##_f = nullptr;         \
    LF_CHECK(field(m, name, &var##_f)); \
    double *var = var##_f->buf.p

int flag_buf(lf_model *m, const char *name, uint8_t **out)
{
    auto it = m->flags.find(name);
    if (it == m->flags.end()) {
        std::unique_ptr<lf::DevBuf<uint8_t>> b(new lf::DevBuf<uint8_t>());
        LF_CHECK(b->alloc(m->n));
        LF_CUDA(cudaMemsetAsync(b->p, 0, m->n, lf::stream()));
        m->bytes += m->n;
        *out = b->p;
        m->flags[name] = std::move(b);
        return LF_OK;
    }
    *out = it->second->p;
    return LF_OK;
}

// pointer of land-use row r of a (landuse, pixel) parameter
const double *lu_row(lf_model *m, Field *f, int r) { return f->buf.p + (int64_t)f->row_index[r] * m->n; }

int soil_stage(lf_model *m)
{
    using namespace lfsoil;
    cudaStream_t st = lf::stream();
    Ptrs P;
    memset(&P, 0, sizeof(P));
    Diag D;
    memset(&D, 0, sizeof(D));
    P.n = m->n;
    {
        FIELD(a, "Rain"); P.Rain = a;
    }
#define BIND(member, name)   \
    {                        \
        FIELD(_p, name);     \
        P.member = _p;       \
    }
#define BINDLU(member, name)                                          \
    {                                                                 \
        FIELD(_p, name);                                              \
        (void)_p;                                                     \
        for (int r = 0; r < 3; ++r) P.member[r] = lu_row(m, _p_f, r); \
    }
#define BINDD(member, name)  \
    {                        \
        FIELD(_p, name);     \
        D.member = _p;       \
    }
    BIND(SnowMelt, "SnowMelt") BIND(ETRef, "ETRef") BIND(EWRef, "EWRef") BIND(ESRef, "ESRef") BIND(LAI, "LAI")
    BIND(LAITerm, "LAITerm") BIND(bX, "b_Xinanjiang") BIND(PowPref, "PowerPrefFlow") BIND(UZK, "UpperZoneK")
    BIND(GwPercStep, "GwPercStep") BIND(LZK, "LowerZoneK") BIND(LZThreshold, "LZThreshold") BIND(GwLossStep, "GwLossStep")
    BIND(SoilFraction, "SoilFraction") BIND(DirectRunoffFraction, "DirectRunoffFraction") BIND(WaterFraction, "WaterFraction")
    BINDLU(KSat1a, "KSat1a") BINDLU(KSat1b, "KSat1b") BINDLU(KSat2, "KSat2") BINDLU(InvM1a, "GenuInvM1a")
    BINDLU(InvM1b, "GenuInvM1b") BINDLU(InvM2, "GenuInvM2") BINDLU(WRes1a, "WRes1a") BINDLU(WRes1b, "WRes1b")
    BINDLU(WRes2, "WRes2") BINDLU(WS1a, "WS1a") BINDLU(WS1b, "WS1b") BINDLU(WS2, "WS2") BINDLU(WWP1a, "WWP1a")
    BINDLU(WWP1b, "WWP1b") BINDLU(WFC1a, "WFC1a") BINDLU(WFC1b, "WFC1b") BINDLU(CropCoef, "CropCoef")
    BINDLU(CropGroup, "CropGroupNumber")
    BIND(CumInterception, "CumInterception") BIND(W1a, "W1a") BIND(W1b, "W1b") BIND(W2, "W2") BIND(UZ, "UZ") BIND(DSLR, "DSLR")
    BIND(LZ, "LZ") BIND(CumInterSealed, "CumInterSealed") BIND(LZInflowCUM, "LZInflowCUM") BIND(TaCUM, "TaCUM")
    BIND(TaInterceptionCUM, "TaInterceptionCUM") BIND(ESActCUM, "ESActCUM") BIND(GwLossCUM, "GwLossCUM")
    BIND(DirectRunoff, "DirectRunoff") BIND(SurfOther, "SurfOther") BIND(SurfForest, "SurfForest") BIND(GwToChan, "GwToChan")
    uint8_t *frozen = nullptr;
    LF_CHECK(flag_buf(m, "isFrozenSoil", &frozen));
    P.frozen = frozen;
    P.DtDay = m->DtDay;
    P.InvDtDay = 1 / m->DtDay;
    P.AvWaterThreshold = m->cfg.AvWaterThreshold;
    P.CourantCrit = m->cfg.CourantCrit;
    P.DrainedFraction = m->cfg.DrainedFraction;
    P.LeafDrainageK = m->cfg.LeafDrainageK;
    P.SMaxSealed = m->cfg.SMaxSealed;
    P.TimeSinceStart = (double)(m->steps + 1);
    // contributions handed from the column kernel to the pixel kernel, and the deferred-column lists
    {
        {
            P.cstride = m->cfg.diagnostics ? 8 : 3;  // lean: CS_UZOUT, CS_GWPERC, CS_SURF (the rest is summed in the first pass)
            auto it = m->fields.find("__cbuf");
            if (it == m->fields.end()) {
                std::unique_ptr<Field> f(new Field());
                f->rows = 3 * P.cstride;
                LF_CHECK(f->buf.alloc((size_t)3 * P.cstride * m->n));
                m->bytes += (int64_t)3 * P.cstride * m->n * 8;
                it = m->fields.emplace("__cbuf", std::move(f)).first;
            }
            P.cbuf = it->second->buf.p;
        }
        uint8_t *pdef = nullptr;
        LF_CHECK(flag_buf(m, "__pix_deferred", &pdef));
        P.pix_deferred = pdef;
        if (!m->soil_list.p) {
            // capacity per bucket: a quarter of the columns (typically ~6 % of all columns are deferred in total);
            // a full list degrades gracefully: the column is integrated by k_soil_pixel_flagged
            // capacity per bucket: 3/4 of the columns with diagnostics; an eighth in the lean build, whose queued columns
            // also get a dense side record (1-2 % of the columns are deferred after spin-up, ~6 % from a cold start)
            const bool side = !m->cfg.diagnostics && !(getenv("LF_SOIL_SIDE") && atoi(getenv("LF_SOIL_SIDE")) == 0);
            m->soil_list_cap = (int32_t)std::max<int64_t>(1024, side ? m->n / 8 : (3 * m->n) / 4);
            if (const char *e = getenv("LF_SOIL_LIST_CAP")) m->soil_list_cap = std::max(1, atoi(e));  // tests: force the overflow path
            LF_CHECK(m->soil_list.alloc((size_t)lfsoil::NBUCKET * m->soil_list_cap));
            LF_CHECK(m->soil_list_cnt.alloc(lfsoil::NBUCKET));
            m->bytes += (int64_t)lfsoil::NBUCKET * m->soil_list_cap * 4;
            if (side) {
                LF_CHECK(m->soil_side.alloc((size_t)lfsoil::NBUCKET * lfsoil::NSIDE * m->soil_list_cap));
                m->bytes += (int64_t)lfsoil::NBUCKET * lfsoil::NSIDE * m->soil_list_cap * 8;
            }
        }
        P.list = m->soil_list.p;
        P.list_cnt = m->soil_list_cnt.p;
        P.list_cap = m->soil_list_cap;
        P.side = m->soil_side.p;
        LF_CUDA(cudaMemsetAsync(m->soil_list_cnt.p, 0, lfsoil::NBUCKET * sizeof(int32_t), st));
    }
    auto tick = [&](int k) {
        if (m->soil_profile) cudaEventRecord(m->soil_ev[k], st);
    };
    tick(0);
    // persistent grids of the deferred-column kernel: resident blocks per SM x SMs (never more than the lists can hold)
    const unsigned grid_def_cap = lf::blocks_for((int64_t)lfsoil::NBUCKET * m->soil_list_cap, lfsoil::SOIL_THREADS);
    const unsigned grid_def = std::min<unsigned>(grid_def_cap, (unsigned)(lf::sm_count() * (m->cfg.diagnostics ? 4 : 8)));
    const unsigned grid_def6 = std::min<unsigned>(grid_def_cap, (unsigned)(lf::sm_count() * 6));
    const unsigned grid_pix = lf::blocks_for(m->n, 256);
    if (m->cfg.diagnostics) {
        BINDLU(WWP2, "WWP2") BINDLU(WFC2, "WFC2") BINDLU(Depth1a, "SoilDepth1a") BINDLU(Depth1b, "SoilDepth1b")
        BINDLU(Depth2, "SoilDepth2")
        BINDD(Interception, "Interception") BINDD(TaInterception, "TaInterception") BINDD(LeafDrainage, "LeafDrainage")
        BINDD(potential_transpiration, "potential_transpiration") BINDD(Ta, "Ta") BINDD(ESAct, "ESAct")
        BINDD(PrefFlow, "PrefFlow") BINDD(Infiltration, "Infiltration")
        BINDD(AvailableWaterForInfiltration, "AvailableWaterForInfiltration") BINDD(SeepTopToSubA, "SeepTopToSubA")
        BINDD(SeepTopToSubB, "SeepTopToSubB") BINDD(SeepSubToGW, "SeepSubToGW") BINDD(Theta1a, "Theta1a")
        BINDD(Theta1b, "Theta1b") BINDD(Theta2, "Theta2") BINDD(Sat1a, "Sat1a") BINDD(Sat1b, "Sat1b") BINDD(Sat1, "Sat1")
        BINDD(Sat2, "Sat2") BINDD(UZOutflow, "UZOutflow") BINDD(GwPercUZLZ, "GwPercUZLZ") BINDD(RWS, "RWS")
        BINDD(Theta, "Theta") BINDD(SurfaceRunSoil, "SurfaceRunSoil") BINDD(W1, "W1")
        BINDD(RainSnowmelt, "RainSnowmelt") BINDD(EWaterAct, "EWaterAct") BINDD(InterSealed, "InterSealed")
        BINDD(TASealed, "TASealed") BINDD(TaInterceptionAll, "TaInterceptionAll") BINDD(TaPixel, "TaPixel")
        BINDD(ESActPixel, "ESActPixel") BINDD(PrefFlowPixel, "PrefFlowPixel") BINDD(InfiltrationPixel, "InfiltrationPixel")
        BINDD(ThetaAll, "ThetaAll") BINDD(SeepTopToSubPixelA, "SeepTopToSubPixelA")
        BINDD(SeepTopToSubPixelB, "SeepTopToSubPixelB") BINDD(SeepSubToGWPixel, "SeepSubToGWPixel")
        BINDD(Theta1aPixel, "Theta1aPixel") BINDD(Theta1bPixel, "Theta1bPixel") BINDD(Theta2Pixel, "Theta2Pixel")
        BINDD(UZOutflowPixel, "UZOutflowPixel") BINDD(GwPercUZLZPixel, "GwPercUZLZPixel") BINDD(GwLossLZ, "GwLossLZ")
        BINDD(LZOutflow, "LZOutflow") BINDD(LZAvInflow, "LZAvInflow") BINDD(SurfaceRunoff, "SurfaceRunoff")
        BINDD(TotalRunoff, "TotalRunoff")
        static_assert(sizeof(int32_t) * 2 == sizeof(double), "NoSubS shares a double-sized slot");
        Field *ns = nullptr;
        {
            auto it = m->fields.find("__NoSubS");
            if (it == m->fields.end()) {
                std::unique_ptr<Field> f(new Field());
                f->rows = 3;
                LF_CHECK(f->buf.alloc((size_t)3 * m->n));
                ns = f.get();
                m->fields["__NoSubS"] = std::move(f);
            } else {
                ns = it->second.get();
            }
        }
        D.NoSubS = (int32_t *)ns->buf.p;
    }
    // First pass of the lean build: k_soil_staged (inputs staged in shared memory by bulk async copies); the diagnostics
    // build uses k_soil_fused (direct loads: it writes 50 more maps and is not a performance configuration).
    // LF_SOIL_VARIANT selects the launch shape (tuning): 13 = 64 px x 5 blocks/SM (default), 11 = 64 x 4, 10 = 32 x 9,
    // 12 = 32 x 8.  Measured on C3 (profiles/): direct loads spend 61 % of the warp-stall samples on global-load latency
    // (first pass 29.4 ms at best); staged: 18 ms.
    int variant = 13;
    if (const char *e = getenv("LF_SOIL_VARIANT")) {  // read per call: tools/soil_variants.py switches it between runs
        const int v = atoi(e);
        if (v >= 10 && v <= 13) variant = v;
    }
    const int force_plain = getenv("LF_SOIL_PLAIN") ? atoi(getenv("LF_SOIL_PLAIN")) : 0;
    const int def_mb = getenv("LF_SOIL_DEF_MB") ? atoi(getenv("LF_SOIL_DEF_MB")) : 6;  // resident blocks of the deferred kernel
#define LF_SOIL_REST(DG, MBD)                                                                         \
    do {                                                                                              \
        LF_LAUNCH_CHECK();                                                                            \
        tick(1);                                                                                      \
        if (!DG && P.side)                                                                            \
            k_soil_veg_deferred<false, 6, true><<<grid_def6, lfsoil::SOIL_THREADS, 0, st>>>(P, D);    \
        else if (MBD == 8 && def_mb == 6)                                                             \
            k_soil_veg_deferred<DG, 6><<<grid_def6, lfsoil::SOIL_THREADS, 0, st>>>(P, D);             \
        else                                                                                          \
            k_soil_veg_deferred<DG, MBD><<<grid_def, lfsoil::SOIL_THREADS, 0, st>>>(P, D);            \
        LF_LAUNCH_CHECK();                                                                            \
        for (int b = 0; b < lfsoil::NBUCKET; ++b) tick(2 + b);                                        \
        k_soil_pixel_flagged<DG><<<grid_pix, 256, 0, st>>>(P, D);                                     \
    } while (0)
#define LF_SOIL_LAUNCH(DG, TILE, MB, MBD)                                                              \
    do {                                                                                              \
        k_soil_fused<DG, TILE, MB><<<lf::blocks_for(m->n, TILE), 3 * TILE, 0, st>>>(P, D);            \
        LF_SOIL_REST(DG, MBD);                                                                        \
    } while (0)
#define LF_SOIL_STAGED(TILE, MB)                                                                                          \
    do {                                                                                                                  \
        const size_t smem = lfsoil::staged_smem_bytes<TILE>(G.nrows);                                                     \
        static bool attr_set = false;                                                                                     \
        if (!attr_set) {                                                                                                  \
            LF_CUDA(cudaFuncSetAttribute(k_soil_staged<TILE, MB>, cudaFuncAttributeMaxDynamicSharedMemorySize,            \
                                         (int)lfsoil::staged_smem_bytes<TILE>(lfsoil::MAX_STAGE_ROWS)));                  \
            LF_CUDA(cudaFuncSetAttribute(k_soil_staged<TILE, MB>, cudaFuncAttributePreferredSharedMemoryCarveout,         \
                                         cudaSharedmemCarveoutMaxShared));                                                \
            attr_set = true;                                                                                              \
        }                                                                                                                 \
        k_soil_staged<TILE, MB><<<lf::blocks_for(m->n, TILE), 3 * TILE, smem, st>>>(P, G, force_plain);                   \
        LF_SOIL_REST(false, 8);                                                                                           \
    } while (0)
    if (m->cfg.diagnostics) {
        LF_SOIL_LAUNCH(true, 64, 3, 4);
    } else {
        // copy plan: per-pixel rows, (V,N) rows per fraction, land-use groups stored once per distinct pointer set
        Stage G;
        memset(&G, 0, sizeof(G));
        int nr = 0;
        const double *pix[lfsoil::NPIXROW] = {P.Rain, P.SnowMelt, P.ETRef, P.EWRef, P.ESRef, P.bX, P.PowPref, P.UZK,
                                              P.GwPercStep, P.LZK, P.LZThreshold, P.GwLossStep, P.DirectRunoffFraction,
                                              P.WaterFraction, P.LZ, P.CumInterSealed, P.LZInflowCUM, P.TaCUM,
                                              P.TaInterceptionCUM, P.ESActCUM, P.GwLossCUM};
        for (int r = 0; r < lfsoil::NPIXROW; ++r) G.src[nr++] = pix[r];
        const double *veg[lfsoil::NVEGROW] = {P.SoilFraction, P.LAI, P.LAITerm, P.CumInterception, P.W1a, P.W1b, P.W2,
                                              P.UZ, P.DSLR};
        for (int v = 0; v < 3; ++v)
            for (int r = 0; r < lfsoil::NVEGROW; ++r) G.src[nr++] = veg[r] + (int64_t)v * m->n;
        const double *const *ga[lfsoil::NLUA] = {P.KSat1a, P.KSat1b, P.InvM1a, P.InvM1b, P.WRes1a, P.WRes1b,
                                                 P.WS1a,   P.WS1b,   P.WWP1a,  P.WWP1b,  P.WFC1a,  P.WFC1b};
        const double *const *gb[lfsoil::NLUB] = {P.KSat2, P.InvM2, P.WRes2, P.WS2};
        const double *const *gc[lfsoil::NLUC] = {P.CropCoef, P.CropGroup};
        auto add_group = [&](const double *const *const *rows, int nparam, int32_t *off) {
            for (int v = 0; v < 3; ++v) {
                int same = -1;
                for (int u = 0; u < v && same < 0; ++u) {
                    bool eq = true;
                    for (int q = 0; q < nparam; ++q) eq = eq && rows[q][u] == rows[q][v];
                    if (eq) same = u;
                }
                if (same >= 0) {
                    off[v] = off[same];
                } else {
                    off[v] = nr;
                    for (int q = 0; q < nparam; ++q) G.src[nr++] = rows[q][v];
                }
            }
        };
        add_group(ga, lfsoil::NLUA, G.offA);
        add_group(gb, lfsoil::NLUB, G.offB);
        add_group(gc, lfsoil::NLUC, G.offC);
        G.nrows = nr;
        bool aligned = (reinterpret_cast<uintptr_t>(P.frozen) & 15) == 0;
        for (int r = 0; r < nr; ++r) aligned = aligned && (reinterpret_cast<uintptr_t>(G.src[r]) & 15) == 0;
        G.bulk_ok = aligned ? 1 : 0;
        if (variant == 11) LF_SOIL_STAGED(64, 4);
        else if (variant == 12) LF_SOIL_STAGED(32, 8);
        else if (variant == 10) LF_SOIL_STAGED(32, 9);
        else LF_SOIL_STAGED(64, 5);  // 34 land-use rows (C3) leave room for five 64-pixel tiles per SM
    }
#undef LF_SOIL_STAGED
#undef LF_SOIL_LAUNCH
#undef LF_SOIL_REST
    LF_LAUNCH_CHECK();
    tick(8);
    return LF_OK;
}

int build_a(lf_model *m, const char *alpha_name, int row, const double *dx, double dxs, double dt, const char *a_name,
            int a_row)
{
    FIELD(al, alpha_name);
    FIELD(a, a_name);
    k_make_a<<<lf::blocks_for(m->n, 256), 256, 0, lf::stream()>>>(al + (int64_t)row * m->n, dx, dxs, dt,
                                                                  a + (int64_t)a_row * m->n, m->n);
    LF_LAUNCH_CHECK();
    return LF_OK;
}

int internal(lf_model *m, const std::string &name, Order order, double **out);

int refresh_params(lf_model *m)
{
    if (!m->params_dirty) return LF_OK;
    // internal maps "__aOF" (3,N), "__aCh", "__aCh2"
    for (const char *nm : {"__aOF", "__aCh", "__aCh2"}) {
        if (m->fields.find(nm) == m->fields.end()) {
            std::unique_ptr<Field> f(new Field());
            f->rows = strcmp(nm, "__aOF") == 0 ? 3 : 1;
            f->order = f->rows == 3 ? SOIL : CHAN;
            LF_CHECK(f->buf.alloc((size_t)f->rows * m->n));
            m->bytes += (int64_t)f->rows * m->n * 8;
            m->fields[nm] = std::move(f);
        }
    }
    for (int r = 0; r < 3; ++r)
        LF_CHECK(build_a(m, "OFAlpha", r, nullptr, m->cfg.PixelLength, m->cfg.DtSec, "__aOF", r));
    FIELD(L, "ChanLength");
    LF_CHECK(build_a(m, "ChannelAlpha", 0, L, 0., m->DtRouting, "__aCh", 0));
    if (m->cfg.SplitRouting) {
        LF_CHECK(build_a(m, "ChannelAlpha2", 0, L, 0., m->DtRouting, "__aCh2", 0));
        if (m->quintic) {
            double *zf = nullptr;
            LF_CHECK(internal(m, "__z2floor", CHAN, &zf));
            FIELD(c2s, "Chan2M3Start");
            FIELD(al2, "ChannelAlpha2");
            k_z2floor<<<lf::blocks_for(m->n, 256), 256, 0, lf::stream()>>>(c2s, L, al2, zf, m->n);
            LF_LAUNCH_CHECK();
        }
    }
    m->params_dirty = false;
    return LF_OK;
}

// internal double-buffer of a discharge map: name + "__r0/__r1"
int internal(lf_model *m, const std::string &name, Order order, double **out)
{
    auto it = m->fields.find(name);
    if (it == m->fields.end()) {
        std::unique_ptr<Field> f(new Field());
        f->order = order;
        LF_CHECK(f->buf.alloc(m->n));
        LF_CUDA(cudaMemsetAsync(f->buf.p, 0, m->n * sizeof(double), lf::stream()));
        m->bytes += m->n * 8;
        *out = f->buf.p;
        m->fields[name] = std::move(f);
        return LF_OK;
    }
    *out = it->second->buf.p;
    return LF_OK;
}

int make_view(lf_model *m, lf_model::XSide &xs, int cap, int nsec, lfx::View &X)
{
    memset(&X, 0, sizeof(X));
    if (!m->xchg || !xs.xslot.p) return LF_OK;
    X.xslot = xs.xslot.p;
    X.exp_dst = xs.exp_dst.p;
    X.exp_stride = xs.exp_stride.p;
    X.imp = xs.imp;
    X.imp_parity_stride = xs.imp_parity_stride;
    X.cap = cap;
    X.nsec = nsec;
    X.parity = m->x_parity;
    return lf::xchg_view_base(m->xchg, &X.abort_flag);
}

int surface_stage(lf_model *m)
{
    cudaStream_t st = lf::stream();
    LF_CHECK(refresh_params(m));
    lf_graph *g = m->g_of;
    OfPtrs O;
    memset(&O, 0, sizeof(O));
    if (m->xchg) LF_CHECK(lf_xchg_begin(m->xchg, &m->x_parity));   // closed at the end of the channel stage
    LF_CHECK(make_view(m, m->xs_of, 1, 3, O.X));
    const bool hasx = O.X.xslot != nullptr;
    O.cfirst = g->cfirst.p;
    O.cend = g->cend.p;
    FIELD(dr, "DirectRunoff");
    FIELD(so, "SurfOther");
    FIELD(sf, "SurfForest");
    FIELD(mm, "MMtoM3");
    FIELD(aof, "__aOF");
    O.DirectRunoff = dr;
    O.SurfOther = so;
    O.SurfForest = sf;
    O.MMtoM3 = mm;
    const char *names[3] = {"OFQOther", "OFQForest", "OFQDirect"};
    double *qcur[3], *qalt[3];
    for (int r = 0; r < 3; ++r) {
        Field *f = nullptr;
        LF_CHECK(field(m, names[r], &f));
        qcur[r] = f->buf.p;  // the named map always holds the CURRENT discharge
        LF_CHECK(internal(m, std::string(names[r]) + "__alt", SOIL, &qalt[r]));
        O.a[r] = aof + (int64_t)r * m->n;
        O.Qold[r] = qcur[r];
        O.Qnew[r] = qalt[r];
    }
    O.PixelLength = m->cfg.PixelLength;
    O.InvPixelLength = 1.0 / m->cfg.PixelLength;
    O.InvDtSec = 1 / m->cfg.DtSec;
    O.P = lfkw::make_params(m->cfg.Beta);
    const std::vector<int32_t> &ls = g->h_level_start;
    auto sweep = [&]() -> int {
        for (int l = 0; l < g->n_orders; ++l) {
            int lo = ls[l], hi = ls[l + 1];
            if (hi <= lo) continue;
            if (hasx) {
                if (m->quintic) k_of_level<true, true><<<lf::blocks_for(hi - lo, 128), 128, 0, st>>>(O, lo, hi);
                else k_of_level<false, true><<<lf::blocks_for(hi - lo, 128), 128, 0, st>>>(O, lo, hi);
            } else {
                if (m->quintic) k_of_level<true, false><<<lf::blocks_for(hi - lo, 128), 128, 0, st>>>(O, lo, hi);
                else k_of_level<false, false><<<lf::blocks_for(hi - lo, 128), 128, 0, st>>>(O, lo, hi);
            }
            LF_LAUNCH_CHECK();
        }
        return LF_OK;
    };
    if (m->use_graphs && g->n_orders > 4) LF_CHECK(lf::run_captured(m->graphs_of, &O, sizeof(O), st, sweep));
    else LF_CHECK(sweep());
    // swap: the named maps must hold the new discharge
    for (int r = 0; r < 3; ++r) {
        Field *f = m->fields[names[r]].get();
        Field *alt = m->fields[std::string(names[r]) + "__alt"].get();
        std::swap(f->buf.p, alt->buf.p);
        std::swap(f->buf.n, alt->buf.n);
    }
    FIELD(qo, "OFQOther");
    FIELD(qf, "OFQForest");
    FIELD(qd, "OFQDirect");
    FIELD(gw, "GwToChan");
    FIELD(sideDt, "ToChanM3RunoffDt");
    uint8_t *isch = nullptr;
    LF_CHECK(flag_buf(m, "IsChannel", &isch));
    double *tochan = nullptr, *oftochan = nullptr, *qall = nullptr;
    if (m->cfg.diagnostics) {
        FIELD(t1, "ToChanM3Runoff");
        FIELD(t2, "OFToChanM3");
        FIELD(t3, "Qall");
        tochan = t1;
        oftochan = t2;
        qall = t3;
        const char *m3n[3] = {"OFM3Other", "OFM3Forest", "OFM3Direct"};
        const double *qs[3] = {qo, qf, qd};
        FIELD(ofa, "OFAlpha");
        for (int r = 0; r < 3; ++r) {
            FIELD(m3, m3n[r]);
            k_of_m3<<<lf::blocks_for(m->n, 256), 256, 0, st>>>(qs[r], ofa + (int64_t)r * m->n, m->cfg.PixelLength, m->cfg.Beta,
                                                               m->quintic ? 1 : 0, m3, m->n);
            LF_LAUNCH_CHECK();
        }
    }
    k_of_post<<<lf::blocks_for(m->n, 256), 256, 0, st>>>((int)m->n, qo, qf, qd, gw, mm, isch, m->soil_to_chan.p,
                                                         m->quintic ? 1 : 0, m->cfg.DtSec,
                                                         1 / (double)m->cfg.NoRoutSteps, sideDt, tochan, oftochan, qall);
    LF_LAUNCH_CHECK();
    return LF_OK;
}

// pointers of the channel stage (shared by the early launch of the isolated pixels and the stage itself)
int chan_ptrs(lf_model *m, ChanPtrs &C, double **m32_out, double **c2s_out)
{
    LF_CHECK(refresh_params(m));
    lf_graph *g = m->g_ch;
    memset(&C, 0, sizeof(C));
    C.n = (int32_t)m->n;
    C.lev = g->lev_of_pos.p;
    C.cfirst = g->cfirst.p;
    C.cend = g->cend.p;
    LF_CHECK(make_view(m, m->xs_ch, m->cfg.NoRoutSteps, m->cfg.SplitRouting ? 2 : 1, C.X));
    FIELD(qk, "ChanQKin");
    FIELD(m3, "ChanM3Kin");
    FIELD(sd, "sumDisDay");
    FIELD(cq, "ChanQ");
    FIELD(a, "__aCh");
    FIELD(L, "ChanLength");
    FIELD(al, "ChannelAlpha");
    FIELD(side, "ToChanM3RunoffDt");
    C.Qk = qk;
    C.M3 = m3;
    C.sumDis = sd;
    C.ChanQ = cq;
    C.a = a;
    C.L = L;
    C.alpha = al;
    C.sideDt = side;
    LF_CHECK(internal(m, "ChanQKin__r0", CHAN, &C.Qr0));
    LF_CHECK(internal(m, "ChanQKin__r1", CHAN, &C.Qr1));
    uint8_t *isch = nullptr;
    LF_CHECK(flag_buf(m, "IsChannelKinematic", &isch));
    C.isChan = isch;
    C.InvDtRouting = 1 / m->DtRouting;
    C.DtRouting = m->DtRouting;
    C.S = m->cfg.NoRoutSteps;
    C.split = m->cfg.SplitRouting;
    C.P = lfkw::make_params(m->cfg.Beta);
    if (m->st.n_res + m->st.n_lake > 0) {
        lf_model::Structures &T = m->st;
        C.sid = T.sid.p;
        C.feeds = T.feeds.p;
        C.CQ0 = T.CQ0.p;
        C.CQ1 = T.CQ1.p;
        C.gs0 = (int)((m->steps * (int64_t)m->cfg.NoRoutSteps) & 1);
        auto V = [&](const char *k) -> double * {
            auto it = T.v.find(k);
            return it == T.v.end() ? nullptr : it->second->p;
        };
        C.res.Storage = V("ReservoirStorageM3CC");
        C.res.Fill = V("ReservoirFillCC");
        C.res.OutM3 = V("QResOutM3DtCC");
        C.res.Total = V("TotalReservoirStorageM3CC");
        C.res.Cons = V("ConservativeStorageLimitCC");
        C.res.Norm = V("NormalStorageLimitCC");
        C.res.NormFlood = V("Normal_FloodStorageLimitCC");
        C.res.Flood = V("FloodStorageLimitCC");
        C.res.MinOut = V("MinReservoirOutflowCC");
        C.res.NormOut = V("NormalReservoirOutflowCC");
        C.res.NonDam = V("NonDamagingReservoirOutflowCC");
        C.res.DeltaO = V("DeltaO");
        C.res.DeltaLN = V("DeltaLN");
        C.res.DeltaNFL = V("DeltaNFL");
        C.lake.Storage = V("LakeStorageM3CC");
        C.lake.Outflow = V("LakeOutflowCC");
        C.lake.InflowOld = V("LakeInflowOldCC");
        C.lake.Balance = V("LakeStorageM3BalanceCC");
        C.lake.Level = V("LakeLevelCC");
        C.lake.OutM3 = V("QLakeOutM3DtCC");
        C.lake.Area = V("LakeAreaCC");
        C.lake.Factor = V("LakeFactor");
        C.lake.FactorSqr = V("LakeFactorSqr");
        if (T.cq_dirty) {   // ChanQ of the feeder pixels before the first sub-step (np.bincount(downstruct, ChanQ) at s = 0)
            k_struct_cq_init<<<lf::blocks_for(m->n, 256), 256, 0, lf::stream()>>>(T.feeds.p, cq, T.CQ0.p, T.CQ1.p, m->n);
            LF_LAUNCH_CHECK();
            T.cq_dirty = false;
        }
    }
    if (m32_out) *m32_out = nullptr;
    if (c2s_out) *c2s_out = nullptr;
    if (C.split) {
        FIELD(q2k, "Chan2QKin");
        FIELD(m32_, "Chan2M3Kin");
        FIELD(cs2a, "CrossSection2Area");
        FIELD(s1, "Sideflow1Chan");
        FIELD(snl, "sumDisDay_NOTlast");
        FIELD(a2, "__aCh2");
        FIELD(al2, "ChannelAlpha2");
        FIELD(ql, "QLimit");
        FIELD(ml, "M3Limit");
        FIELD(c2s_, "Chan2M3Start");
        FIELD(c2q, "Chan2QStart");
        C.Q2k = q2k;
        C.M32 = m32_;
        C.CS2A = cs2a;
        C.S1 = s1;
        C.sumNotLast = snl;
        C.a2 = a2;
        C.alpha2 = al2;
        C.QLimit = ql;
        C.M3Limit = ml;
        C.C2M3Start = c2s_;
        C.C2QStart = c2q;
        if (m32_out) *m32_out = m32_;
        if (c2s_out) *c2s_out = c2s_;
        LF_CHECK(internal(m, "Chan2QKin__r0", CHAN, &C.Q2r0));
        LF_CHECK(internal(m, "Chan2QKin__r1", CHAN, &C.Q2r1));
        if (m->quintic) {
            double *zf = nullptr;
            LF_CHECK(internal(m, "__z2floor", CHAN, &zf));
            C.z2floor = zf;
        }
    }
    return LF_OK;
}

int chan_streams(lf_model *m)
{
    if (m->side_stream) return LF_OK;
    int prio_lo = 0, prio_hi = 0;   // high priority: its small blocks take SM slots as the big kernel's blocks retire
    LF_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    LF_CUDA(cudaStreamCreateWithPriority(&m->side_stream, cudaStreamNonBlocking, prio_hi));
    LF_CUDA(cudaStreamCreateWithPriority(&m->early_stream, cudaStreamNonBlocking, prio_lo));
    LF_CUDA(cudaEventCreateWithFlags(&m->ev_fork, cudaEventDisableTiming));
    LF_CUDA(cudaEventCreateWithFlags(&m->ev_join, cudaEventDisableTiming));
    LF_CUDA(cudaEventCreateWithFlags(&m->ev_early_fork, cudaEventDisableTiming));
    LF_CUDA(cudaEventCreateWithFlags(&m->ev_early_join, cudaEventDisableTiming));
    LF_CHECK(m->iso_next.alloc(4));
    LF_CUDA(cudaMemsetAsync(m->iso_next.p, 0, 4 * sizeof(int), lf::stream()));
    // Kernels share an SM only if they agree on its L1 / shared-memory split: the soil stage asks for the maximum
    // shared-memory carve-out, so the kernels meant to run beside it ask for the same (they use no L1-resident reuse)
    const void *same_split[] = {(const void *)k_chan_isolated_ws<true>, (const void *)k_chan_isolated_ws<false>,
                                (const void *)k_chan_diagonal<true, false, false>, (const void *)k_chan_diagonal<false, false, false>,
                                (const void *)k_chan_diagonal<true, true, false>, (const void *)k_chan_diagonal<false, true, false>,
                                (const void *)k_chan_diagonal<true, false, true>, (const void *)k_chan_diagonal<false, false, true>,
                                (const void *)k_chan_diagonal<true, true, true>, (const void *)k_chan_diagonal<false, true, true>,
                                (const void *)k_chan_isolated_list<true>, (const void *)k_chan_isolated_list<false>};
    for (const void *fn : same_split)   // the wavefront on the side stream must be able to share SMs with the isolated-pixel kernel
        LF_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    return LF_OK;
}

// Starts the sub-steps of the non-channel isolated pixels of this model step on the low-priority stream (they need
// nothing the step computes).  Called at the top of lf_model_step; channel_stage() launches the machine-filling
// second drain of the same queue and joins.
int early_isolated(lf_model *m)
{
    m->early_in_flight = false;
    lf_graph *g = m->g_ch;
    const int iso_lo = (int)(m->n - g->n_isolated), iso_hi = (int)m->n;
    if (!m->overlap_isolated || iso_hi <= iso_lo) return LF_OK;
    cudaStream_t st = lf::stream();
    LF_CHECK(chan_streams(m));
    ChanPtrs C;
    LF_CHECK(chan_ptrs(m, C, nullptr, nullptr));
    LF_CUDA(cudaMemsetAsync(m->iso_next.p, 0, sizeof(int), st));
    LF_CUDA(cudaEventRecord(m->ev_early_fork, st));
    LF_CUDA(cudaStreamWaitEvent(m->early_stream, m->ev_early_fork, 0));
    const unsigned nchunks = lf::blocks_for(iso_hi - iso_lo, CH_THREADS);
    const unsigned grid = std::min<unsigned>(nchunks, (unsigned)(lf::sm_count() * std::max(1, m->early_blocks_per_sm)));
    if (m->quintic) k_chan_isolated_ws<true><<<grid, CH_THREADS, 0, m->early_stream>>>(C, iso_lo, iso_hi, m->iso_next.p);
    else k_chan_isolated_ws<false><<<grid, CH_THREADS, 0, m->early_stream>>>(C, iso_lo, iso_hi, m->iso_next.p);
    LF_LAUNCH_CHECK();
    LF_CUDA(cudaEventRecord(m->ev_early_join, m->early_stream));
    m->early_in_flight = true;
    return LF_OK;
}

int channel_stage(lf_model *m)
{
    cudaStream_t st = lf::stream();
    lf_graph *g = m->g_ch;
    ChanPtrs C;
    double *m32 = nullptr, *c2s = nullptr;
    LF_CHECK(chan_ptrs(m, C, &m32, &c2s));
    uint8_t *atlast = nullptr;
    LF_CHECK(flag_buf(m, "AtLastPointC", &atlast));
    // The connected network (many small, dependent launches) and the isolated pixels (one big launch) touch
    // disjoint pixels: the wavefront is issued on a side stream and overlaps the isolated kernel.
    LF_CHECK(chan_streams(m));
    cudaStream_t sw = m->side_stream;
    const std::vector<int32_t> &ls = g->h_level_start;
    int Lc = g->n_orders, S = C.S;
    int iso_lo = (int)(m->n - g->n_isolated), iso_hi = (int)m->n;
    if (iso_hi > iso_lo && m->iso_list_dirty) {   // isolated pixels that are channel pixels: compact list, rebuilt when the flags change
        if (m->iso_chan_list.n < (size_t)(iso_hi - iso_lo)) LF_CHECK(m->iso_chan_list.alloc(iso_hi - iso_lo));
        LF_CUDA(cudaMemsetAsync(m->iso_next.p + 1, 0, sizeof(int), st));
        k_chan_iso_collect<<<lf::blocks_for(iso_hi - iso_lo, 256), 256, 0, st>>>(C.isChan, iso_lo, iso_hi, m->iso_chan_list.p,
                                                                                  m->iso_next.p + 1);
        LF_LAUNCH_CHECK();
        LF_CUDA(cudaMemcpyAsync(&m->iso_chan_count, m->iso_next.p + 1, sizeof(int), cudaMemcpyDeviceToHost, st));
        LF_CUDA(cudaStreamSynchronize(st));
        m->iso_list_dirty = false;
    }
    if (!m->early_in_flight && iso_hi > iso_lo) LF_CUDA(cudaMemsetAsync(m->iso_next.p, 0, sizeof(int), st));
    LF_CUDA(cudaEventRecord(m->ev_fork, st));
    LF_CUDA(cudaStreamWaitEvent(sw, m->ev_fork, 0));
    if (iso_hi > iso_lo) {
        // second drain of the chunk queue.  iso_blocks_per_sm > 0: a persistent grid of that many blocks per SM -- 8 fill the
        // register file and the wavefront on the side stream then only starts once the queue is drained; 6 (default) leave
        // room for two wavefront blocks per SM, so the two overlap (profiles/r02_channel_overlap_experiment.txt); 0: one
        // block per chunk.
        const unsigned nchunks = lf::blocks_for(iso_hi - iso_lo, CH_THREADS);
        const unsigned grid = m->iso_blocks_per_sm > 0 ? std::min<unsigned>(nchunks, (unsigned)(lf::sm_count() * m->iso_blocks_per_sm))
                                                       : nchunks;
        if (m->quintic) k_chan_isolated_ws<true><<<grid, CH_THREADS, 0, st>>>(C, iso_lo, iso_hi, m->iso_next.p);
        else k_chan_isolated_ws<false><<<grid, CH_THREADS, 0, st>>>(C, iso_lo, iso_hi, m->iso_next.p);
        LF_LAUNCH_CHECK();
        if (m->iso_chan_count > 0) {
            const unsigned gl = lf::blocks_for(m->iso_chan_count, CH_THREADS);
            if (m->quintic) k_chan_isolated_list<true><<<gl, CH_THREADS, 0, st>>>(C, m->iso_chan_list.p, m->iso_chan_count);
            else k_chan_isolated_list<false><<<gl, CH_THREADS, 0, st>>>(C, m->iso_chan_list.p, m->iso_chan_count);
            LF_LAUNCH_CHECK();
        }
    }
    // the connected network: space-time wavefront over (level, sub-step)
    auto level_end = [&](int l) { return l == Lc - 1 ? iso_lo : ls[l + 1]; };
    auto width = [&](int d) {
        int lo_lev = d - S + 1 > 0 ? d - S + 1 : 0;
        int hi_lev = d < Lc - 1 ? d : Lc - 1;
        return level_end(hi_lev) - ls[lo_lev];
    };
    const int32_t *lsd = g->level_start.p;
    auto wavefront = [&]() -> int {
        for (int d = 0; d < Lc + S - 1; ++d) {
            if (m->narrow_runs && width(d) <= CH_RUN_THREADS) {   // a run of narrow diagonals: one single-block launch
                int e = d + 1;
                while (e < Lc + S - 1 && width(e) <= CH_RUN_THREADS) ++e;
                if (e - d >= 3) {
#define LF_CHAN_RUN(QZ_, HX_, HS_) k_chan_narrow_run<QZ_, HX_, HS_><<<1, CH_RUN_THREADS, 0, sw>>>(C, d, e, Lc, lsd, iso_lo)
                    if (C.sid && C.X.xslot) {
                        if (m->quintic) LF_CHAN_RUN(true, true, true);
                        else LF_CHAN_RUN(false, true, true);
                    } else if (C.sid) {
                        if (m->quintic) LF_CHAN_RUN(true, false, true);
                        else LF_CHAN_RUN(false, false, true);
                    } else if (C.X.xslot) {
                        if (m->quintic) LF_CHAN_RUN(true, true, false);
                        else LF_CHAN_RUN(false, true, false);
                    } else {
                        if (m->quintic) LF_CHAN_RUN(true, false, false);
                        else LF_CHAN_RUN(false, false, false);
                    }
#undef LF_CHAN_RUN
                    LF_LAUNCH_CHECK();
                    d = e - 1;
                    continue;
                }
            }
            int lo_lev = d - S + 1 > 0 ? d - S + 1 : 0;
            int hi_lev = d < Lc - 1 ? d : Lc - 1;
            int lo = ls[lo_lev], hi = level_end(hi_lev);
            if (hi <= lo) continue;
            const unsigned gb = lf::blocks_for(hi - lo, CH_THREADS);
#define LF_CHAN_LAUNCH(QZ_, HX_, HS_) k_chan_diagonal<QZ_, HX_, HS_><<<gb, CH_THREADS, 0, sw>>>(C, lo, hi, d)
            if (C.sid && C.X.xslot) {
                if (m->quintic) LF_CHAN_LAUNCH(true, true, true);
                else LF_CHAN_LAUNCH(false, true, true);
            } else if (C.sid) {
                if (m->quintic) LF_CHAN_LAUNCH(true, false, true);
                else LF_CHAN_LAUNCH(false, false, true);
            } else if (C.X.xslot) {
                if (m->quintic) LF_CHAN_LAUNCH(true, true, false);
                else LF_CHAN_LAUNCH(false, true, false);
            } else {
                if (m->quintic) LF_CHAN_LAUNCH(true, false, false);
                else LF_CHAN_LAUNCH(false, false, false);
            }
#undef LF_CHAN_LAUNCH
            LF_LAUNCH_CHECK();
        }
        return LF_OK;
    };
    if (m->use_graphs && Lc + S > 6) LF_CHECK(lf::run_captured(m->graphs_ch, &C, sizeof(C), sw, wavefront));
    else LF_CHECK(wavefront());
    LF_CUDA(cudaEventRecord(m->ev_join, sw));
    LF_CUDA(cudaStreamWaitEvent(st, m->ev_join, 0));
    if (m->xchg) LF_CHECK(lf_xchg_end(m->xchg));
    if (m->early_in_flight) {
        LF_CUDA(cudaStreamWaitEvent(st, m->ev_early_join, 0));
        m->early_in_flight = false;
    }
    FIELD(chm3, "ChanM3");
    FIELD(tcs, "TotalCrossSectionArea");
    FIELD(sdis, "sumDis");
    FIELD(qavg, "ChanQAvg");
    FIELD(dout, "DischargeM3Out");
    double *fv = nullptr, *td = nullptr, *pa = nullptr, *cumq = nullptr;
    if (m->accumulate_discharge) {
        FIELD(cq_, "CumQ");
        cumq = cq_;
    }
    if (m->cfg.diagnostics) {
        FIELD(fv_, "FlowVelocity");
        FIELD(td_, "TravelDistance");
        FIELD(pa_, "PixelArea");
        fv = fv_;
        td = td_;
        pa = pa_;
    }
    k_chan_post<<<lf::blocks_for(m->n, 256), 256, 0, st>>>((int)m->n, C.split, m->quintic ? 1 : 0, S, m->cfg.DtSec, C.M3, m32, c2s, C.L,
                                                           C.sumDis, C.ChanQ, C.Qk, atlast, pa, chm3, tcs, sdis, qavg, dout, fv, td,
                                                           m->nancheck ? m->iso_next.p + 2 : nullptr, cumq);
    LF_LAUNCH_CHECK();
    return LF_OK;
}

}  // namespace

extern "C" {

int lf_model_create(const lf_model_config *cfg, const uint8_t *land_mask, const double *ldd_to_chan,
                    const double *ldd_kinematic, lf_model **out)
{
    if (!cfg || !land_mask || !ldd_to_chan || !ldd_kinematic || !out) {
        lf::set_error("lf_model_create: null pointer");
        return LF_ERR_INVALID;
    }
    if (!(cfg->DtSec > 0) || !(cfg->Beta > 0) || cfg->NoRoutSteps < 1 || !(cfg->PixelLength > 0) ||
        !(cfg->CourantCrit > 0)) {
        lf::set_error("lf_model_create: DtSec, Beta, PixelLength, CourantCrit must be > 0 and NoRoutSteps >= 1");
        return LF_ERR_INVALID;
    }
    LF_CHECK(lf::ensure_device());
    *out = nullptr;
    std::unique_ptr<lf_model> m(new lf_model());
    m->cfg = *cfg;
    m->DtDay = cfg->DtSec / 86400.;
    m->DtRouting = cfg->DtSec / cfg->NoRoutSteps;
    m->quintic = lfkw::make_params(cfg->Beta).quintic != 0;
    LF_CHECK(lf_ldd_build(ldd_to_chan, land_mask, cfg->rows, cfg->cols, &m->g_of));
    LF_CHECK(lf_ldd_build(ldd_kinematic, land_mask, cfg->rows, cfg->cols, &m->g_ch));
    m->n = m->g_of->n;
    LF_CHECK(m->stage.alloc((size_t)3 * m->n));
    LF_CHECK(m->stage_u8.alloc(m->n));
    LF_CHECK(m->soil_to_chan.alloc(m->n));
    LF_CHECK(m->flag.alloc(1));
    k_soil_to_chan<<<lf::blocks_for(m->n, 256), 256, 0, lf::stream()>>>(m->g_of->pix_of_pos.p, m->g_ch->pos_of_pix.p,
                                                                        m->soil_to_chan.p, m->n);
    LF_LAUNCH_CHECK();
    LF_CUDA(cudaStreamSynchronize(lf::stream()));
    *out = m.release();
    return LF_OK;
}

int lf_model_create_from_graphs(const lf_model_config *cfg, lf_graph *g_overland, lf_graph *g_channel, lf_model **out)
{
    if (!cfg || !g_overland || !g_channel || !out) {
        lf::set_error("lf_model_create_from_graphs: null pointer");
        return LF_ERR_INVALID;
    }
    if (g_overland->n != g_channel->n) {
        lf::set_error("lf_model_create_from_graphs: the two graphs hold different pixel sets");
        return LF_ERR_INVALID;
    }
    if (!(cfg->DtSec > 0) || !(cfg->Beta > 0) || cfg->NoRoutSteps < 1 || !(cfg->PixelLength > 0) ||
        !(cfg->CourantCrit > 0)) {
        lf::set_error("lf_model_create_from_graphs: DtSec, Beta, PixelLength, CourantCrit must be > 0 and NoRoutSteps >= 1");
        return LF_ERR_INVALID;
    }
    LF_CHECK(lf::ensure_device());
    *out = nullptr;
    std::unique_ptr<lf_model> m(new lf_model());
    m->cfg = *cfg;
    m->DtDay = cfg->DtSec / 86400.;
    m->DtRouting = cfg->DtSec / cfg->NoRoutSteps;
    m->quintic = lfkw::make_params(cfg->Beta).quintic != 0;
    m->n = g_overland->n;
    LF_CHECK(m->stage.alloc((size_t)3 * m->n));
    LF_CHECK(m->stage_u8.alloc(m->n));
    LF_CHECK(m->soil_to_chan.alloc(m->n));
    LF_CHECK(m->flag.alloc(1));
    k_soil_to_chan<<<lf::blocks_for(m->n, 256), 256, 0, lf::stream()>>>(g_overland->pix_of_pos.p, g_channel->pos_of_pix.p,
                                                                        m->soil_to_chan.p, m->n);
    LF_LAUNCH_CHECK();
    LF_CUDA(cudaStreamSynchronize(lf::stream()));
    m->g_of = g_overland;   // owned by the model from here on
    m->g_ch = g_channel;
    *out = m.release();
    return LF_OK;
}

int lf_model_set_exchange(lf_model *m, lf_xchg *x, int32_t which, const int32_t *xslot, int32_t n_export,
                          const int32_t *export_peer, const int64_t *export_offset, const int64_t *export_parity_stride,
                          int32_t n_import, int64_t import_offset)
{
    if (!m || !x || !xslot || (which != 0 && which != 1) || n_export < 0 || n_import < 0 || import_offset < 0 ||
        (n_export > 0 && (!export_peer || !export_offset || !export_parity_stride))) {
        lf::set_error("lf_model_set_exchange: bad arguments");
        return LF_ERR_INVALID;
    }
    if (m->xchg && m->xchg != x) {
        lf::set_error("lf_model_set_exchange: both graphs must use the same exchange region");
        return LF_ERR_STATE;
    }
    LF_CHECK(lf::ensure_device());
    cudaStream_t st = lf::stream();
    lf_model::XSide &xs = which == 0 ? m->xs_of : m->xs_ch;
    lf_graph *g = which == 0 ? m->g_of : m->g_ch;
    const int cap = which == 0 ? 1 : m->cfg.NoRoutSteps, nsec = which == 0 ? 3 : (m->cfg.SplitRouting ? 2 : 1);
    lf::DevBuf<int32_t> tmp;
    LF_CHECK(tmp.alloc(m->n));
    LF_CHECK(xs.xslot.alloc(m->n));
    LF_CUDA(cudaMemcpyAsync(tmp.p, xslot, m->n * sizeof(int32_t), cudaMemcpyDefault, st));
    k_i32_rows_to_pos<<<lf::blocks_for(m->n, 256), 256, 0, st>>>(tmp.p, xs.xslot.p, g->pix_of_pos.p, m->n);
    LF_LAUNCH_CHECK();
    std::vector<double *> dst(std::max(n_export, 1), nullptr);
    std::vector<long long> stride(std::max(n_export, 1), 0);
    for (int k = 0; k < n_export; ++k) {
        uint64_t base = 0;
        LF_CHECK(lf_xchg_peer_base(x, export_peer[k], &base));
        dst[k] = (double *)(uintptr_t)(base + lfx::HEADER_BYTES) + export_offset[k];
        stride[k] = export_parity_stride[k];
    }
    LF_CHECK(xs.exp_dst.alloc(dst.size()));
    LF_CHECK(xs.exp_stride.alloc(stride.size()));
    LF_CUDA(cudaMemcpyAsync(xs.exp_dst.p, dst.data(), dst.size() * sizeof(double *), cudaMemcpyHostToDevice, st));
    LF_CUDA(cudaMemcpyAsync(xs.exp_stride.p, stride.data(), stride.size() * sizeof(long long), cudaMemcpyHostToDevice, st));
    LF_CUDA(cudaStreamSynchronize(st));
    uint64_t own = 0;
    LF_CHECK(lf_xchg_peer_base(x, -1, &own));
    xs.imp = (double *)(uintptr_t)(own + lfx::HEADER_BYTES) + import_offset;
    xs.imp_parity_stride = (long long)n_import * nsec * cap;
    xs.n_export = n_export;
    xs.n_import = n_import;
    m->xchg = x;
    return LF_OK;
}

int lf_model_info(const lf_model *m, int64_t *n_pixels, int64_t *levels_overland, int64_t *levels_channel,
                  int64_t *isolated_channel_pixels, int64_t *device_bytes)
{
    if (!m) {
        lf::set_error("lf_model_info: null model");
        return LF_ERR_INVALID;
    }
    if (n_pixels) *n_pixels = m->n;
    if (levels_overland) *levels_overland = m->g_of->n_orders;
    if (levels_channel) *levels_channel = m->g_ch->n_orders;
    if (isolated_channel_pixels) *isolated_channel_pixels = m->g_ch->n_isolated;
    if (device_bytes) *device_bytes = m->bytes;
    return LF_OK;
}

int lf_model_set(lf_model *m, const char *name, const double *values, int64_t count)
{
    if (!m || !name || !values) {
        lf::set_error("lf_model_set: null pointer");
        return LF_ERR_INVALID;
    }
    LF_CHECK(lf::ensure_device());
    Field *f = nullptr;
    LF_CHECK(field(m, name, &f));
    if (count != (int64_t)f->rows * m->n) {
        lf::set_error("lf_model_set(%s): expected %lld values, got %lld", name, (long long)f->rows * m->n, (long long)count);
        return LF_ERR_INVALID;
    }
    cudaStream_t st = lf::stream();
    const int32_t *pop = f->order == SOIL ? m->g_of->pix_of_pos.p : m->g_ch->pix_of_pos.p;
    // host data goes through the staging buffer; device data is gathered straight from the caller's buffer
    const bool values_on_device = lf::is_device_ptr(values);
    const double *from = values;
    if (!values_on_device || f->landuse) {
        LF_CUDA(cudaMemcpyAsync(m->stage.p, values, count * sizeof(double), cudaMemcpyDefault, st));
        from = m->stage.p;
    }
    k_rows_to_pos<<<lf::blocks_for((m->n + PERM_U - 1) / PERM_U, 256), 256, 0, st>>>(from, f->buf.p, pop, m->n, f->rows);
    LF_LAUNCH_CHECK();
    if (f->as_z) {
        k_q_to_z<<<lf::blocks_for(m->n, 256), 256, 0, st>>>(f->buf.p, m->n);
        LF_LAUNCH_CHECK();
    }
    if (f->landuse) {
        // rows that repeat an earlier land use share its storage row (Irrigated == Rainfed without a third map,
        // Lisflood_initial.py:371-391): the kernel then re-reads cached lines instead of new HBM bytes
        const bool on_device = lf::is_device_ptr(values);
        for (int r = 0; r < 3; ++r) {
            f->row_index[r] = r;
            for (int q = 0; q < r; ++q) {
                bool same;
                if (on_device) {
                    int h = 0;
                    LF_CUDA(cudaMemsetAsync(m->flag.p, 0, sizeof(int), st));
                    k_rows_differ<<<lf::blocks_for(m->n, 256), 256, 0, st>>>(m->stage.p + (int64_t)r * m->n,
                                                                             m->stage.p + (int64_t)q * m->n, m->n, m->flag.p);
                    LF_LAUNCH_CHECK();
                    LF_CUDA(cudaMemcpyAsync(&h, m->flag.p, sizeof(int), cudaMemcpyDeviceToHost, st));
                    LF_CUDA(cudaStreamSynchronize(st));
                    same = h == 0;
                } else {
                    same = memcmp(values + (int64_t)r * m->n, values + (int64_t)q * m->n, m->n * sizeof(double)) == 0;
                }
                if (same) {
                    f->row_index[r] = f->row_index[q];
                    break;
                }
            }
        }
    }
    if (strcmp(name, "OFAlpha") == 0 || strcmp(name, "ChannelAlpha") == 0 || strcmp(name, "ChannelAlpha2") == 0 ||
        strcmp(name, "ChanLength") == 0 || strcmp(name, "Chan2M3Start") == 0)
        m->params_dirty = true;
    if (strcmp(name, "ChanQ") == 0) m->st.cq_dirty = true;
    LF_CUDA(cudaStreamSynchronize(st));
    return LF_OK;
}

int lf_model_set_async(lf_model *m, const char *name, const double *values, int64_t count)
{
    if (!m || !name || !values) {
        lf::set_error("lf_model_set_async: null pointer");
        return LF_ERR_INVALID;
    }
    LF_CHECK(lf::ensure_device());
    Field *f = nullptr;
    LF_CHECK(field(m, name, &f));
    if (count != (int64_t)f->rows * m->n) {
        lf::set_error("lf_model_set_async(%s): expected %lld values, got %lld", name, (long long)f->rows * m->n,
                      (long long)count);
        return LF_ERR_INVALID;
    }
    if (f->landuse) {
        lf::set_error("lf_model_set_async(%s): land-use parameters are set with lf_model_set", name);
        return LF_ERR_INVALID;
    }
    if (!m->copy_stream) LF_CUDA(cudaStreamCreateWithFlags(&m->copy_stream, cudaStreamNonBlocking));
    auto it = m->async_stage.find(name);
    if (it == m->async_stage.end()) {
        std::unique_ptr<lf::DevBuf<double>> b(new lf::DevBuf<double>());
        LF_CHECK(b->alloc(count));
        m->bytes += count * 8;
        it = m->async_stage.emplace(name, std::move(b)).first;
        cudaEvent_t e1, e2;
        LF_CUDA(cudaEventCreateWithFlags(&e1, cudaEventDisableTiming));
        LF_CUDA(cudaEventCreateWithFlags(&e2, cudaEventDisableTiming));
        m->async_copied[name] = e1;
        m->async_consumed[name] = e2;
        LF_CUDA(cudaEventRecord(e2, lf::stream()));
    }
    cudaStream_t st = lf::stream();
    double *stg = it->second->p;
    // the staging buffer is free again once the previous translation kernel has run
    LF_CUDA(cudaStreamWaitEvent(m->copy_stream, m->async_consumed[name], 0));
    LF_CUDA(cudaMemcpyAsync(stg, values, count * sizeof(double), cudaMemcpyDefault, m->copy_stream));
    LF_CUDA(cudaEventRecord(m->async_copied[name], m->copy_stream));
    LF_CUDA(cudaStreamWaitEvent(st, m->async_copied[name], 0));
    const int32_t *pop = f->order == SOIL ? m->g_of->pix_of_pos.p : m->g_ch->pix_of_pos.p;
    k_rows_to_pos<<<lf::blocks_for((m->n + PERM_U - 1) / PERM_U, 256), 256, 0, st>>>(stg, f->buf.p, pop, m->n, f->rows);
    LF_LAUNCH_CHECK();
    if (f->as_z) {
        k_q_to_z<<<lf::blocks_for(m->n, 256), 256, 0, st>>>(f->buf.p, m->n);
        LF_LAUNCH_CHECK();
    }
    LF_CUDA(cudaEventRecord(m->async_consumed[name], st));
    if (strcmp(name, "OFAlpha") == 0 || strcmp(name, "ChannelAlpha") == 0 || strcmp(name, "ChannelAlpha2") == 0 ||
        strcmp(name, "ChanLength") == 0 || strcmp(name, "Chan2M3Start") == 0)
        m->params_dirty = true;
    return LF_OK;
}

int lf_model_get(lf_model *m, const char *name, double *values, int64_t count)
{
    if (!m || !name || !values) {
        lf::set_error("lf_model_get: null pointer");
        return LF_ERR_INVALID;
    }
    LF_CHECK(lf::ensure_device());
    Field *f = nullptr;
    LF_CHECK(field(m, name, &f));
    if (count != (int64_t)f->rows * m->n) {
        lf::set_error("lf_model_get(%s): expected %lld values, got %lld", name, (long long)f->rows * m->n, (long long)count);
        return LF_ERR_INVALID;
    }
    cudaStream_t st = lf::stream();
    const int32_t *pos = f->order == SOIL ? m->g_of->pos_of_pix.p : m->g_ch->pos_of_pix.p;
    k_rows_to_pix<<<lf::blocks_for(m->n, 256), 256, 0, st>>>(f->buf.p, m->stage.p, pos, m->n, f->rows);
    LF_LAUNCH_CHECK();
    if (f->as_z) {
        k_z_to_q<<<lf::blocks_for(m->n, 256), 256, 0, st>>>(m->stage.p, m->n);
        LF_LAUNCH_CHECK();
    }
    LF_CUDA(cudaMemcpyAsync(values, m->stage.p, count * sizeof(double), cudaMemcpyDefault, st));
    LF_CUDA(cudaStreamSynchronize(st));
    return LF_OK;
}

// f32: narrow to float32 on the device before the copy (half the bytes over the host link)
static int get_async_impl(lf_model *m, const char *name, void *values, int64_t count, bool f32)
{
    LF_CHECK(lf::ensure_device());
    Field *f = nullptr;
    LF_CHECK(field(m, name, &f));
    if (count != (int64_t)f->rows * m->n) {
        lf::set_error("lf_model_get_async(%s): expected %lld values, got %lld", name, (long long)f->rows * m->n, (long long)count);
        return LF_ERR_INVALID;
    }
    cudaStream_t st = lf::stream();
    if (!m->out_stream) LF_CUDA(cudaStreamCreateWithFlags(&m->out_stream, cudaStreamNonBlocking));
    const std::string key = f32 ? std::string(name) + "#f32" : std::string(name);
    auto it = m->out_stage.find(key);
    if (it == m->out_stage.end()) {
        std::unique_ptr<lf::DevBuf<double>> b(new lf::DevBuf<double>());
        const int64_t words = f32 ? (count + 1) / 2 : count;
        LF_CHECK(b->alloc(words));
        m->bytes += words * 8;
        it = m->out_stage.emplace(key, std::move(b)).first;
        cudaEvent_t e1, e2;
        LF_CUDA(cudaEventCreateWithFlags(&e1, cudaEventDisableTiming));
        LF_CUDA(cudaEventCreateWithFlags(&e2, cudaEventDisableTiming));
        m->out_ready[key] = e1;
        m->out_copied[key] = e2;
        LF_CUDA(cudaEventRecord(e2, m->out_stream));
    }
    double *stg = it->second->p;
    LF_CUDA(cudaStreamWaitEvent(st, m->out_copied[key], 0));   // the previous copy out of this staging buffer is done
    const int32_t *pos = f->order == SOIL ? m->g_of->pos_of_pix.p : m->g_ch->pos_of_pix.p;
    if (f32) {
        k_rows_to_pix_f32<<<lf::blocks_for(m->n, 256), 256, 0, st>>>(f->buf.p, (float *)stg, pos, m->n, f->rows, f->as_z ? 1 : 0);
        LF_LAUNCH_CHECK();
    } else {
        k_rows_to_pix<<<lf::blocks_for(m->n, 256), 256, 0, st>>>(f->buf.p, stg, pos, m->n, f->rows);
        LF_LAUNCH_CHECK();
        if (f->as_z) {
            k_z_to_q<<<lf::blocks_for(m->n, 256), 256, 0, st>>>(stg, m->n);
            LF_LAUNCH_CHECK();
        }
    }
    LF_CUDA(cudaEventRecord(m->out_ready[key], st));
    LF_CUDA(cudaStreamWaitEvent(m->out_stream, m->out_ready[key], 0));
    LF_CUDA(cudaMemcpyAsync(values, stg, count * (f32 ? sizeof(float) : sizeof(double)), cudaMemcpyDefault, m->out_stream));
    LF_CUDA(cudaEventRecord(m->out_copied[key], m->out_stream));
    return LF_OK;
}

int lf_model_get_async(lf_model *m, const char *name, double *values, int64_t count)
{
    if (!m || !name || !values) {
        lf::set_error("lf_model_get_async: null pointer");
        return LF_ERR_INVALID;
    }
    return get_async_impl(m, name, values, count, false);
}

int lf_model_get_async_f32(lf_model *m, const char *name, float *values, int64_t count)
{
    if (!m || !name || !values) {
        lf::set_error("lf_model_get_async_f32: null pointer");
        return LF_ERR_INVALID;
    }
    return get_async_impl(m, name, values, count, true);
}

int lf_model_wait_outputs(lf_model *m)
{
    if (!m) {
        lf::set_error("lf_model_wait_outputs: null model");
        return LF_ERR_INVALID;
    }
    LF_CHECK(lf::ensure_device());
    if (m->out_stream) LF_CUDA(cudaStreamSynchronize(m->out_stream));
    return LF_OK;
}

int lf_model_set_flags(lf_model *m, const char *name, const uint8_t *values, int64_t count)
{
    if (!m || !name || !values) {
        lf::set_error("lf_model_set_flags: null pointer");
        return LF_ERR_INVALID;
    }
    Order order;
    if (strcmp(name, "isFrozenSoil") == 0 || strcmp(name, "IsChannel") == 0) order = SOIL;
    else if (strcmp(name, "IsChannelKinematic") == 0 || strcmp(name, "AtLastPointC") == 0) order = CHAN;
    else {
        lf::set_error("lf_model_set_flags: unknown flag map '%s'", name);
        return LF_ERR_INVALID;
    }
    if (count != m->n) {
        lf::set_error("lf_model_set_flags(%s): expected %lld values", name, (long long)m->n);
        return LF_ERR_INVALID;
    }
    LF_CHECK(lf::ensure_device());
    cudaStream_t st = lf::stream();
    uint8_t *dst = nullptr;
    LF_CHECK(flag_buf(m, name, &dst));
    const int32_t *pop = order == SOIL ? m->g_of->pix_of_pos.p : m->g_ch->pix_of_pos.p;
    if (strcmp(name, "IsChannelKinematic") == 0) m->iso_list_dirty = true;
    LF_CUDA(cudaMemcpyAsync(m->stage_u8.p, values, count, cudaMemcpyDefault, st));
    k_u8_to_pos<<<lf::blocks_for((m->n + PERM_U - 1) / PERM_U, 256), 256, 0, st>>>(m->stage_u8.p, dst, pop, m->n);
    LF_LAUNCH_CHECK();
    LF_CUDA(cudaStreamSynchronize(st));
    return LF_OK;
}

int lf_model_soil(lf_model *m)
{
    if (!m) {
        lf::set_error("lf_model_soil: null model");
        return LF_ERR_INVALID;
    }
    LF_CHECK(lf::ensure_device());
    return soil_stage(m);
}
int lf_model_surface_routing(lf_model *m)
{
    if (!m) {
        lf::set_error("lf_model_surface_routing: null model");
        return LF_ERR_INVALID;
    }
    LF_CHECK(lf::ensure_device());
    return surface_stage(m);
}
int lf_model_channel(lf_model *m)
{
    if (!m) {
        lf::set_error("lf_model_channel: null model");
        return LF_ERR_INVALID;
    }
    LF_CHECK(lf::ensure_device());
    LF_CHECK(channel_stage(m));
    m->steps += 1;
    return LF_OK;
}
static int mark(lf_model *m)
{
    size_t k = m->ev_used.size();
    if (k >= 4096) return LF_OK;  // stop recording, keep running
    if (k >= m->ev_pool.size()) {
        cudaEvent_t e;
        LF_CUDA(cudaEventCreate(&e));
        m->ev_pool.push_back(e);
    }
    LF_CUDA(cudaEventRecord(m->ev_pool[k], lf::stream()));
    m->ev_used.push_back((int)k);
    return LF_OK;
}

int lf_model_step(lf_model *m)
{
    if (!m) {
        lf::set_error("lf_model_step: null model");
        return LF_ERR_INVALID;
    }
    LF_CHECK(lf::ensure_device());
    LF_CHECK(mark(m));
    LF_CHECK(early_isolated(m));
    LF_CHECK(soil_stage(m));
    LF_CHECK(mark(m));
    LF_CHECK(surface_stage(m));
    LF_CHECK(mark(m));
    LF_CHECK(channel_stage(m));
    LF_CHECK(mark(m));
    m->steps += 1;
    return LF_OK;
}

int lf_model_stage_times(lf_model *m, int reset, double *soil_ms, double *overland_ms, double *channel_ms,
                         int64_t *steps)
{
    if (!m) {
        lf::set_error("lf_model_stage_times: null model");
        return LF_ERR_INVALID;
    }
    LF_CHECK(lf::ensure_device());
    LF_CUDA(cudaStreamSynchronize(lf::stream()));
    for (size_t k = 0; k + 3 < m->ev_used.size(); k += 4) {
        float a = 0, b = 0, c = 0;
        LF_CUDA(cudaEventElapsedTime(&a, m->ev_pool[k], m->ev_pool[k + 1]));
        LF_CUDA(cudaEventElapsedTime(&b, m->ev_pool[k + 1], m->ev_pool[k + 2]));
        LF_CUDA(cudaEventElapsedTime(&c, m->ev_pool[k + 2], m->ev_pool[k + 3]));
        m->t_soil += a;
        m->t_of += b;
        m->t_chan += c;
        m->t_steps += 1;
    }
    m->ev_used.clear();
    if (soil_ms) *soil_ms = m->t_soil;
    if (overland_ms) *overland_ms = m->t_of;
    if (channel_ms) *channel_ms = m->t_chan;
    if (steps) *steps = m->t_steps;
    if (reset) {
        m->t_soil = m->t_of = m->t_chan = 0;
        m->t_steps = 0;
    }
    return LF_OK;
}

int lf_model_soil_stats(lf_model *m, int enable_timing, int64_t *deferred_columns, double *kernel_ms)
{
    if (!m) {
        lf::set_error("lf_model_soil_stats: null model");
        return LF_ERR_INVALID;
    }
    LF_CHECK(lf::ensure_device());
    LF_CUDA(cudaStreamSynchronize(lf::stream()));
    if (deferred_columns) {
        int32_t h[lfsoil::NBUCKET] = {0};
        if (m->soil_list_cnt.p)
            LF_CUDA(cudaMemcpy(h, m->soil_list_cnt.p, sizeof(h), cudaMemcpyDeviceToHost));
        for (int b = 0; b < lfsoil::NBUCKET; ++b) deferred_columns[b] = h[b];
    }
    if (kernel_ms) {
        for (int k = 0; k < 8; ++k) kernel_ms[k] = 0.;
        if (m->soil_profile && m->soil_ev.size() == 9)
            for (int k = 0; k < 8; ++k) {
                float t = 0.f;
                if (cudaEventElapsedTime(&t, m->soil_ev[k], m->soil_ev[k + 1]) == cudaSuccess) kernel_ms[k] = t;
                else cudaGetLastError();
            }
    }
    if (enable_timing && m->soil_ev.empty()) {
        for (int k = 0; k < 9; ++k) {
            cudaEvent_t e;
            LF_CUDA(cudaEventCreate(&e));
            m->soil_ev.push_back(e);
        }
    }
    m->soil_profile = enable_timing != 0;
    return LF_OK;
}


static const char *const FEED_PARAMS[] = {"PrScaling", "CalEvaporation", "DeltaTSnow", "SnowSeason", "TempSnow", "SnowFactor",
                                          "SnowMeltCoef", "TempMelt", "lat_rad", "Kfrost", "Afrost", "FrostIndexThreshold",
                                          "SnowWaterEquivalent", "kgb"};

int lf_model_set_scalar(lf_model *m, const char *name, double value)
{
    if (!m || !name) {
        lf::set_error("lf_model_set_scalar: null pointer");
        return LF_ERR_INVALID;
    }
    for (const char *nm : FEED_PARAMS)
        if (strcmp(nm, name) == 0) {
            m->scalars[name] = value;
            m->fields.erase(name);   // a scalar replaces a map set earlier
            return LF_OK;
        }
    lf::set_error("lf_model_set_scalar: '%s' is not a scalar-or-map parameter", name);
    return LF_ERR_INVALID;
}

static int scalar_or_map(lf_model *m, const char *name, ScalarOrMap *out)
{
    auto it = m->fields.find(name);
    if (it != m->fields.end()) {
        out->map = it->second->buf.p;
        out->value = 0.;
        return LF_OK;
    }
    auto is = m->scalars.find(name);
    if (is == m->scalars.end()) {
        lf::set_error("parameter '%s' is not set (lf_model_set for a map, lf_model_set_scalar for a scalar)", name);
        return LF_ERR_STATE;
    }
    out->map = nullptr;
    out->value = is->second;
    return LF_OK;
}

// dtype: 0 float64, 1 float32, 2 int16 with scale / offset (one pair per map)
static int feed_impl(lf_model *m, const void *precipitation, const void *tavg, const void *et0, const void *e0, int32_t dtype,
                     const double *scale, const double *offset, int32_t decode_f32, double snowmelt_coeff,
                     double ice_melt_coeff_north, double ice_melt_coeff_south, int32_t async)
{
    LF_CHECK(lf::ensure_device());
    cudaStream_t st = lf::stream();
    const size_t esz = dtype == 2 ? 2 : dtype == 1 ? 4 : 8;
    const size_t map_bytes = ((size_t)m->n * esz + 255) / 256 * 256;
    const void *src[4] = {precipitation, tavg, et0, e0};
    const void *dev[4];
    const bool on_device = lf::is_device_ptr(precipitation);
    if (on_device) {
        for (int k = 0; k < 4; ++k) dev[k] = src[k];
    } else {
        const int turn = m->raw_turn;
        m->raw_turn ^= 1;
        if (m->raw_stage[turn].n < 4 * map_bytes) {
            LF_CHECK(m->raw_stage[turn].alloc(4 * map_bytes));
            m->bytes += 4 * map_bytes;
        }
        if (!m->raw_copied[turn]) {
            LF_CUDA(cudaEventCreateWithFlags(&m->raw_copied[turn], cudaEventDisableTiming));
            LF_CUDA(cudaEventCreateWithFlags(&m->raw_consumed[turn], cudaEventDisableTiming));
            LF_CUDA(cudaEventRecord(m->raw_consumed[turn], st));
        }
        cudaStream_t cs = st;
        if (async) {
            if (!m->copy_stream) LF_CUDA(cudaStreamCreateWithFlags(&m->copy_stream, cudaStreamNonBlocking));
            cs = m->copy_stream;
            LF_CUDA(cudaStreamWaitEvent(cs, m->raw_consumed[turn], 0));   // the staging buffer is free again
        }
        for (int k = 0; k < 4; ++k) {
            void *d = m->raw_stage[turn].p + k * map_bytes;
            LF_CUDA(cudaMemcpyAsync(d, src[k], (size_t)m->n * esz, cudaMemcpyDefault, cs));
            dev[k] = d;
        }
        if (async) {
            LF_CUDA(cudaEventRecord(m->raw_copied[turn], cs));
            LF_CUDA(cudaStreamWaitEvent(st, m->raw_copied[turn], 0));
        }
    }
    FeedPtrs F;
    memset(&F, 0, sizeof(F));
    F.n = m->n;
    F.pix_of_pos = m->g_of->pix_of_pos.p;
    F.prec = dev[0];
    F.tavg = dev[1];
    F.et0 = dev[2];
    F.e0 = dev[3];
    for (int k = 0; k < 4; ++k) {
        F.scale[k] = scale ? scale[k] : 1.;
        F.offset[k] = offset ? offset[k] : 0.;
    }
    F.decode_f32 = decode_f32 ? 1 : 0;
    LF_CHECK(scalar_or_map(m, "PrScaling", &F.PrScaling));
    LF_CHECK(scalar_or_map(m, "CalEvaporation", &F.CalEvaporation));
    LF_CHECK(scalar_or_map(m, "DeltaTSnow", &F.DeltaTSnow));
    LF_CHECK(scalar_or_map(m, "SnowSeason", &F.SnowSeason));
    LF_CHECK(scalar_or_map(m, "TempSnow", &F.TempSnow));
    LF_CHECK(scalar_or_map(m, "SnowFactor", &F.SnowFactor));
    LF_CHECK(scalar_or_map(m, "SnowMeltCoef", &F.SnowMeltCoef));
    LF_CHECK(scalar_or_map(m, "TempMelt", &F.TempMelt));
    LF_CHECK(scalar_or_map(m, "lat_rad", &F.lat_rad));
    LF_CHECK(scalar_or_map(m, "Kfrost", &F.Kfrost));
    LF_CHECK(scalar_or_map(m, "Afrost", &F.Afrost));
    LF_CHECK(scalar_or_map(m, "FrostIndexThreshold", &F.FrostIndexThreshold));
    LF_CHECK(scalar_or_map(m, "SnowWaterEquivalent", &F.SnowWaterEquivalent));
    F.DtDay = m->DtDay;
    F.snowmelt_coeff = snowmelt_coeff;
    F.ice_n = ice_melt_coeff_north;
    F.ice_s = ice_melt_coeff_south;
    FIELD(scs, "SnowCoverS");
    FIELD(fi, "FrostIndex");
    FIELD(tp, "TotalPrecipitation");
    FIELD(rain, "Rain");
    FIELD(sm, "SnowMelt");
    FIELD(etr, "ETRef");
    FIELD(ewr, "EWRef");
    FIELD(esr, "ESRef");
    F.SnowCoverS = scs;
    F.FrostIndex = fi;
    F.TotalPrecipitation = tp;
    F.Rain = rain;
    F.SnowMelt = sm;
    F.ETRef = etr;
    F.EWRef = ewr;
    F.ESRef = esr;
    LF_CHECK(flag_buf(m, "isFrozenSoil", &F.frozen));
    if (m->cfg.diagnostics) {
        FIELD(sn, "Snow");
        FIELD(sc, "SnowCover");
        FIELD(pr, "Precipitation");
        FIELD(ta, "Tavg");
        F.Snow = sn;
        F.SnowCover = sc;
        F.Precipitation = pr;
        F.Tavg = ta;
    }
    if (dtype == 2) k_feeder<2><<<lf::blocks_for(m->n, 256), 256, 0, st>>>(F);
    else if (dtype == 1) k_feeder<1><<<lf::blocks_for(m->n, 256), 256, 0, st>>>(F);
    else k_feeder<0><<<lf::blocks_for(m->n, 256), 256, 0, st>>>(F);
    LF_LAUNCH_CHECK();
    if (!on_device) {
        LF_CUDA(cudaEventRecord(m->raw_consumed[m->raw_turn ^ 1], st));
        if (!async) LF_CUDA(cudaStreamSynchronize(st));   // the host buffers are borrowed for the duration of the call
    }
    return LF_OK;
}

int lf_model_feed(lf_model *m, const void *precipitation, const void *tavg, const void *et0, const void *e0, int32_t dtype,
                  double snowmelt_coeff, double ice_melt_coeff_north, double ice_melt_coeff_south, int32_t async)
{
    if (!m || !precipitation || !tavg || !et0 || !e0 || (dtype != 0 && dtype != 1)) {
        lf::set_error("lf_model_feed: null pointer or bad dtype (0 = float64, 1 = float32)");
        return LF_ERR_INVALID;
    }
    return feed_impl(m, precipitation, tavg, et0, e0, dtype, nullptr, nullptr, 0, snowmelt_coeff, ice_melt_coeff_north,
                     ice_melt_coeff_south, async);
}

int lf_model_feed_packed(lf_model *m, const int16_t *precipitation, const int16_t *tavg, const int16_t *et0, const int16_t *e0,
                         const double *scale_factor, const double *add_offset, int32_t decode_float32, double snowmelt_coeff,
                         double ice_melt_coeff_north, double ice_melt_coeff_south, int32_t async)
{
    if (!m || !precipitation || !tavg || !et0 || !e0 || !scale_factor || !add_offset) {
        lf::set_error("lf_model_feed_packed: null pointer");
        return LF_ERR_INVALID;
    }
    return feed_impl(m, precipitation, tavg, et0, e0, 2, scale_factor, add_offset, decode_float32, snowmelt_coeff,
                     ice_melt_coeff_north, ice_melt_coeff_south, async);
}

int lf_model_set_lai(lf_model *m, const double *lai, int64_t count)
{
    if (!m || !lai) {
        lf::set_error("lf_model_set_lai: null pointer");
        return LF_ERR_INVALID;
    }
    LF_CHECK(lf_model_set(m, "LAI", lai, count));
    ScalarOrMap kgb;
    LF_CHECK(scalar_or_map(m, "kgb", &kgb));
    FIELD(l, "LAI");
    FIELD(t, "LAITerm");
    k_lai_term<<<lf::blocks_for(m->n, 256), 256, 0, lf::stream()>>>(l, kgb, t, m->n);
    LF_LAUNCH_CHECK();
    LF_CUDA(cudaStreamSynchronize(lf::stream()));
    return LF_OK;
}

static const char *const RES_NAMES[] = {"TotalReservoirStorageM3CC", "ConservativeStorageLimitCC", "NormalStorageLimitCC",
                                        "Normal_FloodStorageLimitCC", "FloodStorageLimitCC", "MinReservoirOutflowCC",
                                        "NormalReservoirOutflowCC", "NonDamagingReservoirOutflowCC", "DeltaO", "DeltaLN",
                                        "DeltaNFL", "ReservoirStorageM3CC", "ReservoirFillCC", "QResOutM3DtCC"};
static const char *const LAKE_NAMES[] = {"LakeAreaCC", "LakeFactor", "LakeFactorSqr", "LakeStorageM3CC", "LakeOutflowCC",
                                         "LakeInflowOldCC", "LakeStorageM3BalanceCC", "LakeLevelCC", "QLakeOutM3DtCC"};

int lf_model_set_structures(lf_model *m, int32_t n_reservoirs, const int64_t *reservoir_index, int32_t n_lakes,
                            const int64_t *lake_index)
{
    if (!m || n_reservoirs < 0 || n_lakes < 0 || (n_reservoirs > 0 && !reservoir_index) || (n_lakes > 0 && !lake_index)) {
        lf::set_error("lf_model_set_structures: bad arguments");
        return LF_ERR_INVALID;
    }
    // On a cut raster (lf_model_create_from_graphs) the caller passes the structures this rank owns; the pixels that drain
    // into them must be local too (lisflood_code_b200/parallel.py keeps a structure and its feeders on one rank).
    LF_CHECK(lf::ensure_device());
    cudaStream_t st = lf::stream();
    lf_model::Structures &T = m->st;
    std::vector<int32_t> sid(m->n, 0);
    for (int j = 0; j < n_reservoirs; ++j) {
        const int64_t p = reservoir_index[j];
        if (p < 0 || p >= m->n || sid[p] != 0) {
            lf::set_error("lf_model_set_structures: reservoir %d: bad or duplicate pixel index %lld", j, (long long)p);
            return LF_ERR_INVALID;
        }
        sid[p] = j + 1;
    }
    for (int j = 0; j < n_lakes; ++j) {
        const int64_t p = lake_index[j];
        if (p < 0 || p >= m->n || sid[p] != 0) {
            lf::set_error("lf_model_set_structures: lake %d: bad or duplicate pixel index %lld", j, (long long)p);
            return LF_ERR_INVALID;
        }
        sid[p] = -(j + 1);
    }
    lf::DevBuf<int32_t> tmp;
    LF_CHECK(tmp.alloc(m->n));
    LF_CHECK(T.sid.alloc(m->n));
    LF_CHECK(T.feeds.alloc(m->n));
    LF_CHECK(T.CQ0.alloc(m->n));
    LF_CHECK(T.CQ1.alloc(m->n));
    LF_CUDA(cudaMemcpyAsync(tmp.p, sid.data(), m->n * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    k_i32_rows_to_pos<<<lf::blocks_for(m->n, 256), 256, 0, st>>>(tmp.p, T.sid.p, m->g_ch->pix_of_pos.p, m->n);
    LF_LAUNCH_CHECK();
    LF_CUDA(cudaMemsetAsync(T.feeds.p, 0, m->n, st));
    LF_CUDA(cudaMemsetAsync(T.CQ0.p, 0, m->n * sizeof(double), st));
    LF_CUDA(cudaMemsetAsync(T.CQ1.p, 0, m->n * sizeof(double), st));
    k_struct_feeds<<<lf::blocks_for(m->n, 256), 256, 0, st>>>(T.sid.p, m->g_ch->cfirst.p, m->g_ch->cend.p, T.feeds.p, m->n);
    LF_LAUNCH_CHECK();
    T.v.clear();
    for (const char *nm : RES_NAMES) {
        std::unique_ptr<lf::DevBuf<double>> b(new lf::DevBuf<double>());
        LF_CHECK(b->alloc(std::max(n_reservoirs, 1)));
        LF_CUDA(cudaMemsetAsync(b->p, 0, std::max(n_reservoirs, 1) * sizeof(double), st));
        T.v[nm] = std::move(b);
    }
    for (const char *nm : LAKE_NAMES) {
        std::unique_ptr<lf::DevBuf<double>> b(new lf::DevBuf<double>());
        LF_CHECK(b->alloc(std::max(n_lakes, 1)));
        LF_CUDA(cudaMemsetAsync(b->p, 0, std::max(n_lakes, 1) * sizeof(double), st));
        T.v[nm] = std::move(b);
    }
    LF_CUDA(cudaStreamSynchronize(st));
    T.n_res = n_reservoirs;
    T.n_lake = n_lakes;
    T.cq_dirty = true;
    return LF_OK;
}

int lf_model_structure_array(lf_model *m, const char *name, double *values, int64_t count, int32_t set)
{
    if (!m || !name || !values) {
        lf::set_error("lf_model_structure_array: null pointer");
        return LF_ERR_INVALID;
    }
    auto it = m->st.v.find(name);
    if (it == m->st.v.end()) {
        lf::set_error("lf_model_structure_array: unknown per-structure array '%s' (or no structures set)", name);
        return LF_ERR_INVALID;
    }
    bool is_lake = false;
    for (const char *nm : LAKE_NAMES) is_lake = is_lake || strcmp(nm, name) == 0;
    const int64_t want = is_lake ? m->st.n_lake : m->st.n_res;
    if (count != want) {
        lf::set_error("lf_model_structure_array(%s): expected %lld values, got %lld", name, (long long)want, (long long)count);
        return LF_ERR_INVALID;
    }
    LF_CHECK(lf::ensure_device());
    cudaStream_t st = lf::stream();
    if (count > 0) {
        if (set) LF_CUDA(cudaMemcpyAsync(it->second->p, values, count * sizeof(double), cudaMemcpyDefault, st));
        else LF_CUDA(cudaMemcpyAsync(values, it->second->p, count * sizeof(double), cudaMemcpyDefault, st));
    }
    LF_CUDA(cudaStreamSynchronize(st));
    return LF_OK;
}

int lf_model_set_option(lf_model *m, const char *name, double value)
{
    if (!m || !name) {
        lf::set_error("lf_model_set_option: null pointer");
        return LF_ERR_INVALID;
    }
    if (strcmp(name, "overlap_isolated") == 0) m->overlap_isolated = value != 0;
    else if (strcmp(name, "early_blocks_per_sm") == 0) m->early_blocks_per_sm = (int)value;
    else if (strcmp(name, "isolated_blocks_per_sm") == 0) m->iso_blocks_per_sm = (int)value;
    else if (strcmp(name, "flagnancheck") == 0) m->nancheck = value != 0;
    else if (strcmp(name, "cuda_graphs") == 0) m->use_graphs = value != 0;
    else if (strcmp(name, "narrow_runs") == 0) {
        m->narrow_runs = value != 0;
        m->graphs_ch.clear();
    }
    else if (strcmp(name, "accumulate_discharge") == 0) m->accumulate_discharge = value != 0;
    else {
        lf::set_error("lf_model_set_option: unknown option '%s'", name);
        return LF_ERR_INVALID;
    }
    return LF_OK;
}

int lf_model_nonfinite(lf_model *m, int *nonfinite)
{
    if (!m || !nonfinite) {
        lf::set_error("lf_model_nonfinite: null pointer");
        return LF_ERR_INVALID;
    }
    *nonfinite = 0;
    if (!m->nancheck || !m->iso_next.p) return LF_OK;
    LF_CHECK(lf::ensure_device());
    cudaStream_t st = lf::stream();
    LF_CUDA(cudaMemcpyAsync(nonfinite, m->iso_next.p + 2, sizeof(int), cudaMemcpyDeviceToHost, st));
    LF_CUDA(cudaMemsetAsync(m->iso_next.p + 2, 0, sizeof(int), st));
    LF_CUDA(cudaStreamSynchronize(st));
    return LF_OK;
}

void lf_model_destroy(lf_model *m) { delete m; }

}  // extern "C"
