// lf_soil_ops.cu -- the reference's two Numba kernels as stand-alone operators on the device:
//   interception_water_balance   hydrological_modules/soilloop.py:27-70
//   soilColumnsWaterBalance      hydrological_modules/soilloop.py:78-355 (helpers :360-396)
//   suctionUnsaturatedSoilPF     hydrological_modules/soilloop.py:402-432 (option simulatePF; CUDA libm pow / log10)
// One thread per (vegetation fraction, pixel) column, the reference's statement order, the arithmetic library of the
// fused stage (lf_math.cuh; unsat_k / substeps_of from lf_soil_kernel.cuh).  Unlike the fused stage the operators take
// every derived parameter (StoreMaxPervious, PowerInfPot, GenuM, PoreSpaceNotZero, W1, WRes1 ...) from the caller, as
// the reference kernels do.  Host arrays are staged through device temporaries; this entry point is for kernel-level
// drop-in parity, not for speed (the resident path is lf_model_soil).
#include <stddef.h>

#include <vector>

#include "lf_common.cuh"
#include "lf_soil_kernel.cuh"

namespace {

using lfm::dmax;
using lfm::dmin;
using lfm::div_nr;
using lfm::MathTab;

constexpr int OP_THREADS = 128;

__global__ void __launch_bounds__(OP_THREADS) k_op_interception(double *__restrict__ Interception, double *__restrict__ TaInterception,
                                                                double *__restrict__ LeafDrainage, double *__restrict__ CumInterception,
                                                                const double *__restrict__ LAI, const double *__restrict__ Rain,
                                                                const double *__restrict__ TaInterceptionMax, double drainageK,
                                                                int64_t V, int64_t N)
{
    __shared__ MathTab tab;
    lfm::tab_to_shared(&tab, threadIdx.x, OP_THREADS);
    __syncthreads();
    const int64_t k = (int64_t)blockIdx.x * OP_THREADS + threadIdx.x;
    if (k >= V * N) return;
    const int64_t pix = k % N;
    const double lai = LAI[k], rain = Rain[pix];
    double smax;
    if (lai <= .1) smax = 0.;
    else if (lai <= 43.3) smax = 0.935 + 0.498 * lai - 0.00575 * (lai * lai);
    else smax = 11.718;
    double cum = CumInterception[k], interception = 0., ta, leaf;
    if (smax > 0) {
        const double wet = 1. - lfm::exp_neg_tab(div_nr(-0.046 * lai * rain, smax), &tab);
        interception = dmin(dmin(smax - cum, smax * wet), rain);
        cum += interception;
    }
    if (cum > 0.) {
        ta = dmax(dmin(cum, TaInterceptionMax[k]), 0.);
        cum = dmax(cum - ta, 0.);
        leaf = drainageK * cum;
        cum = dmax(cum - leaf, 0.);
    } else {
        ta = 0.;
        leaf = 0.;
    }
    Interception[k] = interception;
    TaInterception[k] = ta;
    LeafDrainage[k] = leaf;
    CumInterception[k] = cum;
}

__global__ void __launch_bounds__(OP_THREADS) k_op_soil_columns(const __grid_constant__ lf_soil_columns_args A)
{
    __shared__ MathTab tab;
    lfm::tab_to_shared(&tab, threadIdx.x, OP_THREADS);
    __syncthreads();
    const MathTab *MT = &tab;
    const int64_t N = A.num_pixs;
    const int64_t v = (int64_t)blockIdx.x * OP_THREADS + threadIdx.x;  // (veg, pix) flat index
    if (v >= A.num_vegs * N) return;
    const int64_t veg = v / N, pix = v - veg * N;
    const int64_t l = A.index_landuse_all[veg] * N + pix;
    const bool drained = A.is_irrigated[veg] && (A.DrainedFraction > 0);  // :115
    const bool frozen = A.isFrozenSoil[pix] != 0;
    // available water, days since last rain (:131-140)
    double avail = dmax((A.Rain[pix] + A.SnowMelt[pix]) + A.LeafDrainage[v] - A.Interception[v], 0.);
    double dslr = A.DSLR[v];
    if (avail > A.AvWaterThreshold) dslr = 1;
    else dslr += A.DtDay;
    A.DSLR[v] = dslr;
    // actual soil evaporation (:148-162)
    double w1a = A.W1a[v], w1b = A.W1b[v], w2 = A.W2[v], w1 = A.W1[v];
    const double wres1a = A.WRes1a[l], wres1b = A.WRes1b[l], wres2 = A.WRes2[l];
    const double ws1a = A.WS1a[l], ws1b = A.WS1b[l], ws2 = A.WS2[l];
    double esact;
    if (frozen) {
        esact = 0.;
    } else {
        esact = A.ESMax[v] * (lfm::sqrt_nr(dslr) - lfm::sqrt_nr(dslr - 1));
        esact = dmax(dmin(esact, w1 - A.WRes1[l]), 0.);
        const double supply1a = w1a - wres1a;
        const double es1a = dmin(esact, supply1a), es1b = dmax(esact - supply1a, 0.);
        w1a = dmax(w1a - es1a, wres1a);
        w1b = dmax(w1b - es1b, wres1b);
    }
    A.ESAct[v] = esact;
    w1 = w1a + w1b;
    // infiltration capacity, preferential flow, infiltration (:168-211)
    const bool pore1a = A.PoreSpaceNotZero1a[l] != 0, pore1b = A.PoreSpaceNotZero1b[l] != 0, pore2 = A.PoreSpaceNotZero2[l] != 0;
    const double relsat1 = pore1a ? dmin(div_nr(w1, A.WS1[l]), 1.0) : 0.0;
    const double satfrac = 1.0 - lfm::pw_tab<true>(1.0 - relsat1, A.b_Xinanjiang[pix], MT);
    const double infpot = frozen ? 0.0 : A.StoreMaxPervious[l] * lfm::pw_tab<true>(1. - satfrac, A.PowerInfPot[pix], MT) * A.DtDay;
    const double prefflow = lfm::pw_tab<true>(relsat1, A.PowerPrefFlow[pix], MT) * avail;
    avail -= prefflow;
    A.PrefFlow[v] = prefflow;
    A.AvailableWaterForInfiltration[v] = avail;
    double infil = dmax(dmin(avail, infpot), 0.);
    {
        const double test = w1a + infil;
        w1a = dmin(ws1a, test);
        w1b += dmax(test - ws1a, 0.);
    }
    // conductivities, Courant number, Darcy sub-steps (:220-312)
    const double ks1a = A.KSat1a[l], ks1b = A.KSat1b[l], ks2 = A.KSat2[l];
    const double im1a = A.GenuInvM1a[l], im1b = A.GenuInvM1b[l], im2 = A.GenuInvM2[l];
    const double m1a = A.GenuM1a[l], m1b = A.GenuM1b[l], m2 = A.GenuM2[l];
    double k1a = lfsoil::unsat_k(w1a, pore1a, wres1a, ws1a, ks1a, im1a, m1a, MT);
    double k1b = lfsoil::unsat_k(w1b, pore1b, wres1b, ws1b, ks1b, im1b, m1b, MT);
    double k2 = lfsoil::unsat_k(w2, pore2, wres2, ws2, ks2, im2, m2, MT);
    double av1a = w1a - wres1a, av1b = w1b - wres1b, av2 = w2 - wres2;
    double cap1 = ws1b - w1b, cap2 = ws2 - w2;
    lfsoil::Ptrs P;  // substeps_of reads the two scalars only
    P.DtDay = A.DtDay;
    P.CourantCrit = A.CourantCrit;
    const int nsub = lfsoil::substeps_of(P, k1a, k1b, k2, av1a, av1b, av2);
    if (A.NoSubS_out) A.NoSubS_out[v] = nsub;
    const double dtsub = div_nr(A.DtDay, (double)nsub);
    double seepA = 0., seepB = 0., seepG = 0.;
    double wt1a = w1a, wt1b = w1b, wt2 = w2;
    for (int s = 0; s < nsub; ++s) {
        if (s > 0) {
            k1a = lfsoil::unsat_k(wt1a, pore1a, wres1a, ws1a, ks1a, im1a, m1a, MT);
            k1b = lfsoil::unsat_k(wt1b, pore1b, wres1b, ws1b, ks1b, im1b, m1b, MT);
            k2 = lfsoil::unsat_k(wt2, pore2, wres2, ws2, ks2, im2, m2, MT);
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
    if (frozen) seepA = seepB = seepG = 0.;  // :313-316
    A.SeepTopToSubA[v] = seepA;
    A.SeepTopToSubB[v] = seepB;
    A.SeepSubToGW[v] = seepG;
    // storages (:319-325)
    w1a -= seepA;
    w1b = w1b + seepA - seepB;
    w2 = w2 + seepB - seepG;
    w1 = w1a + w1b;
    infil -= dmax(w1a - ws1a, 0.);
    w1a = dmin(w1a, ws1a);
    A.Infiltration[v] = infil;
    A.W1a[v] = w1a;
    A.W1b[v] = w1b;
    A.W2[v] = w2;
    A.W1[v] = w1;
    // theta / saturation (:330-336)
    A.Theta1a[v] = pore1a ? w1a / A.SoilDepth1a[l] : 0.;
    A.Theta1b[v] = pore1b ? w1b / A.SoilDepth1b[l] : 0.;
    A.Theta2[v] = pore2 ? w2 / A.SoilDepth2[l] : 0.;
    A.Sat1a[v] = (w1a - A.WWP1a[l]) / (A.WFC1a[l] - A.WWP1a[l]);
    A.Sat1b[v] = (w1b - A.WWP1b[l]) / (A.WFC1b[l] - A.WWP1b[l]);
    A.Sat1[v] = (w1 - A.WWP1[l]) / (A.WFC1[l] - A.WWP1[l]);
    A.Sat2[v] = (w2 - A.WWP2[l]) / (A.WFC2[l] - A.WWP2[l]);
    // upper zone (:340-354)
    double uz = A.UZ[v];
    double uzout = dmin(A.UpperZoneK[pix] * uz, uz);
    uz = dmax(uz - uzout, 0.);
    if (drained) {
        uzout += A.DrainedFraction * seepG;
        uz += (1 - A.DrainedFraction) * seepG + prefflow;
    } else {
        uz += seepG + prefflow;
    }
    const double gwp = dmin(A.GwPercStep[pix], uz);
    uz = dmax(uz - gwp, 0.);
    A.UZOutflow[v] = uzout;
    A.GwPercUZLZ[v] = gwp;
    A.UZ[v] = uz;
}

// host <-> device staging of one argument
// suctionUnsaturatedSoilPF, soilloop.py:402-424: pF = log10 of the capillary head of the three soil layers, from the
// saturation term (saturationDegree, :379-383) through the inverse Van Genuchten curve (pressureHead, :428-432).
// An optional output computed once per reported step: the standard pow / log10 are used, not the hot path's tables.
__global__ void __launch_bounds__(OP_THREADS) k_op_suction_pf(const __grid_constant__ lf_soil_pf_args A, const int64_t *__restrict__ landuse_of_veg)
{
    const int64_t k = (int64_t)blockIdx.x * OP_THREADS + threadIdx.x;
    const int64_t V = A.num_vegs, N = A.num_pixs;
    if (k >= V * N) return;
    const int64_t pix = k % N, lu = landuse_of_veg[k / N] * N + pix;
#pragma unroll
    for (int layer = 0; layer < 3; ++layer) {
        double sat = 0.;
        if (A.PoreSpaceNotZero[layer][lu]) {
            const double wres = A.WRes[layer][lu];
            sat = dmax(dmin((A.W[layer][k] - wres) / (A.WS[layer][lu] - wres), 1.), 0.);
        }
        double head = A.HeadMax;
        if (sat != 0) head = fmin(A.HeadMax, A.GenuInvAlpha[layer][lu] * pow(pow(1. / sat, A.GenuInvM[layer][lu]) - 1., A.GenuInvN[layer][lu]));
        A.pF[layer][k] = head > 0 ? log10(head) : -1.;
    }
}

struct Staged {
    void *dev = nullptr;
    void *host = nullptr;
    size_t bytes = 0;
    bool owned = false, copy_back = false;
};

struct Stager {
    std::vector<Staged> items;
    cudaStream_t st;
    explicit Stager(cudaStream_t s) : st(s) {}
    ~Stager()
    {
        for (Staged &s : items)
            if (s.owned && s.dev) cudaFree(s.dev);
    }
    // returns the device address to use for `p` (itself when it already is device memory)
    int in(const void *p, size_t bytes, bool out, void **dev)
    {
        if (!p) {
            *dev = nullptr;
            return LF_OK;
        }
        if (lf::is_device_ptr(p)) {
            *dev = const_cast<void *>(p);
            return LF_OK;
        }
        Staged s;
        s.host = const_cast<void *>(p);
        s.bytes = bytes;
        s.owned = true;
        s.copy_back = out;
        cudaError_t e = cudaMalloc(&s.dev, bytes ? bytes : 1);
        if (e != cudaSuccess) {
            lf::set_error("cudaMalloc(%zu bytes) -> %s", bytes, cudaGetErrorString(e));
            return LF_ERR_CUDA;
        }
        items.push_back(s);
        LF_CUDA(cudaMemcpyAsync(s.dev, p, bytes, cudaMemcpyHostToDevice, st));
        *dev = s.dev;
        return LF_OK;
    }
    int finish()
    {
        for (Staged &s : items)
            if (s.copy_back) LF_CUDA(cudaMemcpyAsync(s.host, s.dev, s.bytes, cudaMemcpyDeviceToHost, st));
        LF_CUDA(cudaStreamSynchronize(st));
        return LF_OK;
    }
};

}  // namespace

extern "C" int lf_interception_water_balance(double *Interception, double *TaInterception, double *LeafDrainage,
                                             double *CumInterception, const double *LAI, const double *Rain,
                                             const double *TaInterceptionMax, double drainageK, int64_t num_vegs,
                                             int64_t num_pixs)
{
    if (!Interception || !TaInterception || !LeafDrainage || !CumInterception || !LAI || !Rain || !TaInterceptionMax ||
        num_vegs <= 0 || num_pixs <= 0) {
        lf::set_error("lf_interception_water_balance: null pointer or empty shape");
        return LF_ERR_INVALID;
    }
    LF_CHECK(lf::ensure_device());
    cudaStream_t st = lf::stream();
    Stager S(st);
    const size_t vn = (size_t)num_vegs * num_pixs * sizeof(double), n = (size_t)num_pixs * sizeof(double);
    void *d[7];
    LF_CHECK(S.in(Interception, vn, true, &d[0]));
    LF_CHECK(S.in(TaInterception, vn, true, &d[1]));
    LF_CHECK(S.in(LeafDrainage, vn, true, &d[2]));
    LF_CHECK(S.in(CumInterception, vn, true, &d[3]));
    LF_CHECK(S.in(LAI, vn, false, &d[4]));
    LF_CHECK(S.in(Rain, n, false, &d[5]));
    LF_CHECK(S.in(TaInterceptionMax, vn, false, &d[6]));
    k_op_interception<<<lf::blocks_for(num_vegs * num_pixs, OP_THREADS), OP_THREADS, 0, st>>>(
        (double *)d[0], (double *)d[1], (double *)d[2], (double *)d[3], (const double *)d[4], (const double *)d[5],
        (const double *)d[6], drainageK, num_vegs, num_pixs);
    LF_LAUNCH_CHECK();
    return S.finish();
}

extern "C" int lf_soil_columns_water_balance(const lf_soil_columns_args *args)
{
    if (!args || args->num_vegs <= 0 || args->num_pixs <= 0 || args->num_landuses <= 0 || !args->index_landuse_all ||
        !args->is_irrigated) {
        lf::set_error("lf_soil_columns_water_balance: null argument block or empty shape");
        return LF_ERR_INVALID;
    }
    LF_CHECK(lf::ensure_device());
    lf_soil_columns_args A = *args;
    const int64_t V = A.num_vegs, N = A.num_pixs, L = A.num_landuses;
    if (lf::is_device_ptr(A.index_landuse_all) || lf::is_device_ptr(A.is_irrigated) ||
        (A.is_paddy_irrig && lf::is_device_ptr(A.is_paddy_irrig))) {
        lf::set_error("lf_soil_columns_water_balance: the per-vegetation index vectors are host arrays");
        return LF_ERR_INVALID;
    }
    for (int64_t v = 0; v < V; ++v) {
        if (A.index_landuse_all[v] < 0 || A.index_landuse_all[v] >= L) {
            lf::set_error("lf_soil_columns_water_balance: index_landuse_all[%lld] outside 0..%lld", (long long)v, (long long)L - 1);
            return LF_ERR_INVALID;
        }
        if (A.is_paddy_irrig && A.is_paddy_irrig[v]) {
            lf::set_error("lf_soil_columns_water_balance: paddy-rice fractions belong to the EPIC crop module (out of scope)");
            return LF_ERR_INVALID;
        }
    }
    cudaStream_t st = lf::stream();
    Stager S(st);
    const size_t vn = (size_t)V * N, ln = (size_t)L * N, n = (size_t)N;
    void *dv = nullptr;
#define OPARG(member, count, type, out)                                                  \
    do {                                                                                 \
        if (!A.member) {                                                                 \
            lf::set_error("lf_soil_columns_water_balance: " #member " is NULL");          \
            return LF_ERR_INVALID;                                                       \
        }                                                                                \
        LF_CHECK(S.in(A.member, (count) * sizeof(type), out, &dv));                      \
        A.member = (decltype(A.member))dv;                                               \
    } while (0)
    OPARG(index_landuse_all, (size_t)V, int64_t, false);
    OPARG(is_irrigated, (size_t)V, uint8_t, false);
    A.is_paddy_irrig = nullptr;
    OPARG(AvailableWaterForInfiltration, vn, double, true);
    OPARG(Rain, n, double, false);
    OPARG(SnowMelt, n, double, false);
    OPARG(LeafDrainage, vn, double, false);
    OPARG(Interception, vn, double, false);
    OPARG(DSLR, vn, double, true);
    OPARG(ESAct, vn, double, true);
    OPARG(ESMax, vn, double, false);
    OPARG(isFrozenSoil, n, uint8_t, false);
    OPARG(b_Xinanjiang, n, double, false);
    OPARG(StoreMaxPervious, ln, double, false);
    OPARG(PowerInfPot, n, double, false);
    OPARG(PrefFlow, vn, double, true);
    OPARG(PowerPrefFlow, n, double, false);
    OPARG(Infiltration, vn, double, true);
    OPARG(PoreSpaceNotZero1a, ln, uint8_t, false);
    OPARG(PoreSpaceNotZero1b, ln, uint8_t, false);
    OPARG(PoreSpaceNotZero2, ln, uint8_t, false);
    OPARG(KSat1a, ln, double, false);
    OPARG(KSat1b, ln, double, false);
    OPARG(KSat2, ln, double, false);
    OPARG(GenuInvM1a, ln, double, false);
    OPARG(GenuInvM1b, ln, double, false);
    OPARG(GenuInvM2, ln, double, false);
    OPARG(GenuM1a, ln, double, false);
    OPARG(GenuM1b, ln, double, false);
    OPARG(GenuM2, ln, double, false);
    OPARG(W1a, vn, double, true);
    OPARG(W1b, vn, double, true);
    OPARG(W1, vn, double, true);
    OPARG(W2, vn, double, true);
    OPARG(Theta1a, vn, double, true);
    OPARG(Theta1b, vn, double, true);
    OPARG(Theta2, vn, double, true);
    OPARG(Sat1a, vn, double, true);
    OPARG(Sat1b, vn, double, true);
    OPARG(Sat1, vn, double, true);
    OPARG(Sat2, vn, double, true);
    OPARG(SeepTopToSubA, vn, double, true);
    OPARG(SeepTopToSubB, vn, double, true);
    OPARG(SeepSubToGW, vn, double, true);
    OPARG(WRes1a, ln, double, false);
    OPARG(WRes1b, ln, double, false);
    OPARG(WRes1, ln, double, false);
    OPARG(WRes2, ln, double, false);
    OPARG(WWP1a, ln, double, false);
    OPARG(WWP1b, ln, double, false);
    OPARG(WWP1, ln, double, false);
    OPARG(WWP2, ln, double, false);
    OPARG(WFC1a, ln, double, false);
    OPARG(WFC1b, ln, double, false);
    OPARG(WFC1, ln, double, false);
    OPARG(WFC2, ln, double, false);
    OPARG(SoilDepth1a, ln, double, false);
    OPARG(SoilDepth1b, ln, double, false);
    OPARG(SoilDepth2, ln, double, false);
    OPARG(WS1a, ln, double, false);
    OPARG(WS1b, ln, double, false);
    OPARG(WS1, ln, double, false);
    OPARG(WS2, ln, double, false);
    OPARG(UpperZoneK, n, double, false);
    OPARG(GwPercStep, n, double, false);
    OPARG(UZOutflow, vn, double, true);
    OPARG(UZ, vn, double, true);
    OPARG(GwPercUZLZ, vn, double, true);
    if (A.NoSubS_out) OPARG(NoSubS_out, vn, int64_t, true);
#undef OPARG
    k_op_soil_columns<<<lf::blocks_for(V * N, OP_THREADS), OP_THREADS, 0, st>>>(A);
    LF_LAUNCH_CHECK();
    return S.finish();
}

extern "C" int lf_suction_unsaturated_soil_pf(const lf_soil_pf_args *args)
{
    if (!args || args->num_vegs <= 0 || args->num_pixs <= 0 || args->num_landuses <= 0 || !args->index_landuse_all) {
        lf::set_error("lf_suction_unsaturated_soil_pf: null argument block or empty shape");
        return LF_ERR_INVALID;
    }
    LF_CHECK(lf::ensure_device());
    lf_soil_pf_args A = *args;
    const int64_t V = A.num_vegs, N = A.num_pixs, L = A.num_landuses;
    if (lf::is_device_ptr(A.index_landuse_all)) {
        lf::set_error("lf_suction_unsaturated_soil_pf: index_landuse_all is a host array");
        return LF_ERR_INVALID;
    }
    for (int64_t v = 0; v < V; ++v)
        if (A.index_landuse_all[v] < 0 || A.index_landuse_all[v] >= L) {
            lf::set_error("lf_suction_unsaturated_soil_pf: index_landuse_all[%lld] outside 0..%lld", (long long)v, (long long)L - 1);
            return LF_ERR_INVALID;
        }
    cudaStream_t st = lf::stream();
    Stager S(st);
    const size_t vn = (size_t)V * N, ln = (size_t)L * N;
    void *dv = nullptr, *d_index = nullptr;
    LF_CHECK(S.in(A.index_landuse_all, (size_t)V * sizeof(int64_t), false, &d_index));
    for (int layer = 0; layer < 3; ++layer) {
        if (!A.pF[layer] || !A.W[layer] || !A.WRes[layer] || !A.WS[layer] || !A.PoreSpaceNotZero[layer] || !A.GenuInvAlpha[layer] ||
            !A.GenuInvM[layer] || !A.GenuInvN[layer]) {
            lf::set_error("lf_suction_unsaturated_soil_pf: a map of layer %d is NULL", layer);
            return LF_ERR_INVALID;
        }
        LF_CHECK(S.in(A.pF[layer], vn * sizeof(double), true, &dv));
        A.pF[layer] = (double *)dv;
        LF_CHECK(S.in(A.W[layer], vn * sizeof(double), false, &dv));
        A.W[layer] = (const double *)dv;
        LF_CHECK(S.in(A.WRes[layer], ln * sizeof(double), false, &dv));
        A.WRes[layer] = (const double *)dv;
        LF_CHECK(S.in(A.WS[layer], ln * sizeof(double), false, &dv));
        A.WS[layer] = (const double *)dv;
        LF_CHECK(S.in(A.PoreSpaceNotZero[layer], ln * sizeof(uint8_t), false, &dv));
        A.PoreSpaceNotZero[layer] = (const uint8_t *)dv;
        LF_CHECK(S.in(A.GenuInvAlpha[layer], ln * sizeof(double), false, &dv));
        A.GenuInvAlpha[layer] = (const double *)dv;
        LF_CHECK(S.in(A.GenuInvM[layer], ln * sizeof(double), false, &dv));
        A.GenuInvM[layer] = (const double *)dv;
        LF_CHECK(S.in(A.GenuInvN[layer], ln * sizeof(double), false, &dv));
        A.GenuInvN[layer] = (const double *)dv;
    }
    k_op_suction_pf<<<lf::blocks_for(V * N, OP_THREADS), OP_THREADS, 0, st>>>(A, (const int64_t *)d_index);
    LF_LAUNCH_CHECK();
    return S.finish();
}
