/*
 * lisf_oracle_soil.c -- CPU restatement of the reference's soil Numba kernels.  TEST INFRASTRUCTURE ONLY
 * (see the header of lisf_oracle.c).  Reference: src/lisflood/hydrological_modules/soilloop.py.
 *
 *   lfo_interception   <- interception_water_balance       soilloop.py:27-70
 *   lfo_soil_columns   <- soilColumnsWaterBalance           soilloop.py:78-355
 *                         unsaturatedConductivity :360-367, saturationDegree :378-383,
 *                         thetaFun :386-389, satFun :393-396
 * Arrays are C-contiguous float64: (V, N) per vegetation fraction, (L, N) per land use, (N) per pixel;
 * booleans are uint8.  The paddy branch (:107-113) is inactive without EPIC and is not restated.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

static inline double dmin(double a, double b) { return a < b ? a : b; }
static inline double dmax(double a, double b) { return a > b ? a : b; }

/* interception_water_balance, soilloop.py:27-70 */
void lfo_interception(double *Interception, double *TaInterception, double *LeafDrainage, double *CumInterception,
                      const double *LAI, const double *Rain, const double *TaInterceptionMax, double drainageK,
                      int64_t num_vegs, int64_t num_pixs)
{
    for (int64_t veg = 0; veg < num_vegs; ++veg) {
#pragma omp parallel for schedule(static)
        for (int64_t pix = 0; pix < num_pixs; ++pix) {
            int64_t i = veg * num_pixs + pix;
            double lai = LAI[i], smax;
            if (lai <= .1) smax = 0.;
            else if (lai <= 43.3) smax = 0.935 + 0.498 * lai - 0.00575 * (lai * lai);
            else smax = 11.718;
            if (smax > 0) {
                double a = smax - CumInterception[i];
                double b = smax * (1. - exp(-0.046 * lai * Rain[pix] / smax));
                Interception[i] = dmin(dmin(a, b), Rain[pix]);
                CumInterception[i] += Interception[i];
            } else {
                Interception[i] = 0.;
            }
            if (CumInterception[i] > 0.) {
                TaInterception[i] = dmax(dmin(CumInterception[i], TaInterceptionMax[i]), 0.);
                CumInterception[i] = dmax(CumInterception[i] - TaInterception[i], 0.);
                LeafDrainage[i] = drainageK * CumInterception[i];
                CumInterception[i] = dmax(CumInterception[i] - LeafDrainage[i], 0.);
            } else {
                TaInterception[i] = 0.;
                LeafDrainage[i] = 0.;
            }
        }
    }
}

/* saturationDegree :378-383 + unsaturatedConductivity :360-367 */
static inline double unsat_conductivity(double w, int pore, double wres, double ws, double ksat, double invm, double m)
{
    double sat = pore ? dmax(dmin((w - wres) / (ws - wres), 1.), 0.) : 0.;
    double t = 1. - pow(1. - pow(sat, invm), m);
    return ksat * sqrt(sat) * (t * t);
}

typedef struct {
    int64_t num_vegs, num_pixs;
    const int64_t *index_landuse_all;
    const uint8_t *is_irrigated;
    double DtDay, AvWaterThreshold, CourantCrit, DrainedFraction;
    double *AvailableWaterForInfiltration;
    const double *Rain, *SnowMelt, *LeafDrainage, *Interception;
    double *DSLR, *ESAct;
    const double *ESMax;
    const uint8_t *isFrozenSoil;
    const double *b_Xinanjiang, *StoreMaxPervious, *PowerInfPot;
    double *PrefFlow;
    const double *PowerPrefFlow;
    double *Infiltration;
    const uint8_t *PoreSpaceNotZero1a, *PoreSpaceNotZero1b, *PoreSpaceNotZero2;
    const double *KSat1a, *KSat1b, *KSat2, *GenuInvM1a, *GenuInvM1b, *GenuInvM2, *GenuM1a, *GenuM1b, *GenuM2;
    double *W1a, *W1b, *W1, *W2, *Theta1a, *Theta1b, *Theta2, *Sat1a, *Sat1b, *Sat1, *Sat2;
    double *SeepTopToSubA, *SeepTopToSubB, *SeepSubToGW;
    const double *WRes1a, *WRes1b, *WRes1, *WRes2, *WWP1a, *WWP1b, *WWP1, *WWP2, *WFC1a, *WFC1b, *WFC1, *WFC2;
    const double *SoilDepth1a, *SoilDepth1b, *SoilDepth2, *WS1a, *WS1b, *WS1, *WS2;
    const double *UpperZoneK, *GwPercStep;
    double *UZOutflow, *UZ, *GwPercUZLZ;
    int64_t *NoSubS_out; /* optional (V,N) diagnostic: sub-steps taken */
} lfo_soil_args;

/* soilColumnsWaterBalance, soilloop.py:78-355 */
void lfo_soil_columns(const lfo_soil_args *A)
{
    const int64_t N = A->num_pixs;
    for (int64_t veg = 0; veg < A->num_vegs; ++veg) { /* serial over vegetation, :105 */
        const int64_t lu = A->index_landuse_all[veg];
        const int drained = A->is_irrigated[veg] && (A->DrainedFraction > 0); /* :115 */
#pragma omp parallel for schedule(static)
        for (int64_t pix = 0; pix < N; ++pix) {
            const int64_t v = veg * N + pix, l = lu * N + pix;
            /* available water, :131 */
            A->AvailableWaterForInfiltration[v] =
                dmax((A->Rain[pix] + A->SnowMelt[pix]) + A->LeafDrainage[v] - A->Interception[v], 0.);
            /* days since last rain, :137-140 */
            if (A->AvailableWaterForInfiltration[v] > A->AvWaterThreshold) A->DSLR[v] = 1;
            else A->DSLR[v] += A->DtDay;
            /* actual soil evaporation, :148-162 */
            if (A->isFrozenSoil[pix]) {
                A->ESAct[v] = 0.;
            } else {
                A->ESAct[v] = A->ESMax[v] * (sqrt(A->DSLR[v]) - sqrt(A->DSLR[v] - 1));
                A->ESAct[v] = dmax(dmin(A->ESAct[v], A->W1[v] - A->WRes1[l]), 0.);
                double supply1a = A->W1a[v] - A->WRes1a[l];
                double es1a = dmin(A->ESAct[v], supply1a);
                double es1b = dmax(A->ESAct[v] - supply1a, 0.);
                A->W1a[v] = dmax(A->W1a[v] - es1a, A->WRes1a[l]);
                A->W1b[v] = dmax(A->W1b[v] - es1b, A->WRes1b[l]);
            }
            A->W1[v] = A->W1a[v] + A->W1b[v];
            /* infiltration capacity, :168-179 */
            double RelSat1 = A->PoreSpaceNotZero1a[l] ? dmin(A->W1[v] / A->WS1[l], 1.0) : 0.0;
            double SatFraction = 1.0 - pow(1.0 - RelSat1, A->b_Xinanjiang[pix]);
            double InfPot = A->isFrozenSoil[pix]
                                ? 0.0
                                : A->StoreMaxPervious[l] * pow(1. - SatFraction, A->PowerInfPot[pix]) * A->DtDay;
            /* preferential flow, :190-194 */
            A->PrefFlow[v] = pow(RelSat1, A->PowerPrefFlow[pix]) * A->AvailableWaterForInfiltration[v];
            A->AvailableWaterForInfiltration[v] -= A->PrefFlow[v];
            /* infiltration, :201-211 */
            A->Infiltration[v] = dmax(dmin(A->AvailableWaterForInfiltration[v], InfPot), 0.);
            double testW1a = A->W1a[v] + A->Infiltration[v];
            A->W1a[v] = dmin(A->WS1a[l], testW1a);
            A->W1b[v] += dmax(testW1a - A->WS1a[l], 0.);
            /* fluxes between layers: conductivities and Courant numbers, :220-249 */
            double K1a = unsat_conductivity(A->W1a[v], A->PoreSpaceNotZero1a[l], A->WRes1a[l], A->WS1a[l], A->KSat1a[l],
                                            A->GenuInvM1a[l], A->GenuM1a[l]);
            double K1b = unsat_conductivity(A->W1b[v], A->PoreSpaceNotZero1b[l], A->WRes1b[l], A->WS1b[l], A->KSat1b[l],
                                            A->GenuInvM1b[l], A->GenuM1b[l]);
            double K2 = unsat_conductivity(A->W2[v], A->PoreSpaceNotZero2[l], A->WRes2[l], A->WS2[l], A->KSat2[l],
                                           A->GenuInvM2[l], A->GenuM2[l]);
            double Av1a = A->W1a[v] - A->WRes1a[l], Av1b = A->W1b[v] - A->WRes1b[l], Av2 = A->W2[v] - A->WRes2[l];
            double Cap1 = A->WS1b[l] - A->W1b[v], Cap2 = A->WS2[l] - A->W2[v];
            double cA = Av1a == 0 ? 0. : K1a * A->DtDay / Av1a;
            double cB = Av1b == 0 ? 0. : K1b * A->DtDay / Av1b;
            double cG = Av2 == 0 ? 0. : K2 * A->DtDay / Av2;
            double courant = dmax(dmax(cA, cB), cG);
            double nsub_f = dmax(1., ceil(courant / A->CourantCrit));
            int64_t NoSubS = (int64_t)nsub_f;
            if (A->NoSubS_out) A->NoSubS_out[v] = NoSubS;
            /* sub-step loop, :264-312 */
            double WT1a = A->W1a[v], WT1b = A->W1b[v], WT2 = A->W2[v];
            double seepA = 0., seepB = 0., seepG = 0.;
            double DtSub = A->DtDay / (double)NoSubS;
            for (int64_t i = 0; i < NoSubS; ++i) {
                if (i > 0) {
                    K1a = unsat_conductivity(WT1a, A->PoreSpaceNotZero1a[l], A->WRes1a[l], A->WS1a[l], A->KSat1a[l],
                                             A->GenuInvM1a[l], A->GenuM1a[l]);
                    K1b = unsat_conductivity(WT1b, A->PoreSpaceNotZero1b[l], A->WRes1b[l], A->WS1b[l], A->KSat1b[l],
                                             A->GenuInvM1b[l], A->GenuM1b[l]);
                    K2 = unsat_conductivity(WT2, A->PoreSpaceNotZero2[l], A->WRes2[l], A->WS2[l], A->KSat2[l],
                                            A->GenuInvM2[l], A->GenuM2[l]);
                }
                double sA = dmin(K1a * DtSub, Cap1);
                double sB = dmin(K1b * DtSub, Cap2);
                double sG = dmin(K2 * DtSub, Av2);
                Av1a -= sA;
                Av1b += sA - sB;
                Av2 += sB - sG;
                WT1a = Av1a + A->WRes1a[l];
                WT1b = Av1b + A->WRes1b[l];
                WT2 = Av2 + A->WRes2[l];
                Cap1 = A->WS1b[l] - WT1b;
                Cap2 = A->WS2[l] - WT2;
                seepA += sA;
                seepB += sB;
                seepG += sG;
            }
            if (A->isFrozenSoil[pix]) seepA = seepB = seepG = 0.; /* :313-316 */
            A->SeepTopToSubA[v] = seepA;
            A->SeepTopToSubB[v] = seepB;
            A->SeepSubToGW[v] = seepG;
            /* update storages, :319-325 */
            A->W1a[v] -= seepA;
            A->W1b[v] = A->W1b[v] + seepA - seepB;
            A->W2[v] = A->W2[v] + seepB - seepG;
            A->W1[v] = A->W1a[v] + A->W1b[v];
            A->Infiltration[v] -= dmax(A->W1a[v] - A->WS1a[l], 0.);
            A->W1a[v] = dmin(A->W1a[v], A->WS1a[l]);
            /* theta / saturation diagnostics, :330-336 */
            A->Theta1a[v] = A->PoreSpaceNotZero1a[l] ? A->W1a[v] / A->SoilDepth1a[l] : 0.;
            A->Theta1b[v] = A->PoreSpaceNotZero1b[l] ? A->W1b[v] / A->SoilDepth1b[l] : 0.;
            A->Theta2[v] = A->PoreSpaceNotZero2[l] ? A->W2[v] / A->SoilDepth2[l] : 0.;
            A->Sat1a[v] = (A->W1a[v] - A->WWP1a[l]) / (A->WFC1a[l] - A->WWP1a[l]);
            A->Sat1b[v] = (A->W1b[v] - A->WWP1b[l]) / (A->WFC1b[l] - A->WWP1b[l]);
            A->Sat1[v] = (A->W1[v] - A->WWP1[l]) / (A->WFC1[l] - A->WWP1[l]);
            A->Sat2[v] = (A->W2[v] - A->WWP2[l]) / (A->WFC2[l] - A->WWP2[l]);
            /* upper zone, :340-354 */
            A->UZOutflow[v] = dmin(A->UpperZoneK[pix] * A->UZ[v], A->UZ[v]);
            A->UZ[v] = dmax(A->UZ[v] - A->UZOutflow[v], 0.);
            if (drained) {
                A->UZOutflow[v] += A->DrainedFraction * A->SeepSubToGW[v];
                A->UZ[v] += (1 - A->DrainedFraction) * A->SeepSubToGW[v] + A->PrefFlow[v];
            } else {
                A->UZ[v] += A->SeepSubToGW[v] + A->PrefFlow[v];
            }
            A->GwPercUZLZ[v] = dmin(A->GwPercStep[pix], A->UZ[v]);
            A->UZ[v] = dmax(A->UZ[v] - A->GwPercUZLZ[v], 0.);
        }
    }
}
