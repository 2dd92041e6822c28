/*
 * lisf_oracle.c -- CPU restatement of the LISFLOOD raster hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is the parity ORACLE for lisflood_code_b200: a plain-C (C99 + OpenMP), float64 restatement
 * of the reference's algorithm, each function citing the reference file:line it follows.  It is NOT
 * part of the product: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` legs may load it.  The product path (lisflood_code_b200/csrc) never links,
 * imports or calls it and fails loudly when its CUDA library is missing.
 *
 * Pinning: the reference ships no kernel-level golden vectors.  This restatement is
 * pinned against the UNMODIFIED reference kernels imported live from /root/reference in the build
 * container (oracle/ref_loader.py) -- see tests/golden/make_golden.py, which commits the reference's
 * outputs as fixtures under tests/golden/, and tests/test_oracle_vs_golden.py which checks this
 * file against them on every run.
 *
 * Reference paths below are relative to /root/reference/src/lisflood/hydrological_modules/.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define NEWTON_TOL 1e-12 /* kinematic_wave_parallel_tools.py:26 */
#define MAX_ITERS 3000   /* kinematic_wave_parallel_tools.py:27 */

int lfo_set_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
    return omp_get_max_threads();
#else
    (void)n;
    return 1;
#endif
}

/* ------------------------------------------------------------------------------------------------
 * D8 graph: kinematic_wave_parallel.py:47-71 (decode), :73-90 + tools:111-130 (lookups),
 * :92-106 (topoDistFromSea), :140-158 (_setRoutingOrders).
 * ---------------------------------------------------------------------------------------------- */

/* keypad code -> direction index 0..7 (8 = pit); FLOW_CODE = [2,3,6,9,8,7,4,1,5], SEA_CODE 0 -> pit
 * (kinematic_wave_parallel.py:47-51, :64-71).  Returns -1 for a code outside 0..9. */
static int decode_code(double c)
{
    static const int map[10] = {8, 7, 0, 1, 6, 8, 2, 5, 4, 3};
    int k = (int)c;
    if (!(c >= 0.0 && c <= 9.0) || (double)k != c) return -1;
    return map[k];
}
/* IX_ADDS, kinematic_wave_parallel.py:47 */
static const int IX_ADDS[8][2] = {{1, 0}, {1, 1}, {0, 1}, {-1, 1}, {-1, 0}, {-1, -1}, {0, -1}, {1, -1}};

/*
 * Builds every graph member of the reference's kinematicWave object.
 *   ldd_codes      f64[N]   compressed keypad codes (0..9; 0 and 5 = pit)
 *   land_mask      u8[R*C]  1 = active pixel
 * outputs (caller-allocated):
 *   downstream     f64[N]   (-1 = none)              tools:113,126
 *   upstream       i64[N*8] (-1 fill, slot order = row-major order of the source pixel) tools:114,127
 *   num_ups        i64[N]                              kinematic_wave_parallel.py:136
 *   pixels_ordered i64[N], order_start_stop i64[2*n_orders] (caller gives room for 2*N) :148-158
 *   n_orders, max_ups (= number of non-empty upstream columns, >= 1; :89)
 * returns 0, -1 bad code, -2 cycle in the LDD (the reference would loop forever, :99).
 */
int lfo_ldd_graph(const double *ldd_codes, const uint8_t *land_mask, int64_t rows, int64_t cols,
                  double *downstream, int64_t *upstream, int64_t *num_ups, int64_t *pixels_ordered,
                  int64_t *order_start_stop, int64_t *n_orders, int64_t *max_ups)
{
    int64_t ncell = rows * cols, n = 0;
    int64_t *land_points = (int64_t *)malloc(sizeof(int64_t) * (size_t)ncell);
    for (int64_t i = 0; i < ncell; ++i) land_points[i] = land_mask[i] ? n++ : -1; /* :85-86 */
    for (int64_t p = 0; p < n; ++p) {
        downstream[p] = -1.0;
        num_ups[p] = 0;
        for (int k = 0; k < 8; ++k) upstream[p * 8 + k] = -1;
    }
    /* serial double loop in row-major source order: tools:118-129 */
    for (int64_t r = 0; r < rows; ++r)
        for (int64_t c = 0; c < cols; ++c) {
            int64_t src = land_points[r * cols + c];
            if (src < 0) continue; /* off-mask pixels carry direction 8: :84 */
            int d = decode_code(ldd_codes[src]);
            if (d < 0) { free(land_points); return -1; }
            if (d >= 8) continue;
            int64_t rr = r + IX_ADDS[d][0], cc = c + IX_ADDS[d][1];
            if (rr == -1 || cc == -1 || rr == rows || cc == cols || !land_mask[rr * cols + cc]) continue;
            int64_t dst = land_points[rr * cols + cc];
            downstream[src] = (double)dst;
            upstream[dst * 8 + num_ups[dst]] = src;
            num_ups[dst] += 1;
        }
    free(land_points);
    int64_t k = 0;
    for (int64_t p = 0; p < n; ++p) if (num_ups[p] > k) k = num_ups[p];
    *max_ups = k > 1 ? k : 1;

    /* topological distance from the outlets (1 = outlet), O(N) breadth-first instead of the
     * reference's O(N*depth) scan (:92-106); same values. */
    int64_t *dist = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n > 0 ? n : 1));
    int64_t *queue = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n > 0 ? n : 1));
    int64_t head = 0, tail = 0, maxd = 0;
    for (int64_t p = 0; p < n; ++p) {
        dist[p] = -1;
        if (downstream[p] == -1.0) { dist[p] = 1; queue[tail++] = p; }
    }
    while (head < tail) {
        int64_t p = queue[head++];
        if (dist[p] > maxd) maxd = dist[p];
        for (int64_t s = 0; s < num_ups[p]; ++s) {
            int64_t u = upstream[p * 8 + s];
            dist[u] = dist[p] + 1;
            queue[tail++] = u;
        }
    }
    free(queue);
    if (tail != n) { free(dist); return -2; }
    /* routing_order = max - dist; stable sort by (order, pixel) = counting sort (:147-158) */
    int64_t no = maxd;
    *n_orders = no;
    int64_t *count = (int64_t *)calloc((size_t)(no + 1), sizeof(int64_t));
    for (int64_t p = 0; p < n; ++p) count[maxd - dist[p] + 1] += 1;
    for (int64_t o = 0; o < no; ++o) count[o + 1] += count[o];
    for (int64_t o = 0; o < no; ++o) {
        order_start_stop[2 * o] = count[o];
        order_start_stop[2 * o + 1] = count[o + 1];
    }
    for (int64_t p = 0; p < n; ++p) pixels_ordered[count[maxd - dist[p]]++] = p;
    free(count);
    free(dist);
    return 0;
}

/* ------------------------------------------------------------------------------------------------
 * Kinematic wave: tools:34-92
 * ---------------------------------------------------------------------------------------------- */

/* closureError, tools:89-92 */
static inline double closure_error(double q, double upper, double a, double beta)
{
    return q + a * pow(q, beta) - upper;
}

/* solve1Pixel, tools:48-86.  Writes discharge[pix]; returns the Newton iteration count. */
static inline int solve_1_pixel(int64_t pix, double *discharge, const double *constant,
                                const int64_t *upstream, int64_t ups_stride, const int64_t *num_ups,
                                const double *a_dx_div_dt, const double *b_a_dx_div_dt, double beta,
                                double inv_beta, double b_minus_1)
{
    int count = 0;
    double previous = -1.0, upstream_inflow = 0.0;
    for (int64_t k = 0; k < num_ups[pix]; ++k) upstream_inflow += discharge[upstream[pix * ups_stride + k]];
    double c = upstream_inflow + constant[pix];
    if (c <= NEWTON_TOL) { discharge[pix] = 0.0; return 0; }
    double a = a_dx_div_dt[pix], ba = b_a_dx_div_dt[pix];
    double t = ba * pow(c, b_minus_1);
    double secant = (t <= 1.0) ? c / (1.0 + t) : c / (1.0 + pow(t, inv_beta));
    double other = pow((c - secant) / a, inv_beta);
    double q = (secant + other) / 2.0;
    double err = closure_error(q, c, a, beta);
    while (fabs(err) > NEWTON_TOL && q != previous && count < MAX_ITERS) {
        previous = q;
        q -= err / (1.0 + ba * pow(q, b_minus_1));
        q = q > NEWTON_TOL ? q : NEWTON_TOL; /* max(q, NEWTON_TOL) */
        err = closure_error(q, c, a, beta);
        ++count;
    }
    if (q == NEWTON_TOL) q = 0.0;
    discharge[pix] = q;
    return count;
}

/* kinematicRouting, tools:34-46: serial over orders, parallel (prange -> OpenMP) within one. */
/* `fixed` (optional, u8[N]): pixels whose discharge is PRESCRIBED (fixed_values) instead of solved -- the ghost
 * pixels of an LDD-cut partition (tests of lisflood_code_b200/parallel.py); not part of the reference. */
int64_t lfo_kinematic_routing(double *discharge, const double *constant, const int64_t *upstream,
                              int64_t ups_stride, const int64_t *num_ups, const int64_t *ordered,
                              const int64_t *start_stop, int64_t n_orders, double beta,
                              const double *a_dx_div_dt, const double *b_a_dx_div_dt, const uint8_t *fixed,
                              const double *fixed_values)
{
    double inv_beta = 1.0 / beta, b_minus_1 = beta - 1.0;
    int64_t iters = 0;
    int max_thr = 1;
#ifdef _OPENMP
    max_thr = omp_get_max_threads();
#endif
    for (int64_t o = 0; o < n_orders; ++o) {
        int64_t first = start_stop[2 * o], last = start_stop[2 * o + 1];
        /* engage one thread per ~128 pixels of the level so that small levels do not pay a
         * full-team fork/join (Numba's prange chunks similarly) */
        int thr = (int)((last - first) / 128);
        thr = thr < 1 ? 1 : (thr > max_thr ? max_thr : thr);
#pragma omp parallel for schedule(static) reduction(+ : iters) num_threads(thr) if (thr > 1)
        for (int64_t i = first; i < last; ++i) {
            if (fixed && fixed[ordered[i]]) {
                discharge[ordered[i]] = fixed_values[ordered[i]];
                continue;
            }
            iters += solve_1_pixel(ordered[i], discharge, constant, upstream, ups_stride, num_ups, a_dx_div_dt,
                                   b_a_dx_div_dt, beta, inv_beta, b_minus_1);
        }
    }
    return iters;
}

/*
 * kinematicWave.kinematicWaveRouting, kinematic_wave_parallel.py:160-179:
 *   lateral_inflow = q * dx ; constant = a_dx_div_dt * Qold**b + lateral_inflow ; kinematicRouting.
 * dx may be NULL, then dx_scalar is used (space_delta is a scalar for the overland routers,
 * surface_routing.py:108-113).  `work` is f64[N] scratch for `constant`.
 * Returns total Newton iterations (diagnostic).
 */
int64_t lfo_kinematic_wave_routing(double *discharge, const double *q_lat, int64_t n, const double *dx,
                                   double dx_scalar, const double *a_dx_div_dt, const double *b_a_dx_div_dt,
                                   double beta, const int64_t *upstream, int64_t ups_stride,
                                   const int64_t *num_ups, const int64_t *ordered, const int64_t *start_stop,
                                   int64_t n_orders, double *work, const uint8_t *fixed, const double *fixed_values)
{
#pragma omp parallel for schedule(static)
    for (int64_t p = 0; p < n; ++p) {
        double ql = q_lat[p] * (dx ? dx[p] : dx_scalar);
        work[p] = a_dx_div_dt[p] * pow(discharge[p], beta) + ql;
    }
    return lfo_kinematic_routing(discharge, work, upstream, ups_stride, num_ups, ordered, start_stop, n_orders,
                                 beta, a_dx_div_dt, b_a_dx_div_dt, fixed, fixed_values);
}
