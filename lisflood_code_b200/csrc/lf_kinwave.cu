// lf_kinwave.cu -- kinematic-wave router: class kinematicWave of the reference
// (hydrological_modules/kinematic_wave_parallel.py:114-184) on device-resident state.
//
// Storage: every per-pixel array lives in the graph's breadth-first "position" order
// (lf_graph.cu step 7).  In that order
//   * routing level l is the contiguous span [level_start[l], level_start[l+1]),
//   * the upstream pixels of position i are the contiguous positions cfirst[i]..cfirst[i+1]-1 of
//     level l-1, already in the reference's slot order -- the gather is a short contiguous read,
//   * every link goes from level l to level l+1, so the work item (level l, step s) depends only
//     on (l-1, s) and (l, s-1).  All items with l + s = d are independent: a run of S routing steps
//     is executed as L+S-1 "diagonal" launches instead of L*S level launches, each over ONE
//     contiguous span of positions (the levels d-S+1..d), with discharge double-buffered by step
//     parity.  S = 1 degenerates to the reference's level-by-level sweep.
#include <string.h>

#include <vector>

#include "lf_common.cuh"
#include "lf_kw_solve.cuh"
#include "lf_xchg.cuh"

extern "C" int lf_xchg_begin(lf_xchg *, int32_t *);
extern "C" int lf_xchg_end(lf_xchg *);
extern "C" int lf_xchg_peer_base(lf_xchg *, int32_t, uint64_t *);

struct lf_router {
    lf_graph *g = nullptr;
    int64_t n = 0;
    lfkw::Params P;
    double dt = 0, dx_scalar = 0;
    bool dx_is_array = false, has_fp = false;
    int nancheck = 0;
    lf::DevBuf<double> a[2];     // a_dx_div_dt per section            [position order]
    lf::DevBuf<double> dx;       // space_delta if it is a map         [position order]
    lf::DevBuf<double> Q[2][2];  // per section: ping-pong discharge   [position order]
    lf::DevBuf<double> q[2];     // specific lateral inflow            [position order]
    int64_t steps_done[2] = {0, 0};
    lf::DevBuf<double> stage_a, stage_b;  // compressed (user) order staging
    lf::DevBuf<double> scale;
    lf::DevBuf<int> flag;
    // LDD-cut exchange (multi-GPU, lf_xchg.cuh): per position -1 = plain pixel, k >= 0 = export edge k (its new value of
    // every step is also stored into the consumer rank's region), <= -2 = ghost of a pixel owned by another rank (value
    // taken from this rank's import block, not solved)
    lf::DevBuf<int32_t> xslot;
    lf_xchg *xchg = nullptr;
    lf::DevBuf<double *> exp_dst;
    lf::DevBuf<long long> exp_stride;
    double *imp = nullptr;
    long long imp_parity_stride = 0;
    int32_t x_cap_steps = 0, n_export = 0, n_import = 0;
    lf::GraphCache graphs;   // CUDA-graph replay of the diagonals of a run (one variant per argument set)
    int use_graphs = 1, use_coop = 0;   // lf_router_set_option; the cooperative launch is off by default (measured slower)
    int narrow_runs = 0;                // runs of narrow diagonals in one single-block launch (option; measured: no gain)
    lf::DevBuf<unsigned int> coop_counter;
};

namespace {

// to position order: dst[i] = src[pix_of_pos[i]]
__global__ void k_to_pos(const double *__restrict__ src, double *__restrict__ dst,
                         const int32_t *__restrict__ pix_of_pos, int64_t n)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[pix_of_pos[i]];
}
// discharge in: compressed order -> position order, Q -> z = Q^(1/5) when the router runs in 3/5 mode
__global__ void k_q_to_pos(const double *__restrict__ src, double *__restrict__ dst,
                           const int32_t *__restrict__ pix_of_pos, int64_t n, int quintic)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double q = src[pix_of_pos[i]];
    dst[i] = quintic ? lfkw::z_of_q(q) : q;
}
// discharge out: position order -> compressed order, z -> Q = z^5
__global__ void k_q_to_pix(const double *__restrict__ src, double *__restrict__ dst,
                           const int32_t *__restrict__ pos_of_pix, int64_t n, int quintic)
{
    int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    double v = src[pos_of_pix[p]];
    dst[p] = quintic ? lfkw::pow5(v) : v;
}
// alpha -> a_dx_div_dt = alpha * dx / dt   (kinematic_wave_parallel.py:126)
__global__ void k_make_a(const double *__restrict__ alpha, const double *__restrict__ dx, double dx_scalar,
                         double dt, double *__restrict__ a, double *__restrict__ dx_pos,
                         const int32_t *__restrict__ pix_of_pos, int64_t n)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int p = pix_of_pos[i];
    double d = dx ? dx[p] : dx_scalar;
    a[i] = alpha[p] * d / dt;
    if (dx_pos) dx_pos[i] = d;
}
// back to compressed order: dst[p] = src[pos_of_pix[p]]
__global__ void k_to_pix(const double *__restrict__ src, double *__restrict__ dst,
                         const int32_t *__restrict__ pos_of_pix, int64_t n)
{
    int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p < n) dst[p] = src[pos_of_pix[p]];
}
__global__ void k_nonfinite(const double *__restrict__ v, int64_t n, int *__restrict__ flag)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && !isfinite(v[i])) *flag = 1;
}

// One work item (position i, diagonal d) of the space-time wavefront.
struct KwArgs {
    int64_t g0;
    const int32_t *lev, *cfirst, *cend;
    const double *a, *dx, *q, *scale;
    double dx_scalar;
    double *Q0, *Q1;
    lfkw::Params P;
    lfx::View X;
};
// COOP: called from the persistent kernel, where values written by other SMs earlier in the SAME launch are read: the
// discharge buffers are then loaded through L2 (ld.global.cg), never from a possibly stale L1 line.
template <bool QZ, bool HASX, bool COOP = false>
__device__ __forceinline__ void kw_item(const KwArgs &A, int i, int d)
{
    int s = d - A.lev[i];
    int64_t gs = A.g0 + s;  // global index of this routing step; parity selects the buffer
    double *Qnew = (gs & 1) ? A.Q1 : A.Q0;
    const double *Qold = (gs & 1) ? A.Q0 : A.Q1;
    int xs = -1;
    if (HASX) {
        xs = A.X.xslot[i];
        if (xs <= -2) {  // ghost: the owner's value of this step (router-native representation)
            Qnew[i] = xs == lfx::INERT ? 0.0 : lfx::take(lfx::import_slot(A.X, -2 - xs, 0, s), A.X.abort_flag);
            return;
        }
    }
    int c0 = A.cfirst[i], c1 = A.cend ? A.cend[i] : A.cfirst[i + 1];
    double qo = COOP ? __ldcg(Qold + i) : Qold[i];
    double qs = A.q[i];
    if (A.scale) qs *= A.scale[s];
    double lateral = qs * (A.dx ? A.dx[i] : A.dx_scalar);  // lateral_inflow = q * dx, kinematic_wave_parallel.py:163
    double ai = A.a[i];
    double U = 0.0;
    double out;
    if (QZ) {
        for (int k = c0; k < c1; ++k) U += lfkw::pow5(COOP ? __ldcg(Qnew + k) : Qnew[k]);  // buffers hold z = Q^(1/5)
        out = lfkw::solve_z(U, qo, lateral, ai);
    } else {
        for (int k = c0; k < c1; ++k) U += COOP ? __ldcg(Qnew + k) : Qnew[k];  // upstream discharge of this step, slot order (tools:57-58)
        out = lfkw::solve(U, qo, lateral, ai, A.P);
    }
    Qnew[i] = out;
    if (HASX && xs >= 0) lfx::push(lfx::export_slot(A.X, xs, 0, s), out);
}

// Grid-wide barrier of a persistent kernel whose blocks are all resident (cooperative launch): a monotonically growing
// arrival counter; barrier number k is passed when it reaches (k + 1) * gridDim.x.  One L2 atomic and a poll per block
// (cooperative_groups' grid.sync() measured ~10 us per barrier here, more than a kernel boundary in a CUDA graph).
__device__ __forceinline__ void grid_barrier(unsigned int *counter, unsigned int target)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();                       // this block's stores are visible device-wide before it arrives
        atomicAdd(counter, 1u);
        unsigned int seen;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(counter) : "memory");
        } while (seen < target);
    }
    __syncthreads();
}

// One diagonal per launch.  Positions [lo, hi) = levels d-S+1..d.
constexpr int KW_THREADS = 128;
template <bool QZ, bool HASX>
__global__ void __launch_bounds__(KW_THREADS) k_kw_diagonal(int lo, int hi, int d, KwArgs A)
{
    int i = lo + blockIdx.x * KW_THREADS + threadIdx.x;
    if (i >= hi) return;
    kw_item<QZ, HASX>(A, i, d);
}

// A run of consecutive NARROW diagonals [d0, d1) -- at most KW_RUN_THREADS items each -- in ONE block: a __syncthreads()
// between two diagonals instead of a kernel boundary.  Values written by the block's other threads in earlier diagonals
// are visible after the barrier; ghost pixels poll peer memory as in k_kw_diagonal.  An OPTION ("narrow_runs"), off by
// default: the reference's ordering counts levels down from the distance to the outlet, so a deep basin is narrow only in
// its first few hundred diagonals (the far ends of its longest paths); measured on C2 deep (13.0 ms either way) and C4
// (38 of 28 589 diagonals merged, 809 ms either way).
constexpr int KW_RUN_THREADS = 512;
template <bool QZ, bool HASX>
__global__ void __launch_bounds__(KW_RUN_THREADS) k_kw_narrow_run(int d0, int d1, int nlev, int nsteps,
                                                                  const int32_t *__restrict__ level_start, KwArgs A)
{
    for (int d = d0; d < d1; ++d) {
        const int lo_lev = d - nsteps + 1 > 0 ? d - nsteps + 1 : 0;
        const int hi_lev = d < nlev - 1 ? d : nlev - 1;
        const int i = level_start[lo_lev] + threadIdx.x;
        if (i < level_start[hi_lev + 1]) kw_item<QZ, HASX>(A, i, d);
        __syncthreads();
    }
}

// The whole run in ONE cooperative launch: a persistent grid (every block resident) walks the diagonals and meets at a
// grid-wide barrier after each one.  Built for deep networks (tens of thousands of levels: a diagonal is only a few
// microseconds of work) and MEASURED SLOWER than the CUDA-graph replay of one kernel per diagonal on B200 -- C2 deep
// (2102 diagonals): graph replay 13.0 ms per 100 steps, this kernel 17.2 ms with the hand-written barrier below and
// 23.3 ms with cooperative_groups' grid.sync() (profiles/r02_router_coop_sweep.txt) -- so it is an option
// (lf_router_set_option("cooperative", blocks per SM)), off by default.
constexpr int KW_COOP_THREADS = 1024;   // one fat block per SM: the grid barrier costs grow with the number of blocks
template <bool QZ, bool HASX>
__global__ void __launch_bounds__(KW_COOP_THREADS) k_kw_wavefront_coop(int nlev, int nsteps, const int32_t *__restrict__ level_start,
                                                                        KwArgs A, unsigned int *counter)
{
    const int stride = gridDim.x * KW_COOP_THREADS;
    unsigned int target = 0;
    for (int d = 0; d < nlev + nsteps - 1; ++d) {
        const int lo_lev = d - nsteps + 1 > 0 ? d - nsteps + 1 : 0;
        const int hi_lev = d < nlev - 1 ? d : nlev - 1;
        const int lo = level_start[lo_lev], hi = level_start[hi_lev + 1];
        for (int i = lo + blockIdx.x * KW_COOP_THREADS + threadIdx.x; i < hi; i += stride) kw_item<QZ, HASX, true>(A, i, d);
        target += gridDim.x;
        grid_barrier(counter, target);
    }
}

__global__ void k_i32_to_pos(const int32_t *__restrict__ src, int32_t *__restrict__ dst,
                             const int32_t *__restrict__ pix_of_pos, int64_t n)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[pix_of_pos[i]];
}

int run_steps(lf_router *r, int sec, int nsteps, const double *d_scale)
{
    if (r->xslot.p && nsteps > r->x_cap_steps) {
        lf::set_error("lf_router_run: %d steps exceed the exchange buffers (%d steps)", nsteps, r->x_cap_steps);
        return LF_ERR_INVALID;
    }
    lf_graph *g = r->g;
    cudaStream_t st = lf::stream();
    const std::vector<int32_t> &ls = g->h_level_start;
    int L = g->n_orders;
    int64_t g0 = r->steps_done[sec];
    lfx::View X;
    memset(&X, 0, sizeof(X));
    if (r->xslot.p) {
        int32_t parity = 0;
        LF_CHECK(lf_xchg_begin(r->xchg, &parity));
        X.xslot = r->xslot.p;
        X.exp_dst = r->exp_dst.p;
        X.exp_stride = r->exp_stride.p;
        X.imp = r->imp;
        X.imp_parity_stride = r->imp_parity_stride;
        X.cap = r->x_cap_steps;
        X.nsec = 1;
        X.parity = parity;
        LF_CHECK(lf::xchg_view_base(r->xchg, &X.abort_flag));
    }
    KwArgs A;
    memset(&A, 0, sizeof(A));
    A.g0 = g0 & 1;   // only the parity matters to the kernels: equal arguments across runs keep the graph cache warm
    A.lev = g->lev_of_pos.p;
    A.cfirst = g->cfirst.p;
    A.cend = g->cend.p;
    A.a = r->a[sec].p;
    A.dx = r->dx_is_array ? r->dx.p : nullptr;
    A.q = r->q[sec].p;
    A.scale = d_scale;
    A.dx_scalar = r->dx_scalar;
    A.Q0 = r->Q[sec][0].p;
    A.Q1 = r->Q[sec][1].p;
    A.P = r->P;
    A.X = X;
    const bool hasx = r->xslot.p != nullptr;
    const int ndiag = L + nsteps - 1;
    // deep network, little work per diagonal: one persistent cooperative launch with grid barriers
    bool coop = r->use_coop && ndiag > 64 && (double)r->n * nsteps / ndiag < 4.0e6;
    if (coop) {
        void *fn = r->P.quintic ? (hasx ? (void *)k_kw_wavefront_coop<true, true> : (void *)k_kw_wavefront_coop<true, false>)
                                : (hasx ? (void *)k_kw_wavefront_coop<false, true> : (void *)k_kw_wavefront_coop<false, false>);
        int per_sm = 0;
        LF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, KW_COOP_THREADS, 0));
        if (per_sm < 1) {
            coop = false;
        } else {
            const int64_t widest = (int64_t)r->n * nsteps / ndiag * 4 + KW_COOP_THREADS;   // blocks beyond the widest diagonals idle
            const unsigned grid = (unsigned)std::max<int64_t>(1, std::min<int64_t>((int64_t)std::min(per_sm, r->use_coop) * lf::sm_count(),
                                                                                 widest / KW_COOP_THREADS));
            int nlev = L, ns = nsteps;
            const int32_t *lsd = g->level_start.p;
            if (!r->coop_counter.p) LF_CHECK(r->coop_counter.alloc(1));
            LF_CUDA(cudaMemsetAsync(r->coop_counter.p, 0, sizeof(unsigned int), st));
            unsigned int *ctr = r->coop_counter.p;
            void *args[] = {&nlev, &ns, &lsd, &A, &ctr};
            LF_CUDA(cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(KW_COOP_THREADS), args, 0, st));
            lf::count_launch();
        }
    }
    auto width = [&](int d) {
        int lo_lev = d - nsteps + 1 > 0 ? d - nsteps + 1 : 0;
        int hi_lev = d < L - 1 ? d : L - 1;
        return ls[hi_lev + 1] - ls[lo_lev];
    };
    const int32_t *lsd = g->level_start.p;
    auto diagonals = [&]() -> int {
        for (int d = 0; d < ndiag; ++d) {
            if (r->narrow_runs && width(d) <= KW_RUN_THREADS) {   // a run of narrow diagonals: one single-block launch
                int e = d + 1;
                while (e < ndiag && width(e) <= KW_RUN_THREADS) ++e;
                if (e - d >= 3) {
                    if (r->P.quintic) {
                        if (hasx) k_kw_narrow_run<true, true><<<1, KW_RUN_THREADS, 0, st>>>(d, e, L, nsteps, lsd, A);
                        else k_kw_narrow_run<true, false><<<1, KW_RUN_THREADS, 0, st>>>(d, e, L, nsteps, lsd, A);
                    } else {
                        if (hasx) k_kw_narrow_run<false, true><<<1, KW_RUN_THREADS, 0, st>>>(d, e, L, nsteps, lsd, A);
                        else k_kw_narrow_run<false, false><<<1, KW_RUN_THREADS, 0, st>>>(d, e, L, nsteps, lsd, A);
                    }
                    LF_LAUNCH_CHECK();
                    d = e - 1;
                    continue;
                }
            }
            int lo_lev = d - nsteps + 1 > 0 ? d - nsteps + 1 : 0;
            int hi_lev = d < L - 1 ? d : L - 1;
            int lo = ls[lo_lev], hi = ls[hi_lev + 1];
            if (hi <= lo) continue;
            const unsigned gb = lf::blocks_for(hi - lo, KW_THREADS);
            if (r->P.quintic) {
                if (hasx) k_kw_diagonal<true, true><<<gb, KW_THREADS, 0, st>>>(lo, hi, d, A);
                else k_kw_diagonal<true, false><<<gb, KW_THREADS, 0, st>>>(lo, hi, d, A);
            } else {
                if (hasx) k_kw_diagonal<false, true><<<gb, KW_THREADS, 0, st>>>(lo, hi, d, A);
                else k_kw_diagonal<false, false><<<gb, KW_THREADS, 0, st>>>(lo, hi, d, A);
            }
            LF_LAUNCH_CHECK();
        }
        return LF_OK;
    };
    if (!coop) {
        // key of the captured graph: everything the launches depend on
        struct Key {
            int sec, nsteps;
            KwArgs A;
        } key;
        memset(&key, 0, sizeof(key));
        key.sec = sec;
        key.nsteps = nsteps;
        key.A = A;
        if (r->use_graphs && ndiag > 8) LF_CHECK(lf::run_captured(r->graphs, &key, sizeof(key), st, diagonals));
        else LF_CHECK(diagonals());
    }
    r->steps_done[sec] = g0 + nsteps;
    if (r->xslot.p) LF_CHECK(lf_xchg_end(r->xchg));
    return LF_OK;
}

inline double *current_q(lf_router *r, int sec) { return r->Q[sec][(r->steps_done[sec] + 1) & 1].p; }

int check_section(lf_router *r, int section, const char *fn)
{
    if (!r) {
        lf::set_error("%s: null router", fn);
        return LF_ERR_INVALID;
    }
    if (section != LF_SECTION_MAIN && section != LF_SECTION_FLOODPLAIN) {
        // the reference raises Exception("The section parameter must be either 'main_channel' or 'floodplain'!")
        lf::set_error("%s: the section parameter must be either 'main_channel' or 'floodplain'", fn);
        return LF_ERR_INVALID;
    }
    if (section == LF_SECTION_FLOODPLAIN && !r->has_fp) {
        lf::set_error("%s: router was created without alpha_floodplains", fn);
        return LF_ERR_STATE;
    }
    return LF_OK;
}

int nan_check(lf_router *r, int sec, int *nonfinite)
{
    if (nonfinite) *nonfinite = 0;
    if (!r->nancheck) return LF_OK;
    cudaStream_t st = lf::stream();
    LF_CUDA(cudaMemsetAsync(r->flag.p, 0, sizeof(int), st));
    k_nonfinite<<<lf::blocks_for(r->n, 256), 256, 0, st>>>(current_q(r, sec), r->n, r->flag.p);
    LF_LAUNCH_CHECK();
    int h = 0;
    LF_CUDA(cudaMemcpyAsync(&h, r->flag.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    LF_CUDA(cudaStreamSynchronize(st));
    if (nonfinite) *nonfinite = h;
    return LF_OK;
}

}  // namespace

extern "C" {

int lf_router_create(lf_graph *g, const double *alpha, double beta, const double *dx, double dx_scalar, double dt,
                     const double *alpha_floodplains, int flagnancheck, lf_router **out)
{
    if (!g || !alpha || !out) {
        lf::set_error("lf_router_create: null graph / alpha / out");
        return LF_ERR_INVALID;
    }
    if (!(dt > 0) || !(beta > 0)) {
        lf::set_error("lf_router_create: beta and time_delta must be positive");
        return LF_ERR_INVALID;
    }
    LF_CHECK(lf::ensure_device());
    cudaStream_t st = lf::stream();
    lf_router *r = new lf_router();
    *out = nullptr;
    r->g = g;
    r->n = g->n;
    r->P = lfkw::make_params(beta);
    r->dt = dt;
    r->dx_scalar = dx_scalar;
    r->dx_is_array = dx != nullptr;
    r->has_fp = alpha_floodplains != nullptr;
    r->nancheck = flagnancheck;
    int64_t n = r->n;
    int rc = LF_OK;
    auto fail = [&](int code) {
        delete r;
        return code;
    };
    if ((rc = r->stage_a.alloc(n)) || (rc = r->stage_b.alloc(n)) || (rc = r->flag.alloc(1))) return fail(rc);
    if (r->dx_is_array && (rc = r->dx.alloc(n))) return fail(rc);
    int nsec = r->has_fp ? 2 : 1;
    for (int s = 0; s < nsec; ++s) {
        if ((rc = r->a[s].alloc(n)) || (rc = r->Q[s][0].alloc(n)) || (rc = r->Q[s][1].alloc(n)) ||
            (rc = r->q[s].alloc(n)))
            return fail(rc);
        cudaError_t ez = cudaMemsetAsync(r->Q[s][0].p, 0, n * sizeof(double), st);
        if (ez == cudaSuccess) ez = cudaMemsetAsync(r->Q[s][1].p, 0, n * sizeof(double), st);
        if (ez == cudaSuccess) ez = cudaMemsetAsync(r->q[s].p, 0, n * sizeof(double), st);
        if (ez != cudaSuccess) {
            lf::set_error("lf_router_create: clearing the discharge buffers failed: %s", cudaGetErrorString(ez));
            return fail(LF_ERR_CUDA);
        }
    }
    if (r->dx_is_array) {
        cudaError_t e = cudaMemcpyAsync(r->stage_b.p, dx, n * sizeof(double), cudaMemcpyDefault, st);
        if (e != cudaSuccess) {
            lf::set_error("lf_router_create: copy of space_delta failed: %s", cudaGetErrorString(e));
            return fail(LF_ERR_CUDA);
        }
    }
    for (int s = 0; s < nsec; ++s) {
        const double *al = s == 0 ? alpha : alpha_floodplains;
        cudaError_t e = cudaMemcpyAsync(r->stage_a.p, al, n * sizeof(double), cudaMemcpyDefault, st);
        if (e != cudaSuccess) {
            lf::set_error("lf_router_create: copy of alpha failed: %s", cudaGetErrorString(e));
            return fail(LF_ERR_CUDA);
        }
        k_make_a<<<lf::blocks_for(n, 256), 256, 0, st>>>(r->stage_a.p, r->dx_is_array ? r->stage_b.p : nullptr,
                                                         dx_scalar, dt, r->a[s].p,
                                                         (s == 0 && r->dx_is_array) ? r->dx.p : nullptr,
                                                         g->pix_of_pos.p, n);
        lf::count_launch();
    }
    cudaError_t e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) {
        lf::set_error("lf_router_create: %s", cudaGetErrorString(e));
        return fail(LF_ERR_CUDA);
    }
    *out = r;
    return LF_OK;
}

int lf_router_set_discharge(lf_router *r, int section, const double *discharge)
{
    LF_CHECK(check_section(r, section, "lf_router_set_discharge"));
    if (!discharge) {
        lf::set_error("lf_router_set_discharge: null pointer");
        return LF_ERR_INVALID;
    }
    LF_CHECK(lf::ensure_device());
    cudaStream_t st = lf::stream();
    LF_CUDA(cudaMemcpyAsync(r->stage_a.p, discharge, r->n * sizeof(double), cudaMemcpyDefault, st));
    k_q_to_pos<<<lf::blocks_for(r->n, 256), 256, 0, st>>>(r->stage_a.p, current_q(r, section), r->g->pix_of_pos.p, r->n,
                                                          r->P.quintic);
    LF_LAUNCH_CHECK();
    LF_CUDA(cudaStreamSynchronize(st));
    return LF_OK;
}

int lf_router_get_discharge(lf_router *r, int section, double *discharge)
{
    LF_CHECK(check_section(r, section, "lf_router_get_discharge"));
    if (!discharge) {
        lf::set_error("lf_router_get_discharge: null pointer");
        return LF_ERR_INVALID;
    }
    LF_CHECK(lf::ensure_device());
    cudaStream_t st = lf::stream();
    k_q_to_pix<<<lf::blocks_for(r->n, 256), 256, 0, st>>>(current_q(r, section), r->stage_a.p, r->g->pos_of_pix.p, r->n,
                                                          r->P.quintic);
    LF_LAUNCH_CHECK();
    LF_CUDA(cudaMemcpyAsync(discharge, r->stage_a.p, r->n * sizeof(double), cudaMemcpyDefault, st));
    LF_CUDA(cudaStreamSynchronize(st));
    return LF_OK;
}

int lf_router_set_inflow(lf_router *r, int section, const double *specific_lateral_inflow)
{
    LF_CHECK(check_section(r, section, "lf_router_set_inflow"));
    if (!specific_lateral_inflow) {
        lf::set_error("lf_router_set_inflow: null pointer");
        return LF_ERR_INVALID;
    }
    LF_CHECK(lf::ensure_device());
    cudaStream_t st = lf::stream();
    LF_CUDA(cudaMemcpyAsync(r->stage_b.p, specific_lateral_inflow, r->n * sizeof(double), cudaMemcpyDefault, st));
    k_to_pos<<<lf::blocks_for(r->n, 256), 256, 0, st>>>(r->stage_b.p, r->q[section].p, r->g->pix_of_pos.p, r->n);
    LF_LAUNCH_CHECK();
    LF_CUDA(cudaStreamSynchronize(st));
    return LF_OK;
}

int lf_router_run(lf_router *r, int section, int nsteps, const double *q_scale, int *nonfinite)
{
    LF_CHECK(check_section(r, section, "lf_router_run"));
    if (nsteps < 0) {
        lf::set_error("lf_router_run: nsteps must be >= 0");
        return LF_ERR_INVALID;
    }
    if (nonfinite) *nonfinite = 0;
    if (nsteps == 0) return LF_OK;
    LF_CHECK(lf::ensure_device());
    cudaStream_t st = lf::stream();
    const double *d_scale = nullptr;
    if (q_scale) {
        if (r->scale.n < (size_t)nsteps) LF_CHECK(r->scale.alloc(nsteps));
        LF_CUDA(cudaMemcpyAsync(r->scale.p, q_scale, nsteps * sizeof(double), cudaMemcpyDefault, st));
        d_scale = r->scale.p;
    }
    LF_CHECK(run_steps(r, section, nsteps, d_scale));
    LF_CHECK(nan_check(r, section, nonfinite));
    if (q_scale) LF_CUDA(cudaStreamSynchronize(st));  // q_scale is borrowed only for the duration of the call
    return LF_OK;
}

int lf_router_route(lf_router *r, double *discharge, const double *specific_lateral_inflow, int section,
                    int *nonfinite)
{
    LF_CHECK(check_section(r, section, "lf_router_route"));
    if (!discharge || !specific_lateral_inflow) {
        lf::set_error("lf_router_route: null pointer");
        return LF_ERR_INVALID;
    }
    LF_CHECK(lf::ensure_device());
    cudaStream_t st = lf::stream();
    int64_t n = r->n;
    LF_CUDA(cudaMemcpyAsync(r->stage_a.p, discharge, n * sizeof(double), cudaMemcpyDefault, st));
    LF_CUDA(cudaMemcpyAsync(r->stage_b.p, specific_lateral_inflow, n * sizeof(double), cudaMemcpyDefault, st));
    k_q_to_pos<<<lf::blocks_for(n, 256), 256, 0, st>>>(r->stage_a.p, current_q(r, section), r->g->pix_of_pos.p, n,
                                                       r->P.quintic);
    LF_LAUNCH_CHECK();
    k_to_pos<<<lf::blocks_for(n, 256), 256, 0, st>>>(r->stage_b.p, r->q[section].p, r->g->pix_of_pos.p, n);
    LF_LAUNCH_CHECK();
    LF_CHECK(run_steps(r, section, 1, nullptr));
    k_q_to_pix<<<lf::blocks_for(n, 256), 256, 0, st>>>(current_q(r, section), r->stage_a.p, r->g->pos_of_pix.p, n,
                                                       r->P.quintic);
    LF_LAUNCH_CHECK();
    LF_CUDA(cudaMemcpyAsync(discharge, r->stage_a.p, n * sizeof(double), cudaMemcpyDefault, st));
    LF_CHECK(nan_check(r, section, nonfinite));
    LF_CUDA(cudaStreamSynchronize(st));
    return LF_OK;
}

int lf_router_set_exchange(lf_router *r, lf_xchg *x, const int32_t *xslot, int32_t n_export, const int32_t *export_peer,
                           const int64_t *export_offset, const int64_t *export_parity_stride, int32_t n_import,
                           int64_t import_offset, int32_t cap_steps)
{
    if (!r || !x || !xslot || cap_steps < 1 || n_export < 0 || n_import < 0 ||
        (n_export > 0 && (!export_peer || !export_offset || !export_parity_stride)) || import_offset < 0) {
        lf::set_error("lf_router_set_exchange: bad arguments");
        return LF_ERR_INVALID;
    }
    LF_CHECK(lf::ensure_device());
    cudaStream_t st = lf::stream();
    lf::DevBuf<int32_t> tmp;
    LF_CHECK(tmp.alloc(r->n));
    LF_CHECK(r->xslot.alloc(r->n));
    LF_CUDA(cudaMemcpyAsync(tmp.p, xslot, r->n * sizeof(int32_t), cudaMemcpyDefault, st));
    k_i32_to_pos<<<lf::blocks_for(r->n, 256), 256, 0, st>>>(tmp.p, r->xslot.p, r->g->pix_of_pos.p, r->n);
    LF_LAUNCH_CHECK();
    std::vector<double *> dst(std::max(n_export, 1), nullptr);
    std::vector<long long> stride(std::max(n_export, 1), 0);
    for (int k = 0; k < n_export; ++k) {
        uint64_t base = 0;
        LF_CHECK(lf_xchg_peer_base(x, export_peer[k], &base));
        dst[k] = (double *)(uintptr_t)(base + lfx::HEADER_BYTES) + export_offset[k];
        stride[k] = export_parity_stride[k];
    }
    LF_CHECK(r->exp_dst.alloc(dst.size()));
    LF_CHECK(r->exp_stride.alloc(stride.size()));
    LF_CUDA(cudaMemcpyAsync(r->exp_dst.p, dst.data(), dst.size() * sizeof(double *), cudaMemcpyHostToDevice, st));
    LF_CUDA(cudaMemcpyAsync(r->exp_stride.p, stride.data(), stride.size() * sizeof(long long), cudaMemcpyHostToDevice, st));
    LF_CUDA(cudaStreamSynchronize(st));
    uint64_t own = 0;
    LF_CHECK(lf_xchg_peer_base(x, -1, &own));
    r->imp = (double *)(uintptr_t)(own + lfx::HEADER_BYTES) + import_offset;
    r->imp_parity_stride = (long long)n_import * cap_steps;
    r->xchg = x;
    r->x_cap_steps = cap_steps;
    r->n_export = n_export;
    r->n_import = n_import;
    return LF_OK;
}

int lf_router_set_option(lf_router *r, const char *name, double value)
{
    if (!r || !name) {
        lf::set_error("lf_router_set_option: null pointer");
        return LF_ERR_INVALID;
    }
    if (strcmp(name, "cuda_graphs") == 0) r->use_graphs = value != 0;
    else if (strcmp(name, "narrow_runs") == 0) {
        r->narrow_runs = value != 0;
        r->graphs.clear();
    }
    else if (strcmp(name, "cooperative") == 0) r->use_coop = (int)value;   // resident blocks per SM of the persistent grid, 0 = off
    else {
        lf::set_error("lf_router_set_option: unknown option '%s'", name);
        return LF_ERR_INVALID;
    }
    return LF_OK;
}

void lf_router_destroy(lf_router *r) { delete r; }

}  // extern "C"
