// lf_graph.cu -- D8 drainage graph and routing order on the GPU.
//
// Replaces the graph part of kinematicWave.__init__ (reference:
// hydrological_modules/kinematic_wave_parallel.py:59-106,134-158 and
// kinematic_wave_parallel_tools.py:111-130).  The reference builds the lookups with a serial
// double loop over the raster and the routing order with an O(N*depth) scan; here every stage is a
// data-parallel pass:
//   1. exclusive scan of the land mask            -> compressed index of every cell
//   2. decode keypad codes to direction indices    -> dir2d (u8 raster)
//   3. downstream pixel of every pixel             -> int32[N]
//   4. upstream count by inspecting the 8 neighbours' codes on a shared-memory tile with halo
//      (gather form: slot order == neighbour scan order NW,N,NE,W,E,SW,S,SE == row-major order of
//      the source pixel, which is what tools:119-129 produces)
//   5. hops-to-outlet by pointer jumping (log2(depth) rounds), cycle detection for free
//   6. routing order = max(dist) - dist; stable radix sort by order -> pixels_ordered (bit-exact)
//   7. breadth-first storage layout from the outlets: children of a pixel are contiguous and in
//      slot order, so the routing kernels gather upstream discharge from a contiguous range.
// CUB (shipped with the CUDA toolkit) is used for the init-time scan / radix sort only.
#include <cub/cub.cuh>

#include <algorithm>
#include <functional>
#include <memory>
#include <queue>

#include "lf_common.cuh"

namespace {

__constant__ int c_ix_adds[8][2] = {{1, 0}, {1, 1}, {0, 1}, {-1, 1}, {-1, 0}, {-1, -1}, {0, -1}, {1, -1}};
// neighbour scan order NW,N,NE,W,E,SW,S,SE and the direction index a neighbour must carry to drain into me
__constant__ int c_nb_dr[8] = {-1, -1, -1, 0, 0, 1, 1, 1};
__constant__ int c_nb_dc[8] = {-1, 0, 1, -1, 1, -1, 0, 1};
__constant__ int c_nb_need[8] = {1, 0, 7, 2, 6, 3, 4, 5};

struct MaskToInt {
    __host__ __device__ int32_t operator()(const uint8_t &m) const { return m ? 1 : 0; }
};

__global__ void k_land_points(const uint8_t *__restrict__ mask, int32_t *__restrict__ lp,
                              int32_t *__restrict__ cell_of_pix, int64_t ncell)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ncell) return;
    if (mask[i]) {
        cell_of_pix[lp[i]] = (int32_t)i;
    } else {
        lp[i] = -1;
    }
}

// keypad code -> direction index (FLOW_CODE = [2,3,6,9,8,7,4,1,5]; 0 = sea -> pit):
// kinematic_wave_parallel.py:47-51,64-71
__global__ void k_decode(const double *__restrict__ codes, const int32_t *__restrict__ cell_of_pix,
                         uint8_t *__restrict__ dir2d, int64_t n, int *__restrict__ bad)
{
    int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const uint8_t map[10] = {8, 7, 0, 1, 6, 8, 2, 5, 4, 3};
    double c = codes[p];
    int k = (c >= 0.0 && c <= 9.0) ? (int)c : -1;
    if (k < 0 || (double)k != c) {
        atomicExch(bad, 1);
        k = 5;
    }
    dir2d[cell_of_pix[p]] = map[k];
}

// downstream pixel: a link exists only if the D8 target is inside the array and on the land mask
// (kinematic_wave_parallel_tools.py:120-124)
__global__ void k_downstream(const uint8_t *__restrict__ dir2d, const int32_t *__restrict__ lp,
                             const int32_t *__restrict__ cell_of_pix, int32_t *__restrict__ ds, int64_t n,
                             int rows, int cols)
{
    int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    int cell = cell_of_pix[p];
    int r = cell / cols, c = cell - r * cols;
    int d = dir2d[cell];
    int out = -1;
    if (d < 8) {
        int rr = r + c_ix_adds[d][0], cc = c + c_ix_adds[d][1];
        if (rr >= 0 && cc >= 0 && rr < rows && cc < cols) {
            int t = rr * cols + cc;
            if (dir2d[t] != 9) out = lp[t];
        }
    }
    ds[p] = out;
}

// Upstream count per pixel from a (TILE_R+2) x (TILE_C+2) shared-memory tile of direction codes.
constexpr int TILE_C = 64, TILE_R = 16;
__global__ void __launch_bounds__(TILE_C *TILE_R)
    k_count_upstream(const uint8_t *__restrict__ dir2d, const int32_t *__restrict__ lp,
                     uint8_t *__restrict__ nups, int rows, int cols)
{
    __shared__ uint8_t tile[TILE_R + 2][TILE_C + 2 + 2];
    int c0 = blockIdx.x * TILE_C, r0 = blockIdx.y * TILE_R;
    int tid = threadIdx.y * TILE_C + threadIdx.x;
    for (int i = tid; i < (TILE_R + 2) * (TILE_C + 2); i += TILE_C * TILE_R) {
        int tr = i / (TILE_C + 2), tc = i - tr * (TILE_C + 2);
        int r = r0 + tr - 1, c = c0 + tc - 1;
        uint8_t v = 9;
        if (r >= 0 && c >= 0 && r < rows && c < cols) v = dir2d[(int64_t)r * cols + c];
        tile[tr][tc] = v;
    }
    __syncthreads();
    int r = r0 + threadIdx.y, c = c0 + threadIdx.x;
    if (r >= rows || c >= cols) return;
    int me = tile[threadIdx.y + 1][threadIdx.x + 1];
    if (me == 9) return;
    int cnt = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        int v = tile[threadIdx.y + 1 + c_nb_dr[k]][threadIdx.x + 1 + c_nb_dc[k]];
        cnt += (v == c_nb_need[k]);
    }
    nups[lp[(int64_t)r * cols + c]] = (uint8_t)cnt;
}

// pointer jumping: rank = hops to the outlet
__global__ void k_pj_init(const int32_t *__restrict__ ds, int32_t *__restrict__ nxt, int32_t *__restrict__ rnk,
                          int64_t n)
{
    int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    int d = ds[p];
    nxt[p] = d >= 0 ? d : (int32_t)p;
    rnk[p] = d >= 0 ? 1 : 0;
}
__global__ void k_pj_step(const int32_t *__restrict__ nxt, const int32_t *__restrict__ rnk,
                          int32_t *__restrict__ nxt2, int32_t *__restrict__ rnk2, int64_t n,
                          int *__restrict__ active)
{
    int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    int a = nxt[p];
    int b = nxt[a];
    rnk2[p] = rnk[p] + rnk[a];
    nxt2[p] = b;
    if (b != a) *active = 1;  // benign race: everybody writes 1
}
// after convergence every pixel must have reached a true outlet; a pixel on (or draining into) a
// cycle ends on a node that still has a downstream link
__global__ void k_pj_check(const int32_t *__restrict__ nxt, const int32_t *__restrict__ ds, int64_t n,
                           int *__restrict__ cyc)
{
    int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p < n && ds[nxt[p]] >= 0) *cyc = 1;
}
__global__ void k_max(const int32_t *__restrict__ v, int64_t n, int *__restrict__ out)
{
    int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int x = p < n ? v[p] : 0;
    for (int o = 16; o > 0; o >>= 1) x = max(x, __shfl_xor_sync(0xffffffffu, x, o));
    if ((threadIdx.x & 31) == 0) atomicMax(out, x);
}
__global__ void k_max_u8(const uint8_t *__restrict__ v, int64_t n, int *__restrict__ out)
{
    int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int x = p < n ? v[p] : 0;
    for (int o = 16; o > 0; o >>= 1) x = max(x, __shfl_xor_sync(0xffffffffu, x, o));
    if ((threadIdx.x & 31) == 0) atomicMax(out, x);
}
__global__ void k_level(const int32_t *__restrict__ rnk, int32_t *__restrict__ level, int32_t *__restrict__ iota,
                        int64_t n, int maxrank)
{
    int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    level[p] = maxrank - rnk[p];  // order = max(dist) - dist, dist = rank + 1
    iota[p] = (int32_t)p;
}
__global__ void k_level_start(const int32_t *__restrict__ sorted_lev, int32_t *__restrict__ ls, int64_t n, int nl)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int k = sorted_lev[i];
    if (i == 0 || sorted_lev[i - 1] != k) ls[k] = (int32_t)i;
    if (i == n - 1) ls[nl] = (int32_t)n;
}
struct HasUpstream {
    const uint8_t *nups;
    __host__ __device__ bool operator()(const int32_t &p) const { return nups[p] != 0; }
};
__global__ void k_top_level(const int32_t *__restrict__ top, int32_t *__restrict__ pix_of_pos,
                            int32_t *__restrict__ pos_of_pix, int32_t *__restrict__ cfirst, int s, int e,
                            int64_t n, int ls0_end)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    // level 0 (headwaters farthest from the outlets) has no children: cfirst = 0 over its span
    if (i < ls0_end) cfirst[i] = 0;
    if (i == 0) cfirst[n] = s;
    int64_t j = s + i;
    if (j < e) {
        int p = top[i];
        pix_of_pos[j] = p;
        pos_of_pix[p] = (int32_t)j;
    }
}

// children of pixel p, in slot order, are placed at positions first, first+1, ...
__device__ __forceinline__ void place_children(int p, int first, const uint8_t *__restrict__ dir2d,
                                               const int32_t *__restrict__ lp,
                                               const int32_t *__restrict__ cell_of_pix, int32_t *pix_of_pos,
                                               int32_t *pos_of_pix, int rows, int cols)
{
    int cell = cell_of_pix[p];
    int r = cell / cols, c = cell - r * cols;
    int k = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        int rr = r + c_nb_dr[j], cc = c + c_nb_dc[j];
        if (rr < 0 || cc < 0 || rr >= rows || cc >= cols) continue;
        int t = rr * cols + cc;
        if (dir2d[t] == c_nb_need[j]) {
            int child = lp[t];
            pix_of_pos[first + k] = child;
            pos_of_pix[child] = first + k;
            ++k;
        }
    }
}

// One thread block walks a run of consecutive small levels (parents at `lev`, children at lev-1).
constexpr int BFS_THREADS = 1024;
__global__ void __launch_bounds__(BFS_THREADS)
    k_bfs_small(int lev_hi, int lev_lo, const int32_t *__restrict__ ls, const uint8_t *__restrict__ nups,
                int32_t *pix_of_pos, int32_t *pos_of_pix, int32_t *__restrict__ cfirst,
                const uint8_t *__restrict__ dir2d, const int32_t *__restrict__ lp,
                const int32_t *__restrict__ cell_of_pix, int rows, int cols)
{
    typedef cub::BlockScan<int, BFS_THREADS> Scan;
    __shared__ typename Scan::TempStorage tmp;
    __shared__ int carry;
    for (int lev = lev_hi; lev >= lev_lo; --lev) {
        int s1 = ls[lev], e1 = ls[lev + 1], base = ls[lev - 1];
        if (threadIdx.x == 0) carry = 0;
        __syncthreads();
        for (int c = s1; c < e1; c += BFS_THREADS) {
            int i = c + threadIdx.x;
            int p = -1, cnt = 0;
            if (i < e1) {
                p = pix_of_pos[i];
                cnt = nups[p];
            }
            int ex, total;
            Scan(tmp).ExclusiveSum(cnt, ex, total);
            int first = base + carry + ex;
            if (i < e1) {
                cfirst[i] = first;
                if (cnt) place_children(p, first, dir2d, lp, cell_of_pix, pix_of_pos, pos_of_pix, rows, cols);
            }
            __syncthreads();
            if (threadIdx.x == 0) carry += total;
            __syncthreads();
        }
        __threadfence_block();
        __syncthreads();
    }
}
__global__ void k_bfs_count(const int32_t *__restrict__ pix_of_pos, const uint8_t *__restrict__ nups,
                            int32_t *__restrict__ cnt, int s1, int e1)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x + s1;
    if (i < e1) cnt[i - s1] = nups[pix_of_pos[i]];
}
__global__ void k_bfs_place(const int32_t *__restrict__ scan, const uint8_t *__restrict__ nups,
                            int32_t *pix_of_pos, int32_t *pos_of_pix, int32_t *__restrict__ cfirst, int s1, int e1,
                            int base, const uint8_t *__restrict__ dir2d, const int32_t *__restrict__ lp,
                            const int32_t *__restrict__ cell_of_pix, int rows, int cols)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x + s1;
    if (i >= e1) return;
    int p = pix_of_pos[i];
    int first = base + scan[i - s1];
    cfirst[i] = first;
    if (nups[p]) place_children(p, first, dir2d, lp, cell_of_pix, pix_of_pos, pos_of_pix, rows, cols);
}

// accuflux: acc[i] = x[i] + sum of acc over the upstream positions (PCRaster accuflux)
__global__ void k_accuflux_level(double *__restrict__ acc, const int32_t *__restrict__ cfirst, int lo, int hi)
{
    int i = lo + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= hi) return;
    double s = acc[i];
    for (int k = cfirst[i]; k < cfirst[i + 1]; ++k) s += acc[k];
    acc[i] = s;
}
__global__ void k_gather_f64(const double *__restrict__ src, double *__restrict__ dst, const int32_t *__restrict__ idx,
                             int64_t n)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[idx[i]];
}


// ---- LDD-cut partition (multi-GPU): everything in position order ----
// parent position of every position (-1: outlet)
__global__ void k_parent_pos(const int32_t *__restrict__ pix_of_pos, const int32_t *__restrict__ pos_of_pix,
                             const int32_t *__restrict__ ds, int32_t *__restrict__ ppos, int64_t n)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int d = ds[pix_of_pos[i]];
    ppos[i] = d >= 0 ? pos_of_pix[d] : -1;
}
// trunk = upstream area above the threshold; root = non-trunk pixel hanging off the trunk or draining to an outlet
__global__ void k_part_flags(const double *__restrict__ size, const int32_t *__restrict__ ppos, double thr,
                             uint8_t *__restrict__ trunk, int64_t n)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) trunk[i] = size[i] > thr;
}
// one level, outlets first: a non-trunk pixel carries the position of the root of its sub-tree
__global__ void k_part_label(const uint8_t *__restrict__ trunk, const int32_t *__restrict__ ppos, int32_t *__restrict__ label,
                             int lo, int hi)
{
    int i = lo + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= hi) return;
    if (trunk[i]) {
        label[i] = -1;
        return;
    }
    const int pp = ppos[i];
    label[i] = (pp < 0 || trunk[pp]) ? i : label[pp];
}
struct IsRoot {
    const int32_t *label;
    __host__ __device__ bool operator()(const int32_t &i) const { return label[i] == i; }
};
__global__ void k_iota(int32_t *__restrict__ v, int64_t n)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) v[i] = (int32_t)i;
}
__global__ void k_root_info(const int32_t *__restrict__ roots, const double *__restrict__ size,
                            const int32_t *__restrict__ pix_of_pos, double *__restrict__ rsize, int32_t *__restrict__ rpix,
                            int64_t nr)
{
    int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nr) return;
    rsize[k] = size[roots[k]];
    rpix[k] = pix_of_pos[roots[k]];
}
__global__ void k_scatter_root_owner(const int32_t *__restrict__ roots, const int32_t *__restrict__ rown,
                                     int32_t *__restrict__ owner, int64_t nr)
{
    int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < nr) owner[roots[k]] = rown[k];
}
__global__ void k_owner_from_label(const int32_t *__restrict__ label, int32_t *__restrict__ owner, int64_t n)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int l = label[i];
    if (l >= 0 && l != i) owner[i] = owner[l];   // roots already carry their owner; label[l] == l is never rewritten
}
// one level, headwaters first: a trunk pixel joins the rank of its largest tributary (first in slot order on ties)
__global__ void k_trunk_owner(const uint8_t *__restrict__ trunk, const double *__restrict__ size,
                              const int32_t *__restrict__ cfirst, int32_t *__restrict__ owner, int lo, int hi)
{
    int i = lo + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= hi || !trunk[i]) return;
    double best = -1.0;
    int own = 0;
    for (int k = cfirst[i]; k < cfirst[i + 1]; ++k)
        if (size[k] > best) {
            best = size[k];
            own = owner[k];
        }
    owner[i] = own;
}
__global__ void k_i32_to_pix(const int32_t *__restrict__ src, int32_t *__restrict__ dst, const int32_t *__restrict__ pos_of_pix,
                             int64_t n)
{
    int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p < n) dst[p] = src[pos_of_pix[p]];
}
// cut edges: links whose two ends have different owners (owner in PIXEL order)
__global__ void k_cut_edges(const int32_t *__restrict__ ds, const int32_t *__restrict__ owner, int64_t n, int64_t cap,
                            int32_t *__restrict__ eu, int32_t *__restrict__ ed, unsigned long long *__restrict__ count)
{
    int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const int d = ds[p];
    if (d < 0 || owner[p] == owner[d]) return;
    const unsigned long long k = atomicAdd(count, 1ull);
    if ((int64_t)k < cap) {
        eu[k] = (int32_t)p;
        ed[k] = d;
    }
}
// ---- restriction of a graph to a subset of its pixels (the layout keeps the global order and the global levels) ----
struct KeepAtPosSafe {   // index n (one past the end) counts as dropped
    const uint8_t *keep;
    const int32_t *pix_of_pos;
    int32_t n;
    __host__ __device__ int32_t operator()(const int32_t &i) const { return (i < n && keep[pix_of_pos[i]]) ? 1 : 0; }
};
__global__ void k_fill_f64(double *__restrict__ v, double x, int64_t n)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) v[i] = x;
}
struct U8ToInt {
    __host__ __device__ int32_t operator()(const uint8_t &m) const { return m ? 1 : 0; }
};
__global__ void k_restrict(const uint8_t *__restrict__ keep, const int32_t *__restrict__ pix_of_pos_g,
                           const int32_t *__restrict__ cfirst_g, const int32_t *__restrict__ lev_g,
                           const int32_t *__restrict__ scan_pos, const int32_t *__restrict__ scan_pix,
                           int32_t *__restrict__ pix_of_pos, int32_t *__restrict__ pos_of_pix, int32_t *__restrict__ cfirst,
                           int32_t *__restrict__ cend, int32_t *__restrict__ lev, int64_t n)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int p = pix_of_pos_g[i];
    if (!keep[p]) return;
    const int j = scan_pos[i], q = scan_pix[p];
    pix_of_pos[j] = q;
    pos_of_pix[q] = j;
    lev[j] = lev_g[i];
    cfirst[j] = scan_pos[cfirst_g[i]];       // kept positions before the first child
    cend[j] = scan_pos[cfirst_g[i + 1]];     // ... before the end of the children: dropped children vanish
}
__global__ void k_restrict_levels(const int32_t *__restrict__ ls_g, const int32_t *__restrict__ scan_pos, int32_t *__restrict__ ls,
                                  int nl)
{
    int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l <= nl) ls[l] = scan_pos[ls_g[l]];
}

// ---- export kernels (reference-shaped int64 / float64 arrays) ----
__global__ void k_export_i64(const int32_t *__restrict__ src, int64_t *__restrict__ dst, int64_t n)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[i];
}
__global__ void k_export_u8_i64(const uint8_t *__restrict__ src, int64_t *__restrict__ dst, int64_t n)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[i];
}
__global__ void k_export_f64(const int32_t *__restrict__ src, double *__restrict__ dst, int64_t n)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = (double)src[i];
}
__global__ void k_export_ups(const uint8_t *__restrict__ dir2d, const int32_t *__restrict__ lp,
                             const int32_t *__restrict__ cell_of_pix, int64_t *__restrict__ ups, int64_t n,
                             int K, int rows, int cols)
{
    int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    int cell = cell_of_pix[p];
    int r = cell / cols, c = cell - r * cols;
    int k = 0;
    for (int j = 0; j < 8; ++j) {
        int rr = r + c_nb_dr[j], cc = c + c_nb_dc[j];
        if (rr < 0 || cc < 0 || rr >= rows || cc >= cols) continue;
        int t = rr * cols + cc;
        if (dir2d[t] == c_nb_need[j]) ups[p * K + k++] = lp[t];
    }
    for (; k < K; ++k) ups[p * K + k] = -1;
}

}  // namespace

using lf::blocks_for;
using lf::DevBuf;

static int build_graph(const double *ldd_codes, const uint8_t *land_mask, int64_t rows, int64_t cols, lf_graph *g)
{
    cudaStream_t st = lf::stream();
    const int T = 256;
    int64_t ncell = rows * cols;
    g->rows = rows;
    g->cols = cols;

    // ---- 1. mask -> compressed indices
    DevBuf<uint8_t> d_mask;
    LF_CHECK(d_mask.alloc(ncell));
    LF_CUDA(cudaMemcpyAsync(d_mask.p, land_mask, ncell, cudaMemcpyDefault, st));
    LF_CHECK(g->land_points.alloc(ncell + 1));
    DevBuf<uint8_t> d_tmp;
    size_t tmp_bytes = 0;
    {
        cub::TransformInputIterator<int32_t, MaskToInt, const uint8_t *> it(d_mask.p, MaskToInt());
        LF_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, it, g->land_points.p, (int)ncell, st));
        LF_CHECK(d_tmp.alloc(tmp_bytes));
        LF_CUDA(cub::DeviceScan::ExclusiveSum(d_tmp.p, tmp_bytes, it, g->land_points.p, (int)ncell, st));
        lf::count_launch(2);
    }
    int32_t last_scan = 0;
    uint8_t last_mask = 0;
    LF_CUDA(cudaMemcpyAsync(&last_mask, d_mask.p + (ncell - 1), 1, cudaMemcpyDeviceToHost, st));
    LF_CUDA(cudaMemcpyAsync(&last_scan, g->land_points.p + (ncell - 1), sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    LF_CUDA(cudaStreamSynchronize(st));
    int64_t n = (int64_t)last_scan + (last_mask ? 1 : 0);
    if (n <= 0) {
        lf::set_error("land_mask has no active pixel");
        return LF_ERR_INVALID;
    }
    g->n = n;
    LF_CHECK(g->cell_of_pix.alloc(n));
    k_land_points<<<blocks_for(ncell, T), T, 0, st>>>(d_mask.p, g->land_points.p, g->cell_of_pix.p, ncell);
    LF_LAUNCH_CHECK();

    // ---- 2. decode
    DevBuf<double> d_codes;
    LF_CHECK(d_codes.alloc(n));
    LF_CUDA(cudaMemcpyAsync(d_codes.p, ldd_codes, n * sizeof(double), cudaMemcpyDefault, st));
    LF_CHECK(g->dir2d.alloc(ncell));
    LF_CUDA(cudaMemsetAsync(g->dir2d.p, 9, ncell, st));
    DevBuf<int> d_flag;
    LF_CHECK(d_flag.alloc(2));
    LF_CUDA(cudaMemsetAsync(d_flag.p, 0, 2 * sizeof(int), st));
    k_decode<<<blocks_for(n, T), T, 0, st>>>(d_codes.p, g->cell_of_pix.p, g->dir2d.p, n, d_flag.p);
    LF_LAUNCH_CHECK();
    int h_flag[2] = {0, 0};
    LF_CUDA(cudaMemcpyAsync(h_flag, d_flag.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    LF_CUDA(cudaStreamSynchronize(st));
    if (h_flag[0]) {
        lf::set_error("LDD codes must be integers in 0..9 (kinematicWave contract)");
        return LF_ERR_BAD_LDD;
    }
    d_codes.release();
    d_mask.release();

    // ---- 3./4. downstream pixel and upstream counts
    LF_CHECK(g->downstream.alloc(n));
    k_downstream<<<blocks_for(n, T), T, 0, st>>>(g->dir2d.p, g->land_points.p, g->cell_of_pix.p, g->downstream.p, n,
                                                 (int)rows, (int)cols);
    LF_LAUNCH_CHECK();
    LF_CHECK(g->nups_pix.alloc(n));
    LF_CUDA(cudaMemsetAsync(g->nups_pix.p, 0, n, st));
    {
        dim3 blk(TILE_C, TILE_R), grd((unsigned)((cols + TILE_C - 1) / TILE_C), (unsigned)((rows + TILE_R - 1) / TILE_R));
        k_count_upstream<<<grd, blk, 0, st>>>(g->dir2d.p, g->land_points.p, g->nups_pix.p, (int)rows, (int)cols);
        LF_LAUNCH_CHECK();
    }

    // ---- 5. hops to outlet by pointer jumping
    DevBuf<int32_t> nxt_a, nxt_b, rnk_a, rnk_b;
    LF_CHECK(nxt_a.alloc(n));
    LF_CHECK(nxt_b.alloc(n));
    LF_CHECK(rnk_a.alloc(n));
    LF_CHECK(rnk_b.alloc(n));
    k_pj_init<<<blocks_for(n, T), T, 0, st>>>(g->downstream.p, nxt_a.p, rnk_a.p, n);
    LF_LAUNCH_CHECK();
    int32_t *nx = nxt_a.p, *nx2 = nxt_b.p, *rk = rnk_a.p, *rk2 = rnk_b.p;
    int rounds = 0;
    for (;; ++rounds) {
        if (rounds > 33) {
            lf::set_error("the LDD contains a cycle (no outlet reachable)");
            return LF_ERR_LDD_CYCLE;
        }
        LF_CUDA(cudaMemsetAsync(d_flag.p + 1, 0, sizeof(int), st));
        k_pj_step<<<blocks_for(n, T), T, 0, st>>>(nx, rk, nx2, rk2, n, d_flag.p + 1);
        LF_LAUNCH_CHECK();
        LF_CUDA(cudaMemcpyAsync(h_flag + 1, d_flag.p + 1, sizeof(int), cudaMemcpyDeviceToHost, st));
        LF_CUDA(cudaStreamSynchronize(st));
        std::swap(nx, nx2);
        std::swap(rk, rk2);
        if (!h_flag[1]) break;
    }
    // rk = hops to outlet for every pixel.  Odd cycles never settle (caught above); even cycles
    // collapse onto self-loops that are not outlets:
    LF_CUDA(cudaMemsetAsync(d_flag.p, 0, sizeof(int), st));
    k_pj_check<<<blocks_for(n, T), T, 0, st>>>(nx, g->downstream.p, n, d_flag.p);
    LF_LAUNCH_CHECK();
    LF_CUDA(cudaMemcpyAsync(h_flag, d_flag.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    LF_CUDA(cudaStreamSynchronize(st));
    if (h_flag[0]) {
        lf::set_error("the LDD contains a cycle (no outlet reachable)");
        return LF_ERR_LDD_CYCLE;
    }
    LF_CUDA(cudaMemsetAsync(d_flag.p, 0, sizeof(int), st));
    k_max<<<blocks_for(n, T), T, 0, st>>>(rk, n, d_flag.p);
    LF_LAUNCH_CHECK();
    LF_CUDA(cudaMemcpyAsync(h_flag, d_flag.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    LF_CUDA(cudaStreamSynchronize(st));
    int maxrank = h_flag[0];
    int nl = maxrank + 1;
    g->n_orders = nl;

    // ---- 6. routing order + reference ordering (stable sort by order; ties keep pixel order)
    LF_CHECK(g->level_pix.alloc(n));
    int32_t *iota = nx2;  // reuse
    k_level<<<blocks_for(n, T), T, 0, st>>>(rk, g->level_pix.p, iota, n, maxrank);
    LF_LAUNCH_CHECK();
    LF_CHECK(g->pixels_ordered.alloc(n));
    LF_CHECK(g->lev_of_pos.alloc(n));
    {
        int bits = 1;
        while ((1ll << bits) < nl) ++bits;
        size_t sb = 0;
        LF_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, sb, g->level_pix.p, g->lev_of_pos.p, iota,
                                                g->pixels_ordered.p, (int)n, 0, bits, st));
        if (sb > tmp_bytes) {
            LF_CHECK(d_tmp.alloc(sb));
            tmp_bytes = sb;
        }
        LF_CUDA(cub::DeviceRadixSort::SortPairs(d_tmp.p, sb, g->level_pix.p, g->lev_of_pos.p, iota,
                                                g->pixels_ordered.p, (int)n, 0, bits, st));
        lf::count_launch(2 * ((bits + 7) / 8) + 1);
    }
    LF_CHECK(g->level_start.alloc(nl + 1));
    k_level_start<<<blocks_for(n, T), T, 0, st>>>(g->lev_of_pos.p, g->level_start.p, n, nl);
    LF_LAUNCH_CHECK();
    g->h_level_start.resize(nl + 1);
    LF_CUDA(cudaMemcpyAsync(g->h_level_start.data(), g->level_start.p, (nl + 1) * sizeof(int32_t),
                            cudaMemcpyDeviceToHost, st));
    LF_CUDA(cudaStreamSynchronize(st));
    const std::vector<int32_t> &ls = g->h_level_start;
    g->n_pits = ls[nl] - ls[nl - 1];
    nxt_a.release();
    nxt_b.release();
    rnk_a.release();
    rnk_b.release();

    // ---- 7. breadth-first layout from the outlets
    LF_CHECK(g->pix_of_pos.alloc(n));
    LF_CHECK(g->pos_of_pix.alloc(n));
    LF_CHECK(g->cfirst.alloc(n + 1));
    {
        // outlets with an upstream network first (pixel order), isolated pits (no link at all) last:
        // the fused channel kernel advances isolated pixels through all sub-steps in registers
        int64_t top_n = ls[nl] - ls[nl - 1];
        DevBuf<int32_t> top;
        LF_CHECK(top.alloc(top_n));
        HasUpstream sel{g->nups_pix.p};
        size_t sb = 0;
        int *d_nsel = d_flag.p;
        LF_CUDA(cub::DevicePartition::If(nullptr, sb, g->pixels_ordered.p + ls[nl - 1], top.p, d_nsel, (int)top_n, sel, st));
        if (sb > tmp_bytes) {
            LF_CHECK(d_tmp.alloc(sb));
            tmp_bytes = sb;
        }
        LF_CUDA(cub::DevicePartition::If(d_tmp.p, sb, g->pixels_ordered.p + ls[nl - 1], top.p, d_nsel, (int)top_n, sel, st));
        lf::count_launch(2);
        LF_CUDA(cudaMemcpyAsync(h_flag, d_nsel, sizeof(int), cudaMemcpyDeviceToHost, st));
        LF_CUDA(cudaStreamSynchronize(st));
        g->n_isolated = top_n - h_flag[0];
        int64_t span = std::max<int64_t>(std::max<int64_t>(top_n, ls[1]), 1);
        k_top_level<<<blocks_for(span, T), T, 0, st>>>(top.p, g->pix_of_pos.p, g->pos_of_pix.p, g->cfirst.p, ls[nl - 1],
                                                       ls[nl], n, ls[1]);
        LF_LAUNCH_CHECK();
        LF_CUDA(cudaStreamSynchronize(st));
    }
    const int SMALL = 8 * BFS_THREADS;
    DevBuf<int32_t> cnt, scn;
    int64_t big_cap = 0;
    int lev = nl - 1;
    while (lev >= 1) {
        int64_t span = ls[lev + 1] - ls[lev];
        if (span <= SMALL) {
            int lo = lev;
            while (lo - 1 >= 1 && (ls[lo] - ls[lo - 1]) <= SMALL) --lo;
            k_bfs_small<<<1, BFS_THREADS, 0, st>>>(lev, lo, g->level_start.p, g->nups_pix.p, g->pix_of_pos.p,
                                                   g->pos_of_pix.p, g->cfirst.p, g->dir2d.p, g->land_points.p,
                                                   g->cell_of_pix.p, (int)rows, (int)cols);
            LF_LAUNCH_CHECK();
            lev = lo - 1;
        } else {
            if (span > big_cap) {
                LF_CHECK(cnt.alloc(span));
                LF_CHECK(scn.alloc(span));
                big_cap = span;
                size_t sb = 0;
                LF_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, sb, cnt.p, scn.p, (int)span, st));
                if (sb > tmp_bytes) {
                    LF_CHECK(d_tmp.alloc(sb));
                    tmp_bytes = sb;
                }
            }
            k_bfs_count<<<blocks_for(span, T), T, 0, st>>>(g->pix_of_pos.p, g->nups_pix.p, cnt.p, ls[lev], ls[lev + 1]);
            LF_LAUNCH_CHECK();
            size_t sb = tmp_bytes;
            LF_CUDA(cub::DeviceScan::ExclusiveSum(d_tmp.p, sb, cnt.p, scn.p, (int)span, st));
            lf::count_launch(2);
            k_bfs_place<<<blocks_for(span, T), T, 0, st>>>(scn.p, g->nups_pix.p, g->pix_of_pos.p, g->pos_of_pix.p,
                                                           g->cfirst.p, ls[lev], ls[lev + 1], ls[lev - 1], g->dir2d.p,
                                                           g->land_points.p, g->cell_of_pix.p, (int)rows, (int)cols);
            LF_LAUNCH_CHECK();
            --lev;
        }
    }
    LF_CUDA(cudaStreamSynchronize(st));

    // max_upstream = number of non-empty upstream columns = max count (>= 1), kinematic_wave_parallel.py:89
    LF_CUDA(cudaMemsetAsync(d_flag.p, 0, sizeof(int), st));
    k_max_u8<<<blocks_for(n, T), T, 0, st>>>(g->nups_pix.p, n, d_flag.p);
    LF_LAUNCH_CHECK();
    LF_CUDA(cudaMemcpyAsync(h_flag, d_flag.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    LF_CUDA(cudaStreamSynchronize(st));
    g->max_ups = h_flag[0] > 1 ? h_flag[0] : 1;
    return LF_OK;
}

extern "C" {

int lf_ldd_build(const double *ldd_codes, const uint8_t *land_mask, int64_t rows, int64_t cols, lf_graph **out)
{
    if (!ldd_codes || !land_mask || !out || rows <= 0 || cols <= 0) {
        lf::set_error("lf_ldd_build: null pointer or empty raster");
        return LF_ERR_INVALID;
    }
    if (rows * cols >= (1ll << 31) - 1) {
        lf::set_error("lf_ldd_build: raster of %lld cells exceeds the int32 index range", (long long)(rows * cols));
        return LF_ERR_INVALID;
    }
    LF_CHECK(lf::ensure_device());
    lf_graph *g = new lf_graph();
    int rc = build_graph(ldd_codes, land_mask, rows, cols, g);
    if (rc != LF_OK) {
        delete g;
        *out = nullptr;
        return rc;
    }
    *out = g;
    return LF_OK;
}

int lf_graph_info(const lf_graph *g, int64_t *n_pixels, int64_t *n_orders, int64_t *max_upstream, int64_t *n_pits)
{
    if (!g) {
        lf::set_error("lf_graph_info: null graph");
        return LF_ERR_INVALID;
    }
    if (n_pixels) *n_pixels = g->n;
    if (n_orders) *n_orders = g->n_orders;
    if (max_upstream) *max_upstream = g->max_ups;
    if (n_pits) *n_pits = g->n_pits;
    return LF_OK;
}

int lf_graph_export(const lf_graph *g, int64_t *pixels_ordered, int64_t *order_start_stop, int64_t *upstream_lookup,
                    int64_t *num_upstream, double *downstream)
{
    if (!g) {
        lf::set_error("lf_graph_export: null graph");
        return LF_ERR_INVALID;
    }
    if (g->restricted) {
        lf::set_error("lf_graph_export: a restricted graph (lf_graph_restrict) has no reference-shaped arrays");
        return LF_ERR_STATE;
    }
    LF_CHECK(lf::ensure_device());
    cudaStream_t st = lf::stream();
    const int T = 256;
    int64_t n = g->n;
    if (order_start_stop) {
        for (int o = 0; o < g->n_orders; ++o) {
            order_start_stop[2 * o] = g->h_level_start[o];
            order_start_stop[2 * o + 1] = g->h_level_start[o + 1];
        }
    }
    DevBuf<int64_t> tmp;
    if (pixels_ordered || num_upstream || downstream) LF_CHECK(tmp.alloc(n));
    if (pixels_ordered) {
        k_export_i64<<<blocks_for(n, T), T, 0, st>>>(g->pixels_ordered.p, tmp.p, n);
        LF_LAUNCH_CHECK();
        LF_CUDA(cudaMemcpyAsync(pixels_ordered, tmp.p, n * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
        LF_CUDA(cudaStreamSynchronize(st));
    }
    if (num_upstream) {
        k_export_u8_i64<<<blocks_for(n, T), T, 0, st>>>(g->nups_pix.p, tmp.p, n);
        LF_LAUNCH_CHECK();
        LF_CUDA(cudaMemcpyAsync(num_upstream, tmp.p, n * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
        LF_CUDA(cudaStreamSynchronize(st));
    }
    if (downstream) {
        k_export_f64<<<blocks_for(n, T), T, 0, st>>>(g->downstream.p, (double *)tmp.p, n);
        LF_LAUNCH_CHECK();
        LF_CUDA(cudaMemcpyAsync(downstream, tmp.p, n * sizeof(double), cudaMemcpyDeviceToHost, st));
        LF_CUDA(cudaStreamSynchronize(st));
    }
    if (upstream_lookup) {
        int K = g->max_ups;
        DevBuf<int64_t> ups;
        LF_CHECK(ups.alloc(n * K));
        k_export_ups<<<blocks_for(n, T), T, 0, st>>>(g->dir2d.p, g->land_points.p, g->cell_of_pix.p, ups.p, n, K,
                                                     (int)g->rows, (int)g->cols);
        LF_LAUNCH_CHECK();
        LF_CUDA(cudaMemcpyAsync(upstream_lookup, ups.p, n * K * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
        LF_CUDA(cudaStreamSynchronize(st));
    }
    return LF_OK;
}

int lf_graph_layout(const lf_graph *g, int32_t *pixel_of_position, int32_t *level_start)
{
    if (!g) {
        lf::set_error("lf_graph_layout: null graph");
        return LF_ERR_INVALID;
    }
    LF_CHECK(lf::ensure_device());
    cudaStream_t st = lf::stream();
    if (pixel_of_position) {
        LF_CUDA(cudaMemcpyAsync(pixel_of_position, g->pix_of_pos.p, g->n * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
        LF_CUDA(cudaStreamSynchronize(st));
    }
    if (level_start) memcpy(level_start, g->h_level_start.data(), (g->n_orders + 1) * sizeof(int32_t));
    return LF_OK;
}

int lf_graph_accuflux(const lf_graph *g, const double *x, double *out)
{
    if (!g || !x || !out) {
        lf::set_error("lf_graph_accuflux: null pointer");
        return LF_ERR_INVALID;
    }
    LF_CHECK(lf::ensure_device());
    cudaStream_t st = lf::stream();
    int64_t n = g->n;
    DevBuf<double> a, b;
    LF_CHECK(a.alloc(n));
    LF_CHECK(b.alloc(n));
    LF_CUDA(cudaMemcpyAsync(a.p, x, n * sizeof(double), cudaMemcpyDefault, st));
    k_gather_f64<<<blocks_for(n, 256), 256, 0, st>>>(a.p, b.p, g->pix_of_pos.p, n);  // to position order
    LF_LAUNCH_CHECK();
    const std::vector<int32_t> &ls = g->h_level_start;
    for (int l = 1; l < g->n_orders; ++l) {
        int lo = ls[l], hi = ls[l + 1];
        if (hi <= lo) continue;
        k_accuflux_level<<<blocks_for(hi - lo, 256), 256, 0, st>>>(b.p, g->cfirst.p, lo, hi);
        LF_LAUNCH_CHECK();
    }
    k_gather_f64<<<blocks_for(n, 256), 256, 0, st>>>(b.p, a.p, g->pos_of_pix.p, n);  // back to compressed order
    LF_LAUNCH_CHECK();
    LF_CUDA(cudaMemcpyAsync(out, a.p, n * sizeof(double), cudaMemcpyDefault, st));
    LF_CUDA(cudaStreamSynchronize(st));
    return LF_OK;
}


int lf_graph_partition(const lf_graph *g, int32_t world, double subtree_fraction, int32_t *owner, int64_t *loads,
                       int64_t *n_trunk, int64_t *n_roots)
{
    if (!g || !owner || world < 1 || !(subtree_fraction > 0)) {
        lf::set_error("lf_graph_partition: bad arguments");
        return LF_ERR_INVALID;
    }
    if (g->restricted) {
        lf::set_error("lf_graph_partition: needs a graph built by lf_ldd_build");
        return LF_ERR_STATE;
    }
    LF_CHECK(lf::ensure_device());
    cudaStream_t st = lf::stream();
    const int T = 256;
    const int64_t n = g->n;
    const std::vector<int32_t> &ls = g->h_level_start;
    const int nl = g->n_orders;
    // upstream area in cells, position order
    DevBuf<double> size;
    LF_CHECK(size.alloc(n));
    k_fill_f64<<<blocks_for(n, T), T, 0, st>>>(size.p, 1.0, n);
    LF_LAUNCH_CHECK();
    for (int l = 1; l < nl; ++l) {
        int lo = ls[l], hi = ls[l + 1];
        if (hi <= lo) continue;
        k_accuflux_level<<<blocks_for(hi - lo, T), T, 0, st>>>(size.p, g->cfirst.p, lo, hi);
        LF_LAUNCH_CHECK();
    }
    DevBuf<int32_t> ppos, label, own;
    DevBuf<uint8_t> trunk;
    LF_CHECK(ppos.alloc(n));
    LF_CHECK(label.alloc(n));
    LF_CHECK(own.alloc(n));
    LF_CHECK(trunk.alloc(n));
    k_parent_pos<<<blocks_for(n, T), T, 0, st>>>(g->pix_of_pos.p, g->pos_of_pix.p, g->downstream.p, ppos.p, n);
    LF_LAUNCH_CHECK();
    const double thr = world > 1 ? std::max(1.0, subtree_fraction * (double)n / world) : 1e300;
    k_part_flags<<<blocks_for(n, T), T, 0, st>>>(size.p, ppos.p, thr, trunk.p, n);
    LF_LAUNCH_CHECK();
    for (int l = nl - 1; l >= 0; --l) {   // outlets first
        int lo = ls[l], hi = ls[l + 1];
        if (hi <= lo) continue;
        k_part_label<<<blocks_for(hi - lo, T), T, 0, st>>>(trunk.p, ppos.p, label.p, lo, hi);
        LF_LAUNCH_CHECK();
    }
    // the roots, in position order
    DevBuf<int32_t> iota, roots;
    DevBuf<int> d_cnt;
    LF_CHECK(iota.alloc(n));
    LF_CHECK(roots.alloc(n));
    LF_CHECK(d_cnt.alloc(1));
    k_iota<<<blocks_for(n, T), T, 0, st>>>(iota.p, n);
    LF_LAUNCH_CHECK();
    {
        IsRoot sel{label.p};
        size_t sb = 0;
        DevBuf<uint8_t> tmp;
        LF_CUDA(cub::DeviceSelect::If(nullptr, sb, iota.p, roots.p, d_cnt.p, (int)n, sel, st));
        LF_CHECK(tmp.alloc(sb));
        LF_CUDA(cub::DeviceSelect::If(tmp.p, sb, iota.p, roots.p, d_cnt.p, (int)n, sel, st));
        lf::count_launch(2);
        LF_CUDA(cudaStreamSynchronize(st));
    }
    int nr = 0;
    LF_CUDA(cudaMemcpy(&nr, d_cnt.p, sizeof(int), cudaMemcpyDeviceToHost));
    iota.release();
    DevBuf<double> rsize;
    DevBuf<int32_t> rpix, rown;
    LF_CHECK(rsize.alloc(nr));
    LF_CHECK(rpix.alloc(nr));
    LF_CHECK(rown.alloc(nr));
    if (nr > 0) {
        k_root_info<<<blocks_for(nr, T), T, 0, st>>>(roots.p, size.p, g->pix_of_pos.p, rsize.p, rpix.p, nr);
        LF_LAUNCH_CHECK();
    }
    std::vector<double> h_size(nr);
    std::vector<int32_t> h_pix(nr), h_own(nr);
    LF_CUDA(cudaMemcpyAsync(h_size.data(), rsize.p, nr * sizeof(double), cudaMemcpyDeviceToHost, st));
    LF_CUDA(cudaMemcpyAsync(h_pix.data(), rpix.p, nr * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    LF_CUDA(cudaStreamSynchronize(st));
    // longest-processing-time bin packing of the sub-trees over the ranks (deterministic: size descending, pixel ascending)
    {
        std::vector<int32_t> idx(nr);
        for (int k = 0; k < nr; ++k) idx[k] = k;
        std::sort(idx.begin(), idx.end(), [&](int32_t a, int32_t b) {
            if (h_size[a] != h_size[b]) return h_size[a] > h_size[b];
            return h_pix[a] < h_pix[b];
        });
        typedef std::pair<double, int> Load;   // (load, rank): smallest load first, then smallest rank
        std::priority_queue<Load, std::vector<Load>, std::greater<Load>> heap;
        for (int r = 0; r < world; ++r) heap.push(Load(0.0, r));
        for (int32_t k : idx) {
            Load l = heap.top();
            heap.pop();
            h_own[k] = l.second;
            heap.push(Load(l.first + h_size[k], l.second));
        }
    }
    LF_CUDA(cudaMemcpyAsync(rown.p, h_own.data(), nr * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    LF_CUDA(cudaMemsetAsync(own.p, 0, n * sizeof(int32_t), st));
    if (nr > 0) {
        k_scatter_root_owner<<<blocks_for(nr, T), T, 0, st>>>(roots.p, rown.p, own.p, nr);
        LF_LAUNCH_CHECK();
    }
    k_owner_from_label<<<blocks_for(n, T), T, 0, st>>>(label.p, own.p, n);
    LF_LAUNCH_CHECK();
    int64_t ntr = 0;
    if (world > 1) {
        for (int l = 1; l < nl; ++l) {   // headwaters first: children carry their owner already
            int lo = ls[l], hi = ls[l + 1];
            if (hi <= lo) continue;
            k_trunk_owner<<<blocks_for(hi - lo, T), T, 0, st>>>(trunk.p, size.p, g->cfirst.p, own.p, lo, hi);
            LF_LAUNCH_CHECK();
        }
    }
    // to pixel order, to the caller
    DevBuf<int32_t> own_pix;
    LF_CHECK(own_pix.alloc(n));
    k_i32_to_pix<<<blocks_for(n, T), T, 0, st>>>(own.p, own_pix.p, g->pos_of_pix.p, n);
    LF_LAUNCH_CHECK();
    LF_CUDA(cudaMemcpyAsync(owner, own_pix.p, n * sizeof(int32_t), cudaMemcpyDefault, st));
    LF_CUDA(cudaStreamSynchronize(st));
    if (loads || n_trunk) {
        std::vector<int32_t> h(n);
        std::vector<uint8_t> ht(n);
        LF_CUDA(cudaMemcpy(h.data(), own.p, n * sizeof(int32_t), cudaMemcpyDeviceToHost));
        LF_CUDA(cudaMemcpy(ht.data(), trunk.p, n, cudaMemcpyDeviceToHost));
        if (loads)
            for (int r = 0; r < world; ++r) loads[r] = 0;
        for (int64_t i = 0; i < n; ++i) {
            if (loads && h[i] >= 0 && h[i] < world) loads[h[i]] += 1;
            ntr += ht[i];
        }
    }
    if (n_trunk) *n_trunk = ntr;
    if (n_roots) *n_roots = nr;
    return LF_OK;
}

int lf_graph_cut_edges(const lf_graph *g, const int32_t *owner, int64_t cap, int32_t *edge_u, int32_t *edge_d, int64_t *n_edges)
{
    if (!g || !owner || !n_edges || cap < 0 || (cap > 0 && (!edge_u || !edge_d))) {
        lf::set_error("lf_graph_cut_edges: bad arguments");
        return LF_ERR_INVALID;
    }
    if (g->restricted) {
        lf::set_error("lf_graph_cut_edges: needs a graph built by lf_ldd_build");
        return LF_ERR_STATE;
    }
    LF_CHECK(lf::ensure_device());
    cudaStream_t st = lf::stream();
    const int64_t n = g->n;
    DevBuf<int32_t> d_owner, eu, ed;
    DevBuf<unsigned long long> cnt;
    LF_CHECK(d_owner.alloc(n));
    LF_CHECK(eu.alloc(std::max<int64_t>(cap, 1)));
    LF_CHECK(ed.alloc(std::max<int64_t>(cap, 1)));
    LF_CHECK(cnt.alloc(1));
    LF_CUDA(cudaMemcpyAsync(d_owner.p, owner, n * sizeof(int32_t), cudaMemcpyDefault, st));
    LF_CUDA(cudaMemsetAsync(cnt.p, 0, sizeof(unsigned long long), st));
    k_cut_edges<<<blocks_for(n, 256), 256, 0, st>>>(g->downstream.p, d_owner.p, n, cap, eu.p, ed.p, cnt.p);
    LF_LAUNCH_CHECK();
    unsigned long long h = 0;
    LF_CUDA(cudaMemcpyAsync(&h, cnt.p, sizeof(h), cudaMemcpyDeviceToHost, st));
    LF_CUDA(cudaStreamSynchronize(st));
    *n_edges = (int64_t)h;
    const int64_t m = std::min<int64_t>((int64_t)h, cap);
    if (m > 0) {
        std::vector<int32_t> hu(m), hd(m);
        LF_CUDA(cudaMemcpy(hu.data(), eu.p, m * sizeof(int32_t), cudaMemcpyDeviceToHost));
        LF_CUDA(cudaMemcpy(hd.data(), ed.p, m * sizeof(int32_t), cudaMemcpyDeviceToHost));
        std::vector<int64_t> idx(m);
        for (int64_t k = 0; k < m; ++k) idx[k] = k;
        std::sort(idx.begin(), idx.end(), [&](int64_t a, int64_t b) { return hu[a] < hu[b]; });   // a pixel has one downstream link
        for (int64_t k = 0; k < m; ++k) {
            edge_u[k] = hu[idx[k]];
            edge_d[k] = hd[idx[k]];
        }
    }
    return LF_OK;
}

int lf_graph_restrict(const lf_graph *g, const uint8_t *keep, lf_graph **out)
{
    if (!g || !keep || !out) {
        lf::set_error("lf_graph_restrict: null pointer");
        return LF_ERR_INVALID;
    }
    *out = nullptr;
    if (g->restricted) {
        lf::set_error("lf_graph_restrict: needs a graph built by lf_ldd_build");
        return LF_ERR_STATE;
    }
    LF_CHECK(lf::ensure_device());
    cudaStream_t st = lf::stream();
    const int T = 256;
    const int64_t n = g->n;
    DevBuf<uint8_t> d_keep, tmp;
    DevBuf<int32_t> scan_pos, scan_pix;
    LF_CHECK(d_keep.alloc(n + 1));
    LF_CHECK(scan_pos.alloc(n + 1));
    LF_CHECK(scan_pix.alloc(n + 1));
    LF_CUDA(cudaMemcpyAsync(d_keep.p, keep, n, cudaMemcpyDefault, st));
    LF_CUDA(cudaMemsetAsync(d_keep.p + n, 0, 1, st));
    {
        // kept positions before every position 0..n (position order) and kept pixels before every pixel (pixel order)
        cub::CountingInputIterator<int32_t> cnt_it(0);
        KeepAtPosSafe op{d_keep.p, g->pix_of_pos.p, (int32_t)n};
        cub::TransformInputIterator<int32_t, KeepAtPosSafe, cub::CountingInputIterator<int32_t>> it(cnt_it, op);
        cub::TransformInputIterator<int32_t, U8ToInt, const uint8_t *> it2(d_keep.p, U8ToInt());
        size_t sb = 0, sb2 = 0;
        LF_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, sb, it, scan_pos.p, (int)(n + 1), st));
        LF_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, sb2, it2, scan_pix.p, (int)(n + 1), st));
        LF_CHECK(tmp.alloc(std::max(sb, sb2)));
        LF_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, sb, it, scan_pos.p, (int)(n + 1), st));
        LF_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, sb2, it2, scan_pix.p, (int)(n + 1), st));
        lf::count_launch(4);
    }
    int32_t nloc = 0, iso_before = 0;
    LF_CUDA(cudaMemcpyAsync(&nloc, scan_pos.p + n, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    LF_CUDA(cudaMemcpyAsync(&iso_before, scan_pos.p + (n - g->n_isolated), sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    LF_CUDA(cudaStreamSynchronize(st));
    if (nloc <= 0) {
        lf::set_error("lf_graph_restrict: the subset is empty");
        return LF_ERR_INVALID;
    }
    std::unique_ptr<lf_graph> q(new lf_graph());
    q->rows = g->rows;
    q->cols = g->cols;
    q->n = nloc;
    q->n_orders = g->n_orders;
    q->max_ups = g->max_ups;
    q->n_isolated = nloc - iso_before;
    q->restricted = true;
    LF_CHECK(q->pix_of_pos.alloc(nloc));
    LF_CHECK(q->pos_of_pix.alloc(nloc));
    LF_CHECK(q->cfirst.alloc(nloc + 1));
    LF_CHECK(q->cend.alloc(nloc));
    LF_CHECK(q->lev_of_pos.alloc(nloc));
    LF_CHECK(q->level_start.alloc(g->n_orders + 1));
    LF_CUDA(cudaMemsetAsync(q->cfirst.p + nloc, 0, sizeof(int32_t), st));
    k_restrict<<<blocks_for(n, T), T, 0, st>>>(d_keep.p, g->pix_of_pos.p, g->cfirst.p, g->lev_of_pos.p, scan_pos.p, scan_pix.p,
                                               q->pix_of_pos.p, q->pos_of_pix.p, q->cfirst.p, q->cend.p, q->lev_of_pos.p, n);
    LF_LAUNCH_CHECK();
    k_restrict_levels<<<blocks_for(g->n_orders + 1, T), T, 0, st>>>(g->level_start.p, scan_pos.p, q->level_start.p, g->n_orders);
    LF_LAUNCH_CHECK();
    q->h_level_start.resize(g->n_orders + 1);
    LF_CUDA(cudaMemcpyAsync(q->h_level_start.data(), q->level_start.p, (g->n_orders + 1) * sizeof(int32_t),
                            cudaMemcpyDeviceToHost, st));
    LF_CUDA(cudaStreamSynchronize(st));
    q->n_pits = q->h_level_start[g->n_orders] - q->h_level_start[g->n_orders - 1];
    *out = q.release();
    return LF_OK;
}

void lf_graph_destroy(lf_graph *g) { delete g; }

}  // extern "C"
