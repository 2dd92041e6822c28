// lf_xchg.cuh -- in-kernel exchange of boundary discharges between the GPUs of one node (LDD-cut decomposition).
//
// Where the drainage graph is cut between two ranks (link u -> d, u on rank A, d on rank B) rank B keeps a GHOST
// of u: a work item of the routing wavefront that is not solved but receives u's new discharge of every routing
// step from rank A.  The transport is part of the routing kernels themselves:
//   * every rank owns one exchange REGION in its HBM, mapped into the other ranks' address space with CUDA IPC
//     (NVLink peer access): a header of per-peer counters followed by the import slots, float64;
//   * the thread of rank A that solves (u, step s) also stores the value straight into B's slot with a system-scope
//     store (8 bytes over NVLink: the value is its own "ready" flag);
//   * the ghost item (u, s) on rank B polls its slot until it no longer holds the EMPTY pattern, takes the value and
//     puts EMPTY back (it is the only reader, and reads once).
// Every rank indexes its wavefront diagonals by the GLOBAL routing level (lf_graph_restrict keeps it), so the
// consumer of a value sits exactly one diagonal after its producer on every rank: a rank waiting in diagonal d waits
// only for work of diagonals < d elsewhere -- no deadlock, for any partition (edges may go owner -> owner in any
// direction, no hub).  Slots are double-buffered by the parity of the RUN (one lf_router_run / one model step); a
// rank starts run k only after every peer has posted the completion of run k-2 (lf::xchg_begin / xchg_end: two tiny
// kernels per run), so a fast rank runs at most one run ahead of a slow one and never overwrites an unread slot.
// A poll gives up after ~10 s (peer died): the value becomes NaN and the region's abort flag is raised.
#pragma once
#include <stdint.h>

struct lf_xchg;

namespace lfx {

constexpr unsigned long long EMPTY = 0x7ff8dead0b200000ull;   // a NaN payload no computation produces
constexpr int HEADER_BYTES = 4096;                              // counters: done[world] (u64), abort flag at [511]
constexpr long long POLL_TIMEOUT_CYCLES = 20000000000ll;        // ~10 s at 2 GHz

// per-router view of the exchange, passed by value to the routing kernels
// An import block is laid out [run parity (2)][edge][section][step (cap)]; `nsec` sections share an edge (main channel +
// floodplain; the three overland routers).
struct View {
    const int32_t *xslot;        // per position: -1 plain; k >= 0 export edge k; <= -2 ghost of import edge -2-k; INERT: ghost without a link in this graph
    double *const *exp_dst;      // [n_export] address of (parity 0, edge, section 0, step 0) in the CONSUMER's region
    const long long *exp_stride; // [n_export] parity stride (doubles) of that consumer's import block
    double *imp;                 // own import block (parity 0)
    long long imp_parity_stride; // doubles
    int32_t cap;                 // steps per (edge, section)
    int32_t nsec;                // sections per edge
    int32_t parity;              // run parity
    unsigned long long *abort_flag;
};
constexpr int32_t INERT = -2147483647 - 1;

__device__ __forceinline__ double *export_slot(const View &X, int k, int sec, int step)
{
    return X.exp_dst[k] + X.parity * X.exp_stride[k] + ((long long)sec * X.cap + step);
}
__device__ __forceinline__ double *import_slot(const View &X, int k, int sec, int step)
{
    return X.imp + X.parity * X.imp_parity_stride + (((long long)k * X.nsec + sec) * X.cap + step);
}

__device__ __forceinline__ void push(double *dst, double v)
{
    asm volatile("st.relaxed.sys.global.f64 [%0], %1;" ::"l"(dst), "d"(v) : "memory");
}
__device__ __forceinline__ unsigned long long peek(const double *src)
{
    unsigned long long u;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(u) : "l"(src) : "memory");
    return u;
}
// value of an import slot: waits for the producer, then empties the slot for the run after next
__device__ __forceinline__ double take(double *slot, unsigned long long *abort_flag)
{
    unsigned long long u = peek(slot);
    if (u == EMPTY) {
        const long long t0 = clock64();
        for (;;) {
            u = peek(slot);
            if (u != EMPTY) break;
            if (*(volatile unsigned long long *)abort_flag != 0ull || clock64() - t0 > POLL_TIMEOUT_CYCLES) {
                *(volatile unsigned long long *)abort_flag = 1ull;
                return __longlong_as_double(0x7ff8000000000000ll);
            }
            __nanosleep(64);
        }
    }
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(slot), "l"(EMPTY) : "memory");
    return __longlong_as_double((long long)u);
}

}  // namespace lfx

// host side (lf_xchg.cu)
namespace lf {
int xchg_view_base(lf_xchg *x, unsigned long long **abort_flag);
}
