// lf_xchg.cu -- exchange regions for the LDD-cut decomposition: allocation, CUDA IPC mapping between the ranks of one
// node, and the two per-run flow-control kernels (see lf_xchg.cuh for the protocol).
#include <string.h>

#include <vector>

#include "lf_common.cuh"
#include "lf_xchg.cuh"

struct lf_xchg {
    int rank = 0, world = 1;
    int64_t bytes = 0;
    void *base = nullptr;                 // own region (cudaMalloc)
    std::vector<void *> peer;             // mapped base of every rank's region ([rank] = base)
    lf::DevBuf<unsigned long long *> d_peer_done;   // device table: address of done[rank] in every peer's header
    int64_t epoch = 0;                    // runs started
    bool in_run = false;
};

namespace {

__global__ void k_xchg_fill(unsigned long long *__restrict__ p, int64_t n_header, int64_t n_total)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_total) p[i] = i < n_header ? 0ull : lfx::EMPTY;
}
// before run k: every peer must have finished run k-2 (its slots of this parity are empty again)
__global__ void k_xchg_wait(const unsigned long long *__restrict__ done, int world, unsigned long long need,
                            unsigned long long *__restrict__ abort_flag)
{
    const int r = threadIdx.x;
    if (r >= world) return;
    const long long t0 = clock64();
    for (;;) {
        unsigned long long v;
        asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(done + r) : "memory");
        if (v >= need) return;
        if (*(volatile unsigned long long *)abort_flag != 0ull || clock64() - t0 > lfx::POLL_TIMEOUT_CYCLES) {
            *(volatile unsigned long long *)abort_flag = 1ull;
            return;
        }
        __nanosleep(256);
    }
}
// after run k: tell every peer (and myself) that k+1 runs are complete here
__global__ void k_xchg_post(unsigned long long *const *__restrict__ peer_done, int world, unsigned long long value)
{
    const int r = threadIdx.x;
    if (r >= world) return;
    __threadfence_system();
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(peer_done[r]), "l"(value) : "memory");
}

}  // namespace

namespace lf {
int xchg_view_base(lf_xchg *x, unsigned long long **abort_flag)
{
    *abort_flag = (unsigned long long *)x->base + 511;
    return LF_OK;
}
}  // namespace lf

extern "C" {

int lf_xchg_create(int32_t rank, int32_t world, int64_t import_doubles, lf_xchg **out)
{
    if (!out || world < 1 || rank < 0 || rank >= world || import_doubles < 0 || world > 256) {
        lf::set_error("lf_xchg_create: bad arguments");
        return LF_ERR_INVALID;
    }
    *out = nullptr;
    LF_CHECK(lf::ensure_device());
    lf_xchg *x = new lf_xchg();
    x->rank = rank;
    x->world = world;
    x->bytes = lfx::HEADER_BYTES + std::max<int64_t>(import_doubles, 1) * 8;
    cudaError_t e = cudaMalloc(&x->base, x->bytes);
    if (e != cudaSuccess) {
        lf::set_error("lf_xchg_create: cudaMalloc(%lld) -> %s", (long long)x->bytes, cudaGetErrorString(e));
        delete x;
        return LF_ERR_CUDA;
    }
    x->peer.assign(world, nullptr);
    x->peer[rank] = x->base;
    cudaStream_t st = lf::stream();
    const int64_t nw = x->bytes / 8;
    k_xchg_fill<<<lf::blocks_for(nw, 256), 256, 0, st>>>((unsigned long long *)x->base, lfx::HEADER_BYTES / 8, nw);
    lf::count_launch();
    e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) {
        lf::set_error("lf_xchg_create: %s", cudaGetErrorString(e));
        cudaFree(x->base);
        delete x;
        return LF_ERR_CUDA;
    }
    *out = x;
    return LF_OK;
}

int lf_xchg_ipc_handle(lf_xchg *x, void *handle64)
{
    if (!x || !handle64) {
        lf::set_error("lf_xchg_ipc_handle: null pointer");
        return LF_ERR_INVALID;
    }
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handles are 64 bytes");
    LF_CHECK(lf::ensure_device());
    cudaIpcMemHandle_t h;
    LF_CUDA(cudaIpcGetMemHandle(&h, x->base));
    memcpy(handle64, &h, 64);
    return LF_OK;
}

int lf_xchg_open_peer(lf_xchg *x, int32_t peer, const void *handle64)
{
    if (!x || !handle64 || peer < 0 || peer >= x->world) {
        lf::set_error("lf_xchg_open_peer: bad arguments");
        return LF_ERR_INVALID;
    }
    if (peer == x->rank) return LF_OK;
    LF_CHECK(lf::ensure_device());
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    void *p = nullptr;
    LF_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    x->peer[peer] = p;
    return LF_OK;
}

/* for single-process tests: two regions of the same process act as "peers" of each other */
int lf_xchg_set_peer_local(lf_xchg *x, int32_t peer, lf_xchg *other)
{
    if (!x || !other || peer < 0 || peer >= x->world) {
        lf::set_error("lf_xchg_set_peer_local: bad arguments");
        return LF_ERR_INVALID;
    }
    x->peer[peer] = other->base;
    return LF_OK;
}

int lf_xchg_peer_base(lf_xchg *x, int32_t peer, uint64_t *address)
{
    if (x && peer == -1) peer = x->rank;   /* -1: this rank's own region */
    if (!x || !address || peer < 0 || peer >= x->world || !x->peer[peer]) {
        lf::set_error("lf_xchg_peer_base: peer %d is not mapped", (int)peer);
        return LF_ERR_STATE;
    }
    *address = (uint64_t)(uintptr_t)x->peer[peer];
    return LF_OK;
}

int lf_xchg_begin(lf_xchg *x, int32_t *parity)
{
    if (!x) {
        lf::set_error("lf_xchg_begin: null pointer");
        return LF_ERR_INVALID;
    }
    if (x->in_run) {
        lf::set_error("lf_xchg_begin: the previous run was not closed with lf_xchg_end");
        return LF_ERR_STATE;
    }
    LF_CHECK(lf::ensure_device());
    cudaStream_t st = lf::stream();
    for (int r = 0; r < x->world; ++r)
        if (!x->peer[r]) {
            lf::set_error("lf_xchg_begin: the region of rank %d is not mapped (lf_xchg_open_peer)", r);
            return LF_ERR_STATE;
        }
    if (!x->d_peer_done.p) {
        std::vector<unsigned long long *> h(x->world);
        for (int r = 0; r < x->world; ++r) h[r] = (unsigned long long *)x->peer[r] + x->rank;
        LF_CHECK(x->d_peer_done.alloc(x->world));
        LF_CUDA(cudaMemcpyAsync(x->d_peer_done.p, h.data(), x->world * sizeof(void *), cudaMemcpyHostToDevice, st));
        LF_CUDA(cudaStreamSynchronize(st));
    }
    if (x->epoch >= 2) {
        k_xchg_wait<<<1, 256, 0, st>>>((const unsigned long long *)x->base, x->world, (unsigned long long)(x->epoch - 1),
                                       (unsigned long long *)x->base + 511);
        LF_LAUNCH_CHECK();
    }
    if (parity) *parity = (int32_t)(x->epoch & 1);
    x->in_run = true;
    return LF_OK;
}

int lf_xchg_end(lf_xchg *x)
{
    if (!x || !x->in_run) {
        lf::set_error("lf_xchg_end: no run in progress");
        return LF_ERR_STATE;
    }
    LF_CHECK(lf::ensure_device());
    x->epoch += 1;
    k_xchg_post<<<1, 256, 0, lf::stream()>>>(x->d_peer_done.p, x->world, (unsigned long long)x->epoch);
    LF_LAUNCH_CHECK();
    x->in_run = false;
    return LF_OK;
}

int lf_xchg_status(lf_xchg *x, int32_t *aborted, int64_t *epoch)
{
    if (!x) {
        lf::set_error("lf_xchg_status: null pointer");
        return LF_ERR_INVALID;
    }
    LF_CHECK(lf::ensure_device());
    unsigned long long h = 0;
    LF_CUDA(cudaStreamSynchronize(lf::stream()));
    LF_CUDA(cudaMemcpy(&h, (unsigned long long *)x->base + 511, sizeof(h), cudaMemcpyDeviceToHost));
    if (aborted) *aborted = h != 0;
    if (epoch) *epoch = x->epoch;
    return LF_OK;
}

void lf_xchg_destroy(lf_xchg *x)
{
    if (!x) return;
    for (int r = 0; r < x->world; ++r)
        if (r != x->rank && x->peer[r] && x->peer[r] != nullptr) {
            // regions set with lf_xchg_set_peer_local belong to another handle of this process: closing them as IPC
            // mappings fails harmlessly
            if (cudaIpcCloseMemHandle(x->peer[r]) != cudaSuccess) cudaGetLastError();
        }
    if (x->base) cudaFree(x->base);
    delete x;
}

}  // extern "C"
