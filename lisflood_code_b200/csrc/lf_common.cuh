// lf_common.cuh -- shared plumbing for liblisf_b200.so (error reporting, stream, handles).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/lisflood_b200.h"

namespace lf {

void set_error(const char *fmt, ...);
cudaStream_t stream();
int ensure_device();  // LF_OK or LF_ERR_NO_DEVICE / LF_ERR_CUDA
void count_launch(int64_t n = 1, int64_t api = -1);   // kernels executed, launch API calls (default: one per kernel)
int sm_count();

#define LF_CUDA(expr)                                                                              \
    do {                                                                                           \
        cudaError_t _e = (expr);                                                                   \
        if (_e != cudaSuccess) {                                                                   \
            lf::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e));   \
            return LF_ERR_CUDA;                                                                    \
        }                                                                                          \
    } while (0)

#define LF_CHECK(expr)                 \
    do {                               \
        int _rc = (expr);              \
        if (_rc != LF_OK) return _rc;  \
    } while (0)

#define LF_LAUNCH_CHECK()                                                                          \
    do {                                                                                           \
        lf::count_launch();                                                                        \
        cudaError_t _e = cudaGetLastError();                                                       \
        if (_e != cudaSuccess) {                                                                   \
            lf::set_error("%s:%d: kernel launch -> %s", __FILE__, __LINE__, cudaGetErrorString(_e)); \
            return LF_ERR_CUDA;                                                                    \
        }                                                                                          \
    } while (0)

// Owning device buffer (freed with the handle).
template <typename T>
struct DevBuf {
    T *p = nullptr;
    size_t n = 0;
    DevBuf() {}
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    ~DevBuf() { release(); }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        n = 0;
    }
    int alloc(size_t count) {
        release();
        if (count == 0) count = 1;
        cudaError_t e = cudaMalloc((void **)&p, count * sizeof(T));
        if (e != cudaSuccess) {
            p = nullptr;
            set_error("cudaMalloc(%zu bytes) -> %s", count * sizeof(T), cudaGetErrorString(e));
            return LF_ERR_CUDA;
        }
        n = count;
        return LF_OK;
    }
};

// true when p is device (or managed) memory; host pointers (pageable or pinned) give false
inline bool is_device_ptr(const void *p)
{
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}

inline unsigned blocks_for(int64_t n, int threads) { return (unsigned)((n + threads - 1) / threads); }

// ---- CUDA-graph replay of launch-bound loops (the level sweep of the overland routers, the diagonals of the channel
// wavefront: hundreds of small dependent launches per step).  A loop is captured once per distinct argument set (the
// kernel arguments are part of the key: buffers that alternate between steps give two or four variants) and replayed
// with one cudaGraphLaunch afterwards.
struct GraphCache {
    struct Entry {
        std::vector<uint8_t> key;
        cudaGraphExec_t exec = nullptr;
        int64_t kernels = 0;
        uint64_t last_use = 0;
    };
    std::vector<Entry> entries;
    uint64_t tick = 0;
    void clear()   // after a change of how the work is cut into launches (the keys do not see it)
    {
        for (Entry &e : entries)
            if (e.exec) cudaGraphExecDestroy(e.exec);
        entries.clear();
    }
    ~GraphCache() { clear(); }
};
template <class F>
inline int run_captured(GraphCache &gc, const void *key, size_t keylen, cudaStream_t s, F &&enqueue)
{
    gc.tick += 1;
    for (GraphCache::Entry &e : gc.entries)
        if (e.key.size() == keylen && memcmp(e.key.data(), key, keylen) == 0) {
            e.last_use = gc.tick;
            LF_CUDA(cudaGraphLaunch(e.exec, s));
            lf::count_launch(e.kernels, 1);   // the kernels of the graph run; the host made one call
            return LF_OK;
        }
    const int64_t before = lf_launch_count(0);
    LF_CUDA(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
    const int rc = enqueue();
    cudaGraph_t graph = nullptr;
    cudaError_t ce = cudaStreamEndCapture(s, &graph);
    if (rc != LF_OK || ce != cudaSuccess || !graph) {
        if (graph) cudaGraphDestroy(graph);
        if (rc != LF_OK) return rc;
        lf::set_error("CUDA graph capture failed: %s", cudaGetErrorString(ce));
        cudaGetLastError();
        return LF_ERR_CUDA;
    }
    GraphCache::Entry e;
    e.key.assign((const uint8_t *)key, (const uint8_t *)key + keylen);
    e.kernels = lf_launch_count(0) - before;
    e.last_use = gc.tick;
    ce = cudaGraphInstantiate(&e.exec, graph, 0);
    cudaGraphDestroy(graph);
    if (ce != cudaSuccess) {
        lf::set_error("cudaGraphInstantiate: %s", cudaGetErrorString(ce));
        return LF_ERR_CUDA;
    }
    if (gc.entries.size() >= 4) {   // drop the least recently used variant
        size_t old = 0;
        for (size_t k = 1; k < gc.entries.size(); ++k)
            if (gc.entries[k].last_use < gc.entries[old].last_use) old = k;
        cudaGraphExecDestroy(gc.entries[old].exec);
        gc.entries.erase(gc.entries.begin() + old);
    }
    LF_CUDA(cudaGraphLaunch(e.exec, s));
    gc.entries.push_back(e);
    return LF_OK;
}

}  // namespace lf

// ------------------------------------------------------------------------------------------------
// handles
// ------------------------------------------------------------------------------------------------
struct lf_graph {
    int64_t rows = 0, cols = 0, n = 0;
    int32_t n_orders = 0, max_ups = 1;
    int64_t n_pits = 0, n_isolated = 0;  // isolated = pits without any upstream pixel (stored last)
    // raster-order members (kept for export and for building routers)
    lf::DevBuf<uint8_t> dir2d;        // [rows*cols] 0..7 direction, 8 pit, 9 off-mask
    lf::DevBuf<int32_t> land_points;  // [rows*cols] compressed index or -1
    lf::DevBuf<int32_t> cell_of_pix;  // [n] linear cell index
    lf::DevBuf<int32_t> downstream;   // [n] pixel index or -1
    lf::DevBuf<uint8_t> nups_pix;     // [n]
    lf::DevBuf<int32_t> level_pix;    // [n] routing order of the pixel
    lf::DevBuf<int32_t> pixels_ordered;  // [n] reference order: sorted by (order, pixel)
    // internal breadth-first layout ("positions")
    lf::DevBuf<int32_t> pix_of_pos;   // [n]
    lf::DevBuf<int32_t> pos_of_pix;   // [n]
    lf::DevBuf<int32_t> cfirst;       // [n+1] children of position i are positions cfirst[i]..cfirst[i+1]-1
    lf::DevBuf<int32_t> cend;         // restricted graphs only (lf_graph_restrict): children of i are cfirst[i]..cend[i]-1
    bool restricted = false;          // no raster members: the reference-shaped exports are not available
    lf::DevBuf<int32_t> lev_of_pos;   // [n]
    lf::DevBuf<int32_t> level_start;  // [n_orders+1] (device copy)
    std::vector<int32_t> h_level_start;  // host copy
};
