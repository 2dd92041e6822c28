// lf_core.cu -- device selection, stream, error text, stopwatch, launch counter.
#include <stdarg.h>
#include <string.h>

#include <atomic>

#include "lf_common.cuh"

namespace {
thread_local char t_error[1024] = "";
cudaStream_t g_stream = nullptr;
int g_device = -1;  // -1 = not initialised
int g_sms = 0;
cudaEvent_t g_ev0 = nullptr, g_ev1 = nullptr;
std::atomic<int64_t> g_launches{0}, g_api_launches{0};
}  // namespace

namespace lf {

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(t_error, sizeof(t_error), fmt, ap);
    va_end(ap);
}

cudaStream_t stream() { return g_stream; }
void count_launch(int64_t n, int64_t api)
{
    g_launches.fetch_add(n, std::memory_order_relaxed);
    g_api_launches.fetch_add(api < 0 ? n : api, std::memory_order_relaxed);
}
int sm_count() { return g_sms; }

static int init_device(int device)
{
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count <= 0) {
        set_error("no CUDA device visible (%s); liblisf_b200 has no CPU fallback",
                  e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
        return LF_ERR_NO_DEVICE;
    }
    if (device < 0 || device >= count) {
        set_error("device %d out of range (0..%d)", device, count - 1);
        return LF_ERR_INVALID;
    }
    cudaDeviceProp prop;
    LF_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        set_error("device %d (%s) is compute capability %d.%d; this library is built for sm_100a only", device,
                  prop.name, prop.major, prop.minor);
        return LF_ERR_NO_DEVICE;
    }
    LF_CUDA(cudaSetDevice(device));
    if (g_stream && g_device != device) {
        cudaStreamDestroy(g_stream);
        g_stream = nullptr;
    }
    if (!g_stream) LF_CUDA(cudaStreamCreateWithFlags(&g_stream, cudaStreamNonBlocking));
    if (!g_ev0) {
        LF_CUDA(cudaEventCreate(&g_ev0));
        LF_CUDA(cudaEventCreate(&g_ev1));
    }
    g_device = device;
    g_sms = prop.multiProcessorCount;
    return LF_OK;
}

int ensure_device()
{
    if (g_device >= 0) {
        // the calling thread may differ from the initialising one
        cudaSetDevice(g_device);
        return LF_OK;
    }
    return init_device(0);
}

}  // namespace lf

extern "C" {

const char *lf_last_error(void) { return t_error; }
int lf_version(void) { return 100; }

int lf_device_init(int device) { return lf::init_device(device); }

int lf_device_info(char *name, int *sm_count, int *cc_major, int *cc_minor, int64_t *hbm_bytes)
{
    LF_CHECK(lf::ensure_device());
    cudaDeviceProp prop;
    LF_CUDA(cudaGetDeviceProperties(&prop, g_device));
    if (name) {
        strncpy(name, prop.name, 127);
        name[127] = 0;
    }
    if (sm_count) *sm_count = prop.multiProcessorCount;
    if (cc_major) *cc_major = prop.major;
    if (cc_minor) *cc_minor = prop.minor;
    if (hbm_bytes) *hbm_bytes = (int64_t)prop.totalGlobalMem;
    return LF_OK;
}

int lf_synchronize(void)
{
    LF_CHECK(lf::ensure_device());
    LF_CUDA(cudaStreamSynchronize(g_stream));
    return LF_OK;
}

int lf_timer_start(void)
{
    LF_CHECK(lf::ensure_device());
    LF_CUDA(cudaEventRecord(g_ev0, g_stream));
    return LF_OK;
}

int lf_timer_stop(double *elapsed_ms)
{
    LF_CHECK(lf::ensure_device());
    LF_CUDA(cudaEventRecord(g_ev1, g_stream));
    LF_CUDA(cudaEventSynchronize(g_ev1));
    float ms = 0.f;
    LF_CUDA(cudaEventElapsedTime(&ms, g_ev0, g_ev1));
    if (elapsed_ms) *elapsed_ms = (double)ms;
    return LF_OK;
}

int64_t lf_launch_count(int reset)
{
    int64_t v = g_launches.load();
    if (reset) {
        g_launches.store(0);
        g_api_launches.store(0);
    }
    return v;
}

int64_t lf_host_launch_count(void) { return g_api_launches.load(); }

int lf_host_register(void *ptr, int64_t bytes)
{
    if (!ptr || bytes <= 0) {
        lf::set_error("lf_host_register: null pointer or empty range");
        return LF_ERR_INVALID;
    }
    LF_CHECK(lf::ensure_device());
    cudaError_t e = cudaHostRegister(ptr, (size_t)bytes, cudaHostRegisterDefault);
    if (e == cudaErrorHostMemoryAlreadyRegistered) {
        cudaGetLastError();
        return LF_OK;
    }
    LF_CUDA(e);
    return LF_OK;
}

int lf_host_alloc(int64_t bytes, void **out)
{
    if (!out || bytes <= 0) {
        lf::set_error("lf_host_alloc: null pointer or empty size");
        return LF_ERR_INVALID;
    }
    LF_CHECK(lf::ensure_device());
    *out = nullptr;
    LF_CUDA(cudaHostAlloc(out, (size_t)bytes, cudaHostAllocDefault));
    return LF_OK;
}

int lf_host_free(void *ptr)
{
    if (!ptr) return LF_OK;
    LF_CHECK(lf::ensure_device());
    LF_CUDA(cudaFreeHost(ptr));
    return LF_OK;
}

int lf_host_unregister(void *ptr)
{
    if (!ptr) return LF_OK;
    LF_CHECK(lf::ensure_device());
    cudaError_t e = cudaHostUnregister(ptr);
    if (e == cudaErrorHostMemoryNotRegistered) {
        cudaGetLastError();
        return LF_OK;
    }
    LF_CUDA(e);
    return LF_OK;
}

}  // extern "C"
