// lf_selftest.cu -- on-device accuracy check of the hand-written math (lf_math.cuh, lf_kw_solve.cuh) against the
// CUDA math library (C ABI: lf_math_selftest).  Test infrastructure for tests/test_gpu_math.py; not on the hot path.
#include "lf_common.cuh"
#include "lf_kw_solve.cuh"
#include "lf_math.cuh"

namespace {

__device__ __forceinline__ uint64_t splitmix(uint64_t &s)
{
    uint64_t z = (s += 0x9e3779b97f4a7c15ull);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}
__device__ __forceinline__ double u01(uint64_t &s) { return (double)(splitmix(s) >> 11) * (1.0 / 9007199254740992.0); }
__device__ __forceinline__ void upd(unsigned long long *slot, double got, double ref, double scale)
{
    double e = fabs(got - ref) / (fabs(ref) * scale);
    if (!(e == e)) e = 1e300;  // NaN
    atomicMax(slot, (unsigned long long)__double_as_longlong(e));  // non-negative doubles order like their bit patterns
}

// slots: 0 div_nr, 1 sqrt_nr, 2 pw_tab (normalised by 1+|y log2 x|), 3 exp_neg_tab (normalised by 1+|x|),
//        4 root5, 5 root3, 6 lfm::pw (normalised by 1+|y log2 x|), 7 pw_tab<false> on the van Genuchten form
__global__ void k_selftest(int64_t n, uint64_t seed, unsigned long long *out)
{
    __shared__ lfm::MathTab tab;
    lfm::tab_to_shared(&tab, threadIdx.x, blockDim.x);
    __syncthreads();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t s = seed + 0x1234567ull * (uint64_t)i;
    {   // quotients over 80 binades
        const double a = ldexp(1.0 + u01(s), (int)(u01(s) * 80) - 40) * (u01(s) < 0.5 ? -1 : 1);
        const double d = ldexp(1.0 + u01(s), (int)(u01(s) * 80) - 40) * (u01(s) < 0.5 ? -1 : 1);
        upd(out + 0, lfm::div_nr(a, d), a / d, 1.0);
    }
    {
        const double x = ldexp(1.0 + u01(s), (int)(u01(s) * 120) - 60);
        upd(out + 1, lfm::sqrt_nr(x), sqrt(x), 1.0);
    }
    {
        double x = u01(s);
        const int mode = (int)(u01(s) * 3);
        if (mode == 1) x = ldexp(1.0 + u01(s), -1 - (int)(u01(s) * 60));
        if (mode == 2) x = 1.0 - ldexp(1.0 + u01(s), -2 - (int)(u01(s) * 50));
        const double y = u01(s) < 0.5 ? 0.04 + 0.6 * u01(s) : 1.0 + 24.0 * u01(s);
        const double ref = pow(x, y);
        if (ref > 1e-290) {
            const double sc = 1.0 + fabs(y * log2(x));
            upd(out + 2, lfm::pw_tab<true>(x, y, &tab), ref, sc);
            upd(out + 6, lfm::pw(x, y), ref, sc);
        }
        // van Genuchten: 1 - (1 - sat^(1/m))^m with m in (0.05, 0.5); compared on the conductivity factor t^2
        const double m = 0.05 + 0.45 * u01(s), invm = 1.0 / m, sat = u01(s);
        const double tr = 1.0 - pow(1.0 - pow(sat, invm), m);
        const double tg = 1.0 - lfm::pw_tab<false>(1.0 - lfm::pw_tab<false>(sat, invm, &tab), m, &tab);
        if (tr > 1e-6) upd(out + 7, tg, tr, 1.0);
    }
    {
        const double x = -700.0 * u01(s) * u01(s);
        upd(out + 3, lfm::exp_neg_tab(x, &tab), exp(x), 1.0 + fabs(x));
    }
    {
        const double q = ldexp(1.0 + u01(s), (int)(u01(s) * 190) - 95);  // inside (1e-30, 1e30)
        if (lfkw::in_float_range(q)) {
            upd(out + 4, lfkw::root5(q), pow(q, 0.2), 1.0);
            upd(out + 5, lfkw::root3(q), cbrt(q), 1.0);
        }
    }
}

}  // namespace

extern "C" int lf_math_selftest(int64_t n, uint64_t seed, double *max_err)
{
    LF_CHECK(lf::ensure_device());
    if (n <= 0 || !max_err) {
        lf::set_error("lf_math_selftest: n > 0 and an output of 8 doubles are required");
        return LF_ERR_INVALID;
    }
    lf::DevBuf<unsigned long long> out;
    LF_CHECK(out.alloc(8));
    cudaStream_t st = lf::stream();
    LF_CUDA(cudaMemsetAsync(out.p, 0, 8 * sizeof(unsigned long long), st));
    k_selftest<<<lf::blocks_for(n, 256), 256, 0, st>>>(n, seed, out.p);
    LF_LAUNCH_CHECK();
    unsigned long long h[8];
    LF_CUDA(cudaMemcpyAsync(h, out.p, sizeof(h), cudaMemcpyDeviceToHost, st));
    LF_CUDA(cudaStreamSynchronize(st));
    for (int k = 0; k < 8; ++k) memcpy(max_err + k, h + k, 8);
    return LF_OK;
}
