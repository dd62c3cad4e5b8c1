// common.cu -- error state, device query, cached twiddle tables
#include "common.cuh"
#include <stdarg.h>
#include <math.h>
#include <map>
#include <mutex>
#include <tuple>
#include <vector>

namespace shb {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char *what) {
    set_error("CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
    return SHB_ERR_CUDA;
}

int sm_count() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

// exp(-2 pi i * num / den) with octant reduction in long double, so that every
// table entry is the correctly rounded fp64 value (parity budget: 1e-12).
static void unit_root(long long num, long long den, double *re, double *im) {
    num %= den;
    if (num < 0) num += den;
    // angle = (pi/4) * t/den with t = 8*num; o = octant, r/den = position inside it
    long long t = 8 * num;
    long long o = t / den;
    long long r = t - o * den;
    const long double qpi = 0.785398163397448309615660845819875721L;
    long double a_dir = qpi * (long double)r / (long double)den;          // in [0, pi/4)
    long double a_ref = qpi * (long double)(den - r) / (long double)den;  // in (0, pi/4]
    long double c, s;
    switch (o) {
    case 0: c = cosl(a_dir);  s = sinl(a_dir);  break;
    case 1: c = sinl(a_ref);  s = cosl(a_ref);  break;
    case 2: c = -sinl(a_dir); s = cosl(a_dir);  break;
    case 3: c = -cosl(a_ref); s = sinl(a_ref);  break;
    case 4: c = -cosl(a_dir); s = -sinl(a_dir); break;
    case 5: c = -sinl(a_ref); s = -cosl(a_ref); break;
    case 6: c = sinl(a_dir);  s = -cosl(a_dir); break;
    default: c = cosl(a_ref); s = -sinl(a_ref); break;
    }
    *re = (double)c;
    *im = (double)(-s);   // exp(-i angle)
}

const cplx *twiddle_table(int kind, int n) {
    static std::mutex mu;
    static std::map<std::tuple<int, int, int>, cplx *> cache;
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lock(mu);
    auto key = std::make_tuple(dev, kind, n);
    auto it = cache.find(key);
    if (it != cache.end()) return it->second;
    int count = (kind == 0) ? n : n + 1;
    std::vector<double> host(2 * (size_t)count);
    for (int k = 0; k < count; ++k) {
        long long den = (kind == 0) ? n : (kind == 1 ? 2LL * n : 4LL * n);
        unit_root(k, den, &host[2 * k], &host[2 * k + 1]);
    }
    cplx *d = nullptr;
    if (cudaMalloc(&d, sizeof(cplx) * count) != cudaSuccess) return nullptr;
    // synchronous copy: tables are created once per (device, kind, n)
    if (cudaMemcpy(d, host.data(), sizeof(cplx) * count, cudaMemcpyHostToDevice) != cudaSuccess) {
        cudaFree(d);
        return nullptr;
    }
    cache[key] = d;
    return d;
}

}  // namespace shb

extern "C" {

const char *shb_last_error(void) { return shb::g_err; }

int shb_version(void) { return 100; }

int shb_device_ok(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
        cudaGetLastError();
        return 0;
    }
    int dev = 0, major = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    return major == 10 ? 1 : 0;
}

}  // extern "C"
