// cf_runtime.cu -- error channel, device checks (C ABI: cf_abi_version, cf_last_error, cf_device_check).
#include <stdarg.h>
#include <atomic>
#include <stdio.h>

#include "cf_common.cuh"

namespace cf {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_status(cudaError_t err, const char *what)
{
    if (err == cudaSuccess) return CF_OK;
    set_error("%s: %s (%s)", what, cudaGetErrorName(err), cudaGetErrorString(err));
    return CF_ERR_LAUNCH;
}

int launch_status(const char *what) { return cuda_status(cudaGetLastError(), what); }

static std::atomic<long long> g_launches{0};
void count_launches(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

static int g_arch_ok[64];  // 0 unknown, 1 ok, -1 bad
static int g_sms[64];

int require_sm100()
{
    int dev = -1;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess || dev < 0) {
        (void)cudaGetLastError();
        set_error("no CUDA device available (%s): libcf_b200 has no CPU fallback", cudaGetErrorString(e));
        return CF_ERR_ARCH;
    }
    if (dev < 64 && g_arch_ok[dev] == 1) return CF_OK;
    int major = 0, minor = 0, sms = 0;
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (major != 10) {
        set_error("device %d is sm_%d%d; libcf_b200 is built for sm_100a only and has no fallback", dev, major, minor);
        if (dev < 64) g_arch_ok[dev] = -1;
        return CF_ERR_ARCH;
    }
    if (dev < 64) {
        g_arch_ok[dev] = 1;
        g_sms[dev] = sms;
    }
    return CF_OK;
}

int sm_count()
{
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (dev < 64 && g_sms[dev] > 0) return g_sms[dev];
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return sms;
}

}  // namespace cf

extern "C" {
int cf_abi_version(void) { return CF_ABI_VERSION; }
const char *cf_last_error(void) { return cf::g_err; }
int cf_device_check(void) { return cf::require_sm100(); }
long long cf_launch_count(void) { return cf::g_launches.load(std::memory_order_relaxed); }
}
