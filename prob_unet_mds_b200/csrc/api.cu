// Error plumbing, version and launch accounting of the C ABI (include/probunet_b200.h).
#include <stdarg.h>

#include <atomic>

#include "../../include/probunet_b200.h"
#include "common.cuh"

namespace pu {
static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int check_launch(const char* what) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaPeekAtLastError();
    if (e != cudaSuccess) {
        cudaGetLastError();
        set_error("%s: kernel launch failed: %s", what, cudaGetErrorString(e));
        return PU_ERR_CUDA;
    }
    return PU_OK;
}
}  // namespace pu

extern "C" {
const char* pu_last_error(void) { return pu::g_err; }
int pu_version(void) { return 100; }
long long pu_launch_count(int reset) {
    long long v = pu::g_launches.load();
    if (reset) pu::g_launches.store(0);
    return v;
}
int pu_zero(void* dst, long long bytes, void* stream) {
    PU_REQUIRE(dst != nullptr && bytes >= 0, "pu_zero: bad arguments");
    PU_CUDA(cudaMemsetAsync(dst, 0, (size_t)bytes, (cudaStream_t)stream));
    return PU_OK;
}
int pu_copy(void* dst, const void* src, long long bytes, void* stream) {
    PU_REQUIRE(dst != nullptr && src != nullptr && bytes >= 0, "pu_copy: bad arguments");
    PU_CUDA(cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return PU_OK;
}
int pu_device_supports_tc(void) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    int major = 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return 0;
    return major == 10 ? 1 : 0;
}
}
