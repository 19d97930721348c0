// Error plumbing, version and launch accounting of the C ABI (include/probunet_b200.h).
#include <stdarg.h>
#include <stdlib.h>

#include <atomic>
#include <mutex>
#include <set>
#include <tuple>

#include "../../include/probunet_b200.h"
#include "common.cuh"
#include "conv_internal.h"

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
void note_fallback(const char* what, int dtype, int C0, int C1, int Cout, int H, int W) {
    if (dtype != PU_BF16) return;                       // fp32 mode runs on the CUDA-core kernels by design
    if (C0 + C1 < 16 || Cout < 16) return;              // 3-channel heads (Fcomb logits and their gradients): by design too
    static const bool quiet = getenv("PU_QUIET_FALLBACK") != nullptr;
    if (quiet) return;
    static std::mutex mu;
    static std::set<std::tuple<int, int, int, int, int, bool>> seen;
    std::lock_guard<std::mutex> lock(mu);
    if (!seen.insert(std::make_tuple(C0, C1, Cout, H, W, what[9] == '_')).second) return;
    fprintf(stderr,
            "probunet_b200: %s falls back to the CUDA-core kernel for C0=%d C1=%d Cout=%d at %dx%d (tcgen05 tiles need "
            "channel counts that are multiples of 64 and images of at least 8x16 pixels); expect it to be ~100x slower "
            "per FLOP [reported once per shape; PU_QUIET_FALLBACK=1 silences this]\n",
            what, C0, C1, Cout, H, W);
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
