// Empirical CRPS over an ensemble on the device: trainmodel.crps_empirical (trainmodel.py:66-110),
//   CRPS* = E|pred - truth| - 1/2 E|pred - pred'|.
// The reference sorts the S members of every pixel (torch.sort over a [S, ...] tensor) and uses the O(S log S) identity
//   1/2 E|X - X'| = sum_k (x_(k+1) - x_(k)) k (S - k) / S^2 = sum_{i<j} |x_i - x_j| / S^2;
// for S ~ 100 the direct pair sum is cheaper on a GPU than a sort: one thread per pixel, its S members staged in shared
// memory (thread-private column, conflict-free), the i loop tiled by four so each shared-memory read feeds eight FADDs.
// Nothing but the [pixels] result is written: no sorted copy of the ensemble.
#include "../../include/probunet_b200.h"
#include "common.cuh"

namespace pu {

constexpr int CRPS_THREADS = 128;

__global__ void __launch_bounds__(CRPS_THREADS) crps_kernel(const float* __restrict__ pred, const float* __restrict__ truth,
                                                             float* __restrict__ out, int S, long long outer,
                                                             long long inner, long long member_stride,
                                                             long long outer_stride) {
    extern __shared__ float sv[];      // [S][CRPS_THREADS]
    const long long total = outer * inner;
    const long long e = (long long)blockIdx.x * CRPS_THREADS + threadIdx.x;
    const bool valid = e < total;
    const long long o = valid ? e / inner : 0, r = valid ? e % inner : 0;
    const float* p = pred + o * outer_stride + r;
    const float y = valid ? truth[e] : 0.f;
    float* col = sv + threadIdx.x;
    float mae = 0.f;
    for (int s = 0; s < S; ++s) {
        const float v = valid ? p[(long long)s * member_stride] : 0.f;
        col[s * CRPS_THREADS] = v;
        mae += fabsf(v - y);
    }
    // thread-private columns: no barrier needed
    float pairs = 0.f;
    int i = 0;
    for (; i + 4 <= S; i += 4) {
        const float a0 = col[i * CRPS_THREADS], a1 = col[(i + 1) * CRPS_THREADS], a2 = col[(i + 2) * CRPS_THREADS],
                    a3 = col[(i + 3) * CRPS_THREADS];
        float acc = fabsf(a0 - a1) + fabsf(a0 - a2) + fabsf(a0 - a3) + fabsf(a1 - a2) + fabsf(a1 - a3) + fabsf(a2 - a3);
        for (int j = i + 4; j < S; ++j) {
            const float b = col[j * CRPS_THREADS];
            acc += (fabsf(a0 - b) + fabsf(a1 - b)) + (fabsf(a2 - b) + fabsf(a3 - b));
        }
        pairs += acc;
    }
    for (; i < S; ++i) {
        const float a = col[i * CRPS_THREADS];
        for (int j = i + 1; j < S; ++j) pairs += fabsf(a - col[j * CRPS_THREADS]);
    }
    if (valid) out[e] = mae / (float)S - pairs / ((float)S * (float)S);
}

}  // namespace pu

extern "C" int pu_crps_empirical(const float* pred, const float* truth, float* out, int S, long long outer, long long inner,
                                 long long member_stride, long long outer_stride, void* stream) {
    using namespace pu;
    PU_REQUIRE(pred && truth && out && S >= 1 && outer > 0 && inner > 0, "pu_crps_empirical: bad arguments");
    const size_t smem = sizeof(float) * (size_t)S * CRPS_THREADS;
    PU_REQUIRE(smem <= 200 * 1024, "pu_crps_empirical: at most %d ensemble members (got %d)", 200 * 1024 / (4 * CRPS_THREADS), S);
    // per-device attribute and S varies from call to call: set it on every large call (a cheap driver call)
    if (smem > 48 * 1024) PU_CUDA(cudaFuncSetAttribute(crps_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long long blocks = cdivll(outer * inner, CRPS_THREADS);
    PU_REQUIRE(blocks < (1LL << 31), "pu_crps_empirical: too many elements");
    crps_kernel<<<(unsigned)blocks, CRPS_THREADS, smem, (cudaStream_t)stream>>>(pred, truth, out, S, outer, inner, member_stride,
                                                                               outer_stride);
    return check_launch("crps_empirical");
}
