// Multi-tensor AdamW: one launch updates every parameter of the model (replaces the per-tensor loop of
// torch.optim.AdamW, main.py:95, train_prob_unet_model.py:92).  HBM-bound: 16 B read + 12 B written per parameter.
// The host describes the work as a table of chunks (<= 65536 elements of one tensor each) in device memory; blocks
// walk the table round-robin, so 400+ tensors of very different sizes load-balance over the SMs without a launch each.
#include "../../include/probunet_b200.h"
#include "common.cuh"

namespace pu {

struct AdamWConsts {
    float decay;        // 1 - lr * weight_decay
    float b1, omb1;     // beta1, 1 - beta1 (each rounded from double, like torch's scalar arguments)
    float b2, omb2;
    float bc2_sqrt;     // sqrt(1 - beta2 ** step)
    float step_size;    // lr / (1 - beta1 ** step)
    float eps;
};

__device__ __forceinline__ void adamw_update(float& p, float g, float& m, float& v, const AdamWConsts& k) {
    // torch.optim.AdamW (decoupled weight decay, bias-corrected), same operation order as torch's single-tensor path:
    // p.mul_(1 - lr wd); m.lerp_(g, 1 - b1); v.mul_(b2).addcmul_(g, g, 1 - b2); p.addcdiv_(m, sqrt(v)/sqrt(bc2) + eps, -lr/bc1)
    p *= k.decay;
    m = m + (g - m) * k.omb1;
    v = v * k.b2 + k.omb2 * (g * g);
    const float denom = sqrtf(v) / k.bc2_sqrt + k.eps;
    p -= k.step_size * (m / denom);
}

__global__ void __launch_bounds__(256) adamw_multi_kernel(const PuAdamWChunk* __restrict__ chunks, int nchunks,
                                                          const AdamWConsts k) {
    for (int c = blockIdx.x; c < nchunks; c += gridDim.x) {
        const PuAdamWChunk ch = chunks[c];
        float* p = reinterpret_cast<float*>(ch.p);
        const float* g = reinterpret_cast<const float*>(ch.g);
        float* m = reinterpret_cast<float*>(ch.m);
        float* v = reinterpret_cast<float*>(ch.v);
        const int n = ch.n;
        const bool vec = (((ch.p | ch.g | ch.m | ch.v) & 15ull) == 0);
        const int n4 = vec ? (n >> 2) : 0;
        for (int i = threadIdx.x; i < n4; i += blockDim.x) {
            float4 pp = reinterpret_cast<float4*>(p)[i];
            const float4 gg = reinterpret_cast<const float4*>(g)[i];
            float4 mm = reinterpret_cast<float4*>(m)[i];
            float4 vv = reinterpret_cast<float4*>(v)[i];
            adamw_update(pp.x, gg.x, mm.x, vv.x, k);
            adamw_update(pp.y, gg.y, mm.y, vv.y, k);
            adamw_update(pp.z, gg.z, mm.z, vv.z, k);
            adamw_update(pp.w, gg.w, mm.w, vv.w, k);
            reinterpret_cast<float4*>(p)[i] = pp;
            reinterpret_cast<float4*>(m)[i] = mm;
            reinterpret_cast<float4*>(v)[i] = vv;
        }
        for (int i = 4 * n4 + threadIdx.x; i < n; i += blockDim.x) {
            float pi = p[i], mi = m[i], vi = v[i];
            adamw_update(pi, g[i], mi, vi, k);
            p[i] = pi;
            m[i] = mi;
            v[i] = vi;
        }
    }
}

}  // namespace pu

extern "C" int pu_adamw_multi(const PuAdamWChunk* chunks, int nchunks, double lr, double beta1, double beta2, double eps,
                              double weight_decay, int step, void* stream) {
    using namespace pu;
    PU_REQUIRE(chunks && nchunks > 0 && step >= 1, "pu_adamw_multi: bad arguments");
    PU_REQUIRE(beta1 >= 0.0 && beta1 < 1.0 && beta2 >= 0.0 && beta2 < 1.0, "pu_adamw_multi: betas must be in [0, 1)");
    // every derived constant is formed in double from the caller's doubles and rounded once, as torch does
    AdamWConsts k;
    k.decay = (float)(1.0 - lr * weight_decay);
    k.b1 = (float)beta1;
    k.omb1 = (float)(1.0 - beta1);
    k.b2 = (float)beta2;
    k.omb2 = (float)(1.0 - beta2);
    k.bc2_sqrt = (float)sqrt(1.0 - pow(beta2, (double)step));
    k.step_size = (float)(lr / (1.0 - pow(beta1, (double)step)));
    k.eps = (float)eps;
    int grid = nchunks < 148 * 8 ? nchunks : 148 * 8;
    adamw_multi_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(chunks, nchunks, k);
    return check_launch("adamw_multi");
}
