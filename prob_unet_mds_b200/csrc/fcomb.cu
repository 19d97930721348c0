// Fused Fcomb forward: tile z over the image, concat with the U-Net features, 1x1 -> ReLU -> 1x1 -> ReLU -> 1x1.
// Reference: prob_unet.py:100-121.  The tiled z and the concat never exist: layer 0 is split algebraically,
//   W0 . [feat ; z] + b0 = W0f . feat + (W0z . z + b0),
// the second term is a per-(sample, member) bias and, for ensembles, W0f . feat is computed once per pixel and
// reused by all S members (encode once, re-run only Fcomb per sample).
// fp32 CUDA-core math; one thread per pixel, weights broadcast from shared memory.  (bf16 ensembles with S >= 4 go
// to the tensor-core kernel in fcomb_tc.cu.)
#include "../../include/probunet_b200.h"
#include "common.cuh"

namespace pu {

constexpr int FC = 64;   // unet_output_channels == num_filters[0]

// fcomb_tc.cu: tcgen05 ensemble decode (bf16, S >= 4 members per input)
bool fcomb_members_tc_applicable(const PuFcombArgs* a);
int fcomb_members_tc_launch(const PuFcombArgs* a, cudaStream_t st);

template <typename T>
__global__ void __launch_bounds__(128) fcomb_fwd_kernel(PuFcombArgs a) {
    __shared__ __align__(16) float w0t[FC][FC];   // [in][out]
    __shared__ __align__(16) float w1[FC][FC];    // [out][in]
    __shared__ float w2[3][FC];
    __shared__ float b1s[FC], b2s[3];
    __shared__ __align__(16) float zb[FC];
    const int tid = threadIdx.x;
    const int n = blockIdx.y;
    const int p = blockIdx.x * 128 + tid;
    const bool valid = p < a.HW;
    const int K0 = FC + a.L;
    for (int i = tid; i < FC * FC; i += 128) {
        const int o = i / FC, in = i % FC;
        w0t[in][o] = a.w0[(long long)o * K0 + in];
        w1[o][in] = a.w1[i];
    }
    for (int i = tid; i < a.num_classes * FC; i += 128) w2[i / FC][i % FC] = a.w2[i];
    if (tid < FC) b1s[tid] = a.b1[tid];
    if (tid < a.num_classes) b2s[tid] = a.b2[tid];
    __syncthreads();

    float pre[FC];
#pragma unroll
    for (int o = 0; o < FC; ++o) pre[o] = 0.f;
    if (valid) {
        const T* fp = reinterpret_cast<const T*>(a.feat) + ((long long)n * a.HW + p) * FC;
#pragma unroll
        for (int i0 = 0; i0 < FC; i0 += 8) {
            float f8[8];
            ld8(fp + i0, f8);
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const float fi = f8[e];
                const float4* wr = reinterpret_cast<const float4*>(&w0t[i0 + e][0]);
#pragma unroll
                for (int o4 = 0; o4 < FC / 4; ++o4) {
                    const float4 w = wr[o4];
                    pre[o4 * 4 + 0] = fmaf(w.x, fi, pre[o4 * 4 + 0]);
                    pre[o4 * 4 + 1] = fmaf(w.y, fi, pre[o4 * 4 + 1]);
                    pre[o4 * 4 + 2] = fmaf(w.z, fi, pre[o4 * 4 + 2]);
                    pre[o4 * 4 + 3] = fmaf(w.w, fi, pre[o4 * 4 + 3]);
                }
            }
        }
    }

    for (int s = 0; s < a.S; ++s) {
        __syncthreads();   // previous member's zb fully consumed
        if (tid < FC) {
            float v = a.b0[tid];
            const float* zp = a.z + ((long long)n * a.S + s) * a.L;
            for (int l = 0; l < a.L; ++l) v = fmaf(a.w0[(long long)tid * K0 + FC + l], zp[l], v);
            zb[tid] = v;
        }
        __syncthreads();
        float h1[FC];
#pragma unroll
        for (int o = 0; o < FC; ++o) h1[o] = relu_f(pre[o] + zb[o]);
        if (a.h1_out && valid) {
            T* hp = reinterpret_cast<T*>(a.h1_out) + ((long long)n * a.HW + p) * FC;
#pragma unroll
            for (int o = 0; o < FC; o += 8) {
                float v8[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) v8[e] = h1[o + e];
                st8(hp + o, v8);
            }
        }
        float out[3] = {0.f, 0.f, 0.f};
#pragma unroll 1
        for (int o8 = 0; o8 < FC; o8 += 8) {
            float r8[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const int o2 = o8 + e;
                const float4* wr = reinterpret_cast<const float4*>(&w1[o2][0]);
                float acc = b1s[o2];
#pragma unroll
                for (int i4 = 0; i4 < FC / 4; ++i4) {
                    const float4 w = wr[i4];
                    acc = fmaf(w.x, h1[i4 * 4 + 0], acc);
                    acc = fmaf(w.y, h1[i4 * 4 + 1], acc);
                    acc = fmaf(w.z, h1[i4 * 4 + 2], acc);
                    acc = fmaf(w.w, h1[i4 * 4 + 3], acc);
                }
                const float r = relu_f(acc);
                r8[e] = r;
#pragma unroll
                for (int c = 0; c < 3; ++c) out[c] = fmaf(w2[c][o2], r, out[c]);
            }
            if (a.h2_out && valid) {
                T* hp = reinterpret_cast<T*>(a.h2_out) + ((long long)n * a.HW + p) * FC + o8;
                st8(hp, r8);
            }
        }
        if (valid) {
            for (int c = 0; c < a.num_classes; ++c)
                a.out_nchw[(((long long)n * a.S + s) * a.num_classes + c) * a.HW + p] = out[c] + b2s[c];
        }
    }
}

// gradients that flow through the z half of layer 0:  R[n][o] = sum_pixels dpre1[n][p][o] = rmean * HW
//   dw0[o][64 + l] (+)= sum_n R[n][o] z[n][l] ;  db0[o] (+)= sum_n R[n][o] ;  dz[n][l] = sum_o w0[o][64+l] R[n][o]
__global__ void fcomb_z_bwd_kernel(const float* __restrict__ rmean, float hw, const float* __restrict__ z,
                                   const float* __restrict__ w0, float* __restrict__ dz, float* __restrict__ dw0,
                                   float* __restrict__ db0, int N, int L, int accumulate) {
    const int K0 = FC + L;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < FC * L) {
        const int o = i / L, l = i % L;
        float s = 0.f;
        for (int n = 0; n < N; ++n) s = fmaf(rmean[n * FC + o] * hw, z[n * L + l], s);
        float* d = dw0 + (long long)o * K0 + FC + l;
        *d = accumulate ? *d + s : s;
    }
    if (i < FC) {
        float s = 0.f;
        for (int n = 0; n < N; ++n) s += rmean[n * FC + i] * hw;
        db0[i] = accumulate ? db0[i] + s : s;
    }
    if (i < N * L) {
        const int n = i / L, l = i % L;
        float s = 0.f;
        for (int o = 0; o < FC; ++o) s = fmaf(w0[(long long)o * K0 + FC + l], rmean[n * FC + o] * hw, s);
        dz[i] = s;
    }
}

// rsample backward: dmu += dz ; dls += dz * eps * sigma
__global__ void rsample_bwd_kernel(const float* dz, const float* eps, const float* sigma, float* dmu, float* dls, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    dmu[i] += dz[i];
    dls[i] += dz[i] * eps[i] * sigma[i];
}

}  // namespace pu

extern "C" {
using namespace pu;

int pu_fcomb_fwd(const PuFcombArgs* a, void* stream) {
    PU_REQUIRE(a && a->feat && a->z && a->w0 && a->b0 && a->w1 && a->b1 && a->w2 && a->b2 && a->out_nchw,
               "pu_fcomb_fwd: null pointer");
    PU_REQUIRE(a->N > 0 && a->HW > 0 && a->L > 0 && a->S > 0, "pu_fcomb_fwd: bad shape");
    PU_REQUIRE(a->num_classes >= 1 && a->num_classes <= 3, "pu_fcomb_fwd: num_classes must be 1..3 (got %d)", a->num_classes);
    PU_REQUIRE(a->S == 1 || (!a->h1_out && !a->h2_out), "pu_fcomb_fwd: hidden activations can only be saved for S == 1");
    cudaStream_t st = (cudaStream_t)stream;
    if (fcomb_members_tc_applicable(a)) return fcomb_members_tc_launch(a, st);
    dim3 grid(cdiv(a->HW, 128), a->N);
    if (a->dtype == PU_F32)
        fcomb_fwd_kernel<float><<<grid, 128, 0, st>>>(*a);
    else
        fcomb_fwd_kernel<__nv_bfloat16><<<grid, 128, 0, st>>>(*a);
    return check_launch("fcomb_fwd");
}

int pu_fcomb_z_bwd(const float* rmean, float hw, const float* z, const float* w0, float* dz, float* dw0, float* db0,
                   int N, int L, int accumulate, void* stream) {
    PU_REQUIRE(rmean && z && w0 && dz && dw0 && db0 && N > 0 && L > 0, "pu_fcomb_z_bwd: bad arguments");
    int total = FC * L;
    if (N * L > total) total = N * L;
    if (FC > total) total = FC;
    fcomb_z_bwd_kernel<<<cdiv(total, 128), 128, 0, (cudaStream_t)stream>>>(rmean, hw, z, w0, dz, dw0, db0, N, L, accumulate);
    return check_launch("fcomb_z_bwd");
}

int pu_rsample_bwd(const float* dz, const float* eps, const float* sigma, float* dmu, float* dls, int n, void* stream) {
    PU_REQUIRE(dz && eps && sigma && dmu && dls && n > 0, "pu_rsample_bwd: bad arguments");
    rsample_bwd_kernel<<<cdiv(n, 128), 128, 0, (cudaStream_t)stream>>>(dz, eps, sigma, dmu, dls, n);
    return check_launch("rsample_bwd");
}

// fused AdamW (torch.optim.AdamW semantics, main.py:95)
__global__ void adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                             float* __restrict__ v, long long n, float lr, float b1, float b2, float eps, float wd,
                             float bc1, float bc2_sqrt) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        float pi = p[i] * (1.f - lr * wd);
        const float gi = g[i];
        const float mi = b1 * m[i] + (1.f - b1) * gi;
        const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
        m[i] = mi;
        v[i] = vi;
        const float denom = sqrtf(vi) / bc2_sqrt + eps;
        p[i] = pi - (lr / bc1) * (mi / denom);
    }
}

int pu_adamw(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2, float eps,
             float weight_decay, int step, void* stream) {
    PU_REQUIRE(p && g && m && v && n > 0 && step >= 1, "pu_adamw: bad arguments");
    const float bc1 = 1.f - powf(beta1, (float)step);
    const float bc2 = 1.f - powf(beta2, (float)step);
    long long grid = cdivll(n, 256);
    if (grid > 148 * 8) grid = 148 * 8;
    adamw_kernel<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(p, g, m, v, n, lr, beta1, beta2, eps, weight_decay,
                                                                    bc1, sqrtf(bc2));
    return check_launch("adamw");
}
}
