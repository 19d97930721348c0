// CUDA-core flash-style self-attention (fp32 math, online softmax; no T x T matrix in HBM).
// This is the fp32-mode path and the fallback for shapes the tcgen05 attention kernel does not take.
// Reference: networks.py:112-125 (AttentionOp fwd/bwd) and :179-184.  Head dim is fixed at 64
// (channels_per_head=64, networks.py:143).
//
// Layout: qkv [N][T][3C], channel = j*C + head*64 + d (j: 0=q, 1=k, 2=v); out/dout [N][T][C], channel = head*64+d.
#include "../../include/probunet_b200.h"
#include "common.cuh"
#include "attn_internal.h"

namespace pu {

constexpr int AD = 64;       // head dim
constexpr int AQ = 64;       // queries (threads) per block
constexpr int AKT = 32;      // keys per smem tile

template <typename T>
__global__ void __launch_bounds__(AQ) attn_fwd_simple(const T* __restrict__ qkv, T* __restrict__ out,
                                                       float* __restrict__ lse, int T_, int heads) {
    __shared__ float sk[AKT][AD];
    __shared__ float sv[AKT][AD];
    const int C = heads * AD;
    const int nh = blockIdx.y, n = nh / heads, h = nh % heads;
    const int t = blockIdx.x * AQ + threadIdx.x;
    const bool valid = t < T_;
    const T* base = qkv + (long long)n * T_ * 3 * C;
    float q[AD], acc[AD];
#pragma unroll
    for (int d = 0; d < AD; ++d) acc[d] = 0.f;
    if (valid) {
        const T* qp = base + (long long)t * 3 * C + h * AD;
#pragma unroll
        for (int d = 0; d < AD; d += 8) {
            float v8[8];
            ld8(qp + d, v8);
#pragma unroll
            for (int e = 0; e < 8; ++e) q[d + e] = v8[e] * 0.125f;   // 1/sqrt(64)
        }
    } else {
#pragma unroll
        for (int d = 0; d < AD; ++d) q[d] = 0.f;
    }
    float m = -INFINITY, l = 0.f;
    for (int k0 = 0; k0 < T_; k0 += AKT) {
        __syncthreads();
        for (int i = threadIdx.x; i < AKT * AD / 8; i += AQ) {
            const int j = i / (AD / 8), d = (i % (AD / 8)) * 8;
            float kv[8], vv[8];
            if (k0 + j < T_) {
                const T* kp = base + (long long)(k0 + j) * 3 * C + C + h * AD + d;
                ld8(kp, kv);
                ld8(kp + C, vv);
            } else {
#pragma unroll
                for (int e = 0; e < 8; ++e) kv[e] = vv[e] = 0.f;
            }
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                sk[j][d + e] = kv[e];
                sv[j][d + e] = vv[e];
            }
        }
        __syncthreads();
        const int jmax = (T_ - k0) < AKT ? (T_ - k0) : AKT;
        for (int j = 0; j < jmax; ++j) {
            float s = 0.f;
#pragma unroll
            for (int d = 0; d < AD; ++d) s = fmaf(q[d], sk[j][d], s);
            const float mn = fmaxf(m, s);
            const float corr = expf(m - mn);
            const float pj = expf(s - mn);
            l = l * corr + pj;
#pragma unroll
            for (int d = 0; d < AD; ++d) acc[d] = fmaf(acc[d], corr, pj * sv[j][d]);
            m = mn;
        }
    }
    if (valid) {
        const float inv = 1.f / l;
        T* op = out + ((long long)n * T_ + t) * C + h * AD;
#pragma unroll
        for (int d = 0; d < AD; d += 8) {
            float o8[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) o8[e] = acc[d + e] * inv;
            st8(op + d, o8);
        }
        lse[((long long)n * heads + h) * T_ + t] = m + logf(l);
    }
}

// delta[n][h][t] = sum_d out * dout
template <typename T>
__global__ void attn_delta_kernel(const T* __restrict__ out, const T* __restrict__ dout, float* __restrict__ delta,
                                  int N, int T_, int heads) {
    const int C = heads * AD;
    long long total = (long long)N * T_ * heads;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int h = (int)(i % heads);
    const long long nt = i / heads;
    const int t = (int)(nt % T_);
    const int n = (int)(nt / T_);
    const T* op = out + nt * C + h * AD;
    const T* dp = dout + nt * C + h * AD;
    float s = 0.f;
#pragma unroll
    for (int d = 0; d < AD; d += 8) {
        float a[8], b[8];
        ld8(op + d, a);
        ld8(dp + d, b);
#pragma unroll
        for (int e = 0; e < 8; ++e) s = fmaf(a[e], b[e], s);
    }
    delta[((long long)n * heads + h) * T_ + t] = s;
}

// dq: one thread per query, loop over keys
template <typename T>
__global__ void __launch_bounds__(AQ) attn_bwd_dq_simple(const T* __restrict__ qkv, const T* __restrict__ dout,
                                                          const float* __restrict__ lse, const float* __restrict__ delta,
                                                          T* __restrict__ dqkv, int T_, int heads) {
    __shared__ float sk[AKT][AD];
    __shared__ float sv[AKT][AD];
    const int C = heads * AD;
    const int nh = blockIdx.y, n = nh / heads, h = nh % heads;
    const int t = blockIdx.x * AQ + threadIdx.x;
    const bool valid = t < T_;
    const T* base = qkv + (long long)n * T_ * 3 * C;
    float q[AD], go[AD], dq[AD];
    float L = 0.f, Dl = 0.f;
#pragma unroll
    for (int d = 0; d < AD; ++d) dq[d] = 0.f;
    if (valid) {
        const T* qp = base + (long long)t * 3 * C + h * AD;
        const T* gp = dout + ((long long)n * T_ + t) * C + h * AD;
#pragma unroll
        for (int d = 0; d < AD; d += 8) {
            float a[8], b[8];
            ld8(qp + d, a);
            ld8(gp + d, b);
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                q[d + e] = a[e] * 0.125f;
                go[d + e] = b[e];
            }
        }
        L = lse[((long long)n * heads + h) * T_ + t];
        Dl = delta[((long long)n * heads + h) * T_ + t];
    } else {
#pragma unroll
        for (int d = 0; d < AD; ++d) q[d] = go[d] = 0.f;
    }
    for (int k0 = 0; k0 < T_; k0 += AKT) {
        __syncthreads();
        for (int i = threadIdx.x; i < AKT * AD / 8; i += AQ) {
            const int j = i / (AD / 8), d = (i % (AD / 8)) * 8;
            float kv[8], vv[8];
            if (k0 + j < T_) {
                const T* kp = base + (long long)(k0 + j) * 3 * C + C + h * AD + d;
                ld8(kp, kv);
                ld8(kp + C, vv);
            } else {
#pragma unroll
                for (int e = 0; e < 8; ++e) kv[e] = vv[e] = 0.f;
            }
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                sk[j][d + e] = kv[e];
                sv[j][d + e] = vv[e];
            }
        }
        __syncthreads();
        const int jmax = (T_ - k0) < AKT ? (T_ - k0) : AKT;
        for (int j = 0; j < jmax; ++j) {
            float s = 0.f, dp = 0.f;
#pragma unroll
            for (int d = 0; d < AD; ++d) {
                s = fmaf(q[d], sk[j][d], s);
                dp = fmaf(go[d], sv[j][d], dp);
            }
            const float p = expf(s - L);
            const float ds = p * (dp - Dl) * 0.125f;
#pragma unroll
            for (int d = 0; d < AD; ++d) dq[d] = fmaf(ds, sk[j][d], dq[d]);
        }
    }
    if (valid) {
        T* dp_ = dqkv + ((long long)n * T_ + t) * 3 * C + h * AD;
#pragma unroll
        for (int d = 0; d < AD; d += 8) {
            float o8[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) o8[e] = dq[d + e];
            st8(dp_ + d, o8);
        }
    }
}

// dk, dv: one thread per key, loop over queries
template <typename T>
__global__ void __launch_bounds__(AQ) attn_bwd_dkv_simple(const T* __restrict__ qkv, const T* __restrict__ dout,
                                                           const float* __restrict__ lse, const float* __restrict__ delta,
                                                           T* __restrict__ dqkv, int T_, int heads) {
    __shared__ float sq[AKT][AD];
    __shared__ float sg[AKT][AD];
    __shared__ float sl[AKT], sd[AKT];
    const int C = heads * AD;
    const int nh = blockIdx.y, n = nh / heads, h = nh % heads;
    const int t = blockIdx.x * AQ + threadIdx.x;   // key index
    const bool valid = t < T_;
    const T* base = qkv + (long long)n * T_ * 3 * C;
    float k[AD], v[AD], dk[AD], dv[AD];
#pragma unroll
    for (int d = 0; d < AD; ++d) dk[d] = dv[d] = 0.f;
    if (valid) {
        const T* kp = base + (long long)t * 3 * C + C + h * AD;
#pragma unroll
        for (int d = 0; d < AD; d += 8) {
            float a[8], b[8];
            ld8(kp + d, a);
            ld8(kp + C + d, b);
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                k[d + e] = a[e];
                v[d + e] = b[e];
            }
        }
    } else {
#pragma unroll
        for (int d = 0; d < AD; ++d) k[d] = v[d] = 0.f;
    }
    for (int q0 = 0; q0 < T_; q0 += AKT) {
        __syncthreads();
        for (int i = threadIdx.x; i < AKT * AD / 8; i += AQ) {
            const int j = i / (AD / 8), d = (i % (AD / 8)) * 8;
            float a[8], b[8];
            if (q0 + j < T_) {
                ld8(base + (long long)(q0 + j) * 3 * C + h * AD + d, a);
                ld8(dout + ((long long)n * T_ + q0 + j) * C + h * AD + d, b);
            } else {
#pragma unroll
                for (int e = 0; e < 8; ++e) a[e] = b[e] = 0.f;
            }
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                sq[j][d + e] = a[e] * 0.125f;
                sg[j][d + e] = b[e];
            }
        }
        for (int j = threadIdx.x; j < AKT; j += AQ) {
            const bool ok = q0 + j < T_;
            sl[j] = ok ? lse[((long long)n * heads + h) * T_ + q0 + j] : 0.f;
            sd[j] = ok ? delta[((long long)n * heads + h) * T_ + q0 + j] : 0.f;
        }
        __syncthreads();
        const int jmax = (T_ - q0) < AKT ? (T_ - q0) : AKT;
        for (int j = 0; j < jmax; ++j) {
            float s = 0.f, dp = 0.f;
#pragma unroll
            for (int d = 0; d < AD; ++d) {
                s = fmaf(sq[j][d], k[d], s);
                dp = fmaf(sg[j][d], v[d], dp);
            }
            const float p = expf(s - sl[j]);
            const float ds = p * (dp - sd[j]);
#pragma unroll
            for (int d = 0; d < AD; ++d) {
                dv[d] = fmaf(p, sg[j][d], dv[d]);
                dk[d] = fmaf(ds, sq[j][d], dk[d]);   // sq already carries the 1/sqrt(d) factor
            }
        }
    }
    if (valid) {
        T* kp = dqkv + ((long long)n * T_ + t) * 3 * C + C + h * AD;
#pragma unroll
        for (int d = 0; d < AD; d += 8) {
            float a[8], b[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                a[e] = dk[d + e];
                b[e] = dv[d + e];
            }
            st8(kp + d, a);
            st8(kp + C + d, b);
        }
    }
}

template <typename T>
static int attn_fwd_simple_t(const void* qkv, void* out, float* lse, int N, int T_, int heads, cudaStream_t st) {
    dim3 grid(cdiv(T_, AQ), N * heads);
    attn_fwd_simple<T><<<grid, AQ, 0, st>>>((const T*)qkv, (T*)out, lse, T_, heads);
    return check_launch("attn_fwd_simple");
}

int attention_fwd_simple(const void* qkv, void* out, float* lse, int N, int T_, int heads, int dtype, cudaStream_t st) {
    return dtype == PU_F32 ? attn_fwd_simple_t<float>(qkv, out, lse, N, T_, heads, st)
                           : attn_fwd_simple_t<__nv_bfloat16>(qkv, out, lse, N, T_, heads, st);
}

template <typename T>
static int attn_delta_t(const void* out, const void* dout, float* delta, int N, int T_, int heads, cudaStream_t st) {
    long long total = (long long)N * T_ * heads;
    attn_delta_kernel<T><<<(unsigned)cdivll(total, 128), 128, 0, st>>>((const T*)out, (const T*)dout, delta, N, T_, heads);
    return check_launch("attn_delta");
}

int attention_delta(const void* out, const void* dout, float* delta, int N, int T_, int heads, int dtype,
                    cudaStream_t st) {
    return dtype == PU_F32 ? attn_delta_t<float>(out, dout, delta, N, T_, heads, st)
                           : attn_delta_t<__nv_bfloat16>(out, dout, delta, N, T_, heads, st);
}

template <typename T>
static int attn_bwd_simple_t(const void* qkv, const void* dout, const float* lse, const float* delta, void* dqkv, int N,
                             int T_, int heads, cudaStream_t st) {
    dim3 grid(cdiv(T_, AQ), N * heads);
    attn_bwd_dq_simple<T><<<grid, AQ, 0, st>>>((const T*)qkv, (const T*)dout, lse, delta, (T*)dqkv, T_, heads);
    int rc = check_launch("attn_bwd_dq_simple");
    if (rc) return rc;
    attn_bwd_dkv_simple<T><<<grid, AQ, 0, st>>>((const T*)qkv, (const T*)dout, lse, delta, (T*)dqkv, T_, heads);
    return check_launch("attn_bwd_dkv_simple");
}

int attention_bwd_simple(const void* qkv, const void* dout, const float* lse, const float* delta, void* dqkv, int N,
                         int T_, int heads, int dtype, cudaStream_t st) {
    return dtype == PU_F32 ? attn_bwd_simple_t<float>(qkv, dout, lse, delta, dqkv, N, T_, heads, st)
                           : attn_bwd_simple_t<__nv_bfloat16>(qkv, dout, lse, delta, dqkv, N, T_, heads, st);
}

}  // namespace pu

extern "C" {
int pu_attention_fwd(const void* qkv, void* out, float* lse, int N, int T, int heads, int dtype, int flags,
                     void* stream) {
    PU_REQUIRE(qkv && out && lse && N > 0 && T > 0 && heads > 0, "pu_attention_fwd: bad arguments");
    PU_REQUIRE(dtype == PU_F32 || dtype == PU_BF16, "pu_attention_fwd: bad dtype");
    cudaStream_t st = (cudaStream_t)stream;
    if (!(flags & PU_CONV_FORCE_SIMPLE) && pu::attention_tc_applicable(N, T, heads, dtype))
        return pu::attention_fwd_tc(qkv, out, lse, N, T, heads, st);
    PU_REQUIRE(!(flags & PU_CONV_FORCE_TC), "pu_attention_fwd: tcgen05 kernel does not apply (T=%d dtype=%d)", T, dtype);
    return pu::attention_fwd_simple(qkv, out, lse, N, T, heads, dtype, st);
}

int pu_attention_bwd(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv,
                     float* delta_ws, float* dq_ws, float* dbias, int N, int T, int heads, int dtype, int flags,
                     void* stream) {
    PU_REQUIRE(qkv && out && dout && lse && dqkv && delta_ws && dq_ws && N > 0 && T > 0 && heads > 0,
               "pu_attention_bwd: bad arguments");
    PU_REQUIRE(dtype == PU_F32 || dtype == PU_BF16, "pu_attention_bwd: bad dtype");
    cudaStream_t st = (cudaStream_t)stream;
    int rc = pu::attention_delta(out, dout, delta_ws, N, T, heads, dtype, st);
    if (rc) return rc;
    bool dbias_done = false;
    if (!(flags & PU_CONV_FORCE_SIMPLE) && pu::attention_bwd_tc_applicable(N, T, heads, dtype)) {
        rc = pu::attention_bwd_tc(qkv, dout, lse, delta_ws, dqkv, dq_ws, dbias, &dbias_done, N, T, heads, st);
    } else {
        PU_REQUIRE(!(flags & PU_CONV_FORCE_TC), "pu_attention_bwd: tcgen05 kernel does not apply (T=%d dtype=%d)", T, dtype);
        rc = pu::attention_bwd_simple(qkv, dout, lse, delta_ws, dqkv, N, T, heads, dtype, st);
    }
    if (rc || !dbias || dbias_done) return rc;
    // paths without the fused column sums: the separate pass
    return pu_bias_grad(dqkv, dbias, (long long)N * T, 3 * heads * 64, dtype, 0, stream);
}
}
