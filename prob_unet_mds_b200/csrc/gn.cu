// GroupNorm (+adaptive scale/shift, SiLU, dropout, 2x resample) forward and backward, NHWC, HBM-bound.
// Replaces F.group_norm / silu / addcmul / dropout and the depthwise resample convs around them
// (networks.py:104,166,170-171,175,82-85).  Statistics are accumulated in fp64.
//
// Thread mapping (all kernels): a block is a (channel-vector, pixel-lane) grid of 256 threads.  Thread (v, pl) owns the
// 8 channels [8v, 8v+8) -- its per-channel constants (mean, rstd, gamma', beta') live in registers -- and walks the
// pixels pl, pl+PL, ... of the block's pixel range, so a warp reads/writes contiguous 16-byte vectors and no integer
// division is needed per element.  Algorithmic traffic per element: apply 2B read + 2B write (bf16); backward
// 4B read (reduce pass) + 4..6B read + 2B write (apply pass).
#include "../../include/probunet_b200.h"
#include "common.cuh"
#include "conv_internal.h"
#include "tc_ptx.cuh"

#include <stdlib.h>

namespace pu {

constexpr int GN_THREADS = 256;
constexpr int GN_BWD_BLOCKS = 2;   // resident blocks per SM of the backward kernels (128 registers, no spills)

// FAST (bf16 kernels): sigmoid(u) = 0.5 tanh(u/2) + 0.5 with the single-MUFU tanh.approx (relative error ~2^-11, far
// below bf16 rounding) instead of ex2 + rcp -- the GroupNorm kernels are bound by instruction issue, not by HBM.
template <bool FAST>
__device__ __forceinline__ float sigmoid_t(float u) {
    if (FAST) {
        float t;
        asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * u));
        return fmaf(t, 0.5f, 0.5f);
    }
    return 1.f / (1.f + expf(-u));
}

// Per-channel constants.  xhat = (x - mu) * rstd and u = xhat * gamma' + beta' are evaluated as single FMAs from x in
// the bf16 kernels (FAST): xhat = x * rstd + nmr, u = x * ag + bg.  The fp32 kernels keep the subtract-first form,
// which does not lose precision when |mean| >> std.
struct ChanConst {
    float mu[8], rstd[8], gam[8], bet[8];
    float nmr[8], ag[8], bg[8];        // -mu*rstd, rstd*gamma', beta' - mu*rstd*gamma'
};

template <bool FAST>
__device__ __forceinline__ float gn_xhat(const ChanConst& k, int e, float x) {
    return FAST ? fmaf(x, k.rstd[e], k.nmr[e]) : (x - k.mu[e]) * k.rstd[e];
}
template <bool FAST>
__device__ __forceinline__ float gn_u(const ChanConst& k, int e, float x) {
    return FAST ? fmaf(x, k.ag[e], k.bg[e]) : fmaf((x - k.mu[e]) * k.rstd[e], k.gam[e], k.bet[e]);
}

// per-thread constants of channels [c0, c0+8) of sample n
__device__ __forceinline__ void gn_load_consts(const PuGnArgs& f, int n, int c0, ChanConst& k) {
    const int C = f.C0 + f.C1;
    const int Cg = C / f.G;
    const double m = (double)Cg * f.H * f.W;
    int gprev = -1;
    float mean = 0.f, rstd = 0.f;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        const int c = c0 + e;
        const int g = c / Cg;
        if (g != gprev) {
            const double sum = f.stats[((long long)n * f.G + g) * 2];
            const double ssq = f.stats[((long long)n * f.G + g) * 2 + 1];
            const double mm = sum / m;
            double var = ssq / m - mm * mm;
            if (var < 0) var = 0;
            mean = (float)mm;
            rstd = (float)(1.0 / sqrt(var + (double)f.eps));
            gprev = g;
        }
        float gam = f.gamma[c], bet = f.beta[c];
        if (f.ada) {
            const float sc = f.ada[c], sh = f.ada[C + c];
            gam = gam * (1.f + sc);
            bet = fmaf(bet, 1.f + sc, sh);
        }
        k.mu[e] = mean;
        k.rstd[e] = rstd;
        k.gam[e] = gam;
        k.bet[e] = bet;
        k.nmr[e] = -mean * rstd;
        k.ag[e] = rstd * gam;
        k.bg[e] = fmaf(-mean * rstd, gam, bet);
    }
}

// 8 channels exactly as loaded (kept packed so that several loads can be in flight per thread)
template <typename T>
struct Raw8;
template <>
struct Raw8<__nv_bfloat16> {
    uint4 v;
};
template <>
struct Raw8<float> {
    float4 a, b;
};
__device__ __forceinline__ Raw8<__nv_bfloat16> ldraw(const __nv_bfloat16* p) {
    Raw8<__nv_bfloat16> r;
    r.v = *reinterpret_cast<const uint4*>(p);
    return r;
}
__device__ __forceinline__ Raw8<float> ldraw(const float* p) {
    Raw8<float> r;
    r.a = *reinterpret_cast<const float4*>(p);
    r.b = *reinterpret_cast<const float4*>(p + 4);
    return r;
}
__device__ __forceinline__ void unpack(const Raw8<__nv_bfloat16>& r, float (&v)[8]) {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r.v);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 f = __bfloat1622float2(h[i]);
        v[2 * i] = f.x;
        v[2 * i + 1] = f.y;
    }
}
__device__ __forceinline__ void unpack(const Raw8<float>& r, float (&v)[8]) {
    v[0] = r.a.x; v[1] = r.a.y; v[2] = r.a.z; v[3] = r.a.w;
    v[4] = r.b.x; v[5] = r.b.y; v[6] = r.b.z; v[7] = r.b.w;
}

constexpr int GN_NB = 2;   // independent 16-byte loads in flight per thread and array

template <typename T>
__device__ __forceinline__ void gn_load_x8(const PuGnArgs& f, long long pix, int c0, float (&v)[8]) {
    if (c0 < f.C0)
        ld8(reinterpret_cast<const T*>(f.src0) + pix * f.C0 + c0, v);
    else
        ld8(reinterpret_cast<const T*>(f.src1) + pix * f.C1 + (c0 - f.C0), v);
}

// ---- statistics ----
template <typename T>
__global__ void __launch_bounds__(GN_THREADS)
gn_stats_kernel(const T* __restrict__ s0, const T* __restrict__ s1, int C0, int C1, int HW, int G, int rows,
                double* __restrict__ stats) {
    // fp32 only inside one trip (four rows): the running sums per thread, per block and per sample are fp64.  (With fp32
    // thread and block sums the statistics depended on the block partition at the 1e-6 level -- enough to move the loss
    // by 1.3e-5 between a batch and its two halves, and the fp32-mode gradients by 5e-5.)
    extern __shared__ double smd_stats[];   // [G][2]
    double* sm = smd_stats;
    const int C = C0 + C1, nvec = C / 8, Cg = C / G;
    const int n = blockIdx.y;
    const int PL = GN_THREADS / nvec;
    const int v = threadIdx.x % nvec, pl = threadIdx.x / nvec;
    for (int i = threadIdx.x; i < 2 * G; i += blockDim.x) sm[i] = 0.0;
    __syncthreads();
    const int r0 = blockIdx.x * rows;
    int r1 = r0 + rows;
    if (r1 > HW) r1 = HW;
    if (pl < PL) {
        double s[8], q[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) s[e] = q[e] = 0.0;
        const int c0 = v * 8;
        const T* base = (c0 < C0) ? (s0 + (long long)n * HW * C0 + c0) : (s1 + (long long)n * HW * C1 + (c0 - C0));
        const int stride = (c0 < C0) ? C0 : C1;
        constexpr int SNB = 4;     // rows per trip; the next trip's loads are issued before this trip's math
        Raw8<T> raw[SNB];
#pragma unroll
        for (int j = 0; j < SNB; ++j) {
            const int rr = r0 + pl + j * PL;
            if (rr < r1) raw[j] = ldraw(base + (long long)rr * stride);
        }
        for (int r = r0 + pl; r < r1; r += SNB * PL) {
            Raw8<T> nxt[SNB];
#pragma unroll
            for (int j = 0; j < SNB; ++j) {
                const int rr = r + (SNB + j) * PL;
                if (rr < r1) nxt[j] = ldraw(base + (long long)rr * stride);
            }
            float ts[8], tq[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) ts[e] = tq[e] = 0.f;
#pragma unroll
            for (int j = 0; j < SNB; ++j) {
                if (r + j * PL >= r1) break;
                float x[8];
                unpack(raw[j], x);
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    ts[e] += x[e];
                    tq[e] = fmaf(x[e], x[e], tq[e]);
                }
            }
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                s[e] += (double)ts[e];
                q[e] += (double)tq[e];
            }
#pragma unroll
            for (int j = 0; j < SNB; ++j) raw[j] = nxt[j];
        }
        // combine channels of the same group before touching shared memory
        int g = c0 / Cg;
        double gs = 0.0, gq = 0.0;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int ge = (c0 + e) / Cg;
            if (ge != g) {
                atomicAdd(&sm[2 * g], gs);
                atomicAdd(&sm[2 * g + 1], gq);
                g = ge;
                gs = gq = 0.0;
            }
            gs += s[e];
            gq += q[e];
        }
        atomicAdd(&sm[2 * g], gs);
        atomicAdd(&sm[2 * g + 1], gq);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * G; i += blockDim.x) atomicAdd(stats + (long long)n * G * 2 + i, sm[i]);
}

// ---- statistics per quad of channels (the layout conv_tc_kernel's epilogue produces), CUDA-core fallback ----
template <typename T>
__global__ void __launch_bounds__(GN_THREADS)
gn_quad_stats_kernel(const T* __restrict__ x, int HW, int C, int rows, double* __restrict__ q) {
    const int nq = C / 4, n = blockIdx.y;
    const int r0 = blockIdx.x * rows;
    const int r1 = min(HW, r0 + rows);
    for (int qd = threadIdx.x; qd < nq; qd += blockDim.x) {        // consecutive threads: consecutive quads (coalesced)
        float s = 0.f, ss = 0.f;
        for (int r = r0; r < r1; ++r) {
            const T* p = x + ((long long)n * HW + r) * C + qd * 4;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float v = ldf(p + e);
                s += v;
                ss = fmaf(v, v, ss);
            }
        }
        atomicAdd(q + ((long long)n * nq + qd) * 2, (double)s);
        atomicAdd(q + ((long long)n * nq + qd) * 2 + 1, (double)ss);
    }
}

int gn_quad_stats_launch(const void* x, int dtype, int N, int HW, int C, double* q, cudaStream_t st) {
    const int rows = 32;
    dim3 grid(cdiv(HW, rows), N);
    if (dtype == PU_F32)
        gn_quad_stats_kernel<float><<<grid, GN_THREADS, 0, st>>>((const float*)x, HW, C, rows, q);
    else
        gn_quad_stats_kernel<__nv_bfloat16><<<grid, GN_THREADS, 0, st>>>((const __nv_bfloat16*)x, HW, C, rows, q);
    return check_launch("gn_quad_stats");
}

// stats[n][g] = sum over the group's quads (of src0's table, then src1's) -- a few thousand numbers
__global__ void gn_stats_from_quads_kernel(const double* __restrict__ q0, const double* __restrict__ q1, int C0, int C1,
                                           int N, int G, double* __restrict__ stats) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N * G) return;
    const int n = i / G, g = i - n * G;
    const int Cg = (C0 + C1) / G;
    double s = 0.0, ss = 0.0;
    for (int c = g * Cg; c < (g + 1) * Cg; c += 4) {
        const double* p = (c < C0) ? q0 + ((long long)n * (C0 / 4) + c / 4) * 2
                                   : q1 + ((long long)n * (C1 / 4) + (c - C0) / 4) * 2;
        s += p[0];
        ss += p[1];
    }
    stats[2 * i] = s;
    stats[2 * i + 1] = ss;
}

// ---- forward apply ----
template <typename T, bool FAST>
__global__ void __launch_bounds__(GN_THREADS, 3) gn_apply_kernel(PuGnArgs f, int rows) {
    const int C = f.C0 + f.C1, nvec = C / 8;
    const int n = blockIdx.y;
    const int PL = GN_THREADS / nvec;
    const int v = threadIdx.x % nvec, pl = threadIdx.x / nvec;
    if (pl >= PL) return;
    const int c0 = v * 8;
    ChanConst k;
    gn_load_consts(f, n, c0, k);
    const int OH = f.resample == PU_RS_UP ? f.H * 2 : (f.resample == PU_RS_DOWN ? f.H / 2 : f.H);
    const int OW = f.resample == PU_RS_UP ? f.W * 2 : (f.resample == PU_RS_DOWN ? f.W / 2 : f.W);
    const int r0 = blockIdx.x * rows;
    int r1 = r0 + rows;
    if (r1 > OH * OW) r1 = OH * OW;
    const float inv_keep = f.dropout_p > 0.f ? 1.f / (1.f - f.dropout_p) : 1.f;
    T* y = reinterpret_cast<T*>(f.y) + (long long)n * OH * OW * C + c0;
    const long long in_base = (long long)n * f.H * f.W;
    if (f.resample == PU_RS_NONE) {
        // fast path: GN_NB pixels per trip, all loads issued before the first use
        const T* xp;
        int stride;
        if (c0 < f.C0) {
            xp = reinterpret_cast<const T*>(f.src0) + in_base * f.C0 + c0;
            stride = f.C0;
        } else {
            xp = reinterpret_cast<const T*>(f.src1) + in_base * f.C1 + (c0 - f.C0);
            stride = f.C1;
        }
        // software pipeline: the loads of the next GN_NB pixels are in flight while the current ones are processed
        Raw8<T> raw[GN_NB];
#pragma unroll
        for (int j = 0; j < GN_NB; ++j) {
            const int rr = r0 + pl + j * PL;
            if (rr < r1) raw[j] = ldraw(xp + (long long)rr * stride);
        }
        for (int op = r0 + pl; op < r1; op += GN_NB * PL) {
            Raw8<T> nxt[GN_NB];
#pragma unroll
            for (int j = 0; j < GN_NB; ++j) {
                const int rr = op + (GN_NB + j) * PL;
                if (rr < r1) nxt[j] = ldraw(xp + (long long)rr * stride);
            }
#pragma unroll
            for (int j = 0; j < GN_NB; ++j) {
                const int rr = op + j * PL;
                if (rr >= r1) break;
                float x[8], o[8];
                unpack(raw[j], x);
                if constexpr (FAST) {
                    // packed fp32 (FFMA2 / FMUL2): u = x * ag + bg, sigmoid(u) = 0.5 tanh(u / 2) + 0.5, silu = u * sigmoid(u)
                    // -- this kernel is issue-bound (66 % issue-slot utilisation, ncu), not HBM-bound
#pragma unroll
                    for (int e = 0; e < 8; e += 2) {
                        const float2 u2 = ffma2(make_float2(x[e], x[e + 1]), make_float2(k.ag[e], k.ag[e + 1]),
                                                make_float2(k.bg[e], k.bg[e + 1]));
                        if (f.silu) {
                            float t0, t1;
                            asm("tanh.approx.f32 %0, %1;" : "=f"(t0) : "f"(0.5f * u2.x));
                            asm("tanh.approx.f32 %0, %1;" : "=f"(t1) : "f"(0.5f * u2.y));
                            const float2 s2 = ffma2(make_float2(t0, t1), make_float2(0.5f, 0.5f), make_float2(0.5f, 0.5f));
                            const float2 o2 = fmul2(u2, s2);
                            o[e] = o2.x;
                            o[e + 1] = o2.y;
                        } else {
                            o[e] = u2.x;
                            o[e + 1] = u2.y;
                        }
                    }
                } else {
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const float u = gn_u<FAST>(k, e, x[e]);
                        o[e] = f.silu ? u * sigmoid_t<FAST>(u) : u;
                    }
                }
                if (f.dropout_p > 0.f) {
                    const long long opix = in_base + rr;
                    const uint32_t keep = dropout_keep8(f.seed, (unsigned long long)((opix * C + c0) >> 3), f.dropout_p);
#pragma unroll
                    for (int e = 0; e < 8; ++e) o[e] = ((keep >> e) & 1u) ? o[e] * inv_keep : 0.f;
                    if (f.keep_mask) reinterpret_cast<uint8_t*>(f.keep_mask)[(opix * C + c0) >> 3] = (uint8_t)keep;
                }
                st8(y + (long long)rr * C, o);
            }
#pragma unroll
            for (int j = 0; j < GN_NB; ++j) raw[j] = nxt[j];
        }
        return;
    }
    for (int op = r0 + pl; op < r1; op += PL) {
        float o[8];
        if (f.resample == PU_RS_NONE) {
            float x[8];
            gn_load_x8<T>(f, in_base + op, c0, x);
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const float u = gn_u<FAST>(k, e, x[e]);
                o[e] = f.silu ? u * sigmoid_t<FAST>(u) : u;
            }
        } else if (f.resample == PU_RS_UP) {
            const int oy = op / OW, ox = op - oy * OW;
            float x[8];
            gn_load_x8<T>(f, in_base + (long long)(oy >> 1) * f.W + (ox >> 1), c0, x);
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const float u = gn_u<FAST>(k, e, x[e]);
                o[e] = f.silu ? u * sigmoid_t<FAST>(u) : u;
            }
        } else {
            const int oy = op / OW, ox = op - oy * OW;
#pragma unroll
            for (int e = 0; e < 8; ++e) o[e] = 0.f;
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                float x[8];
                gn_load_x8<T>(f, in_base + (long long)(oy * 2 + (t >> 1)) * f.W + ox * 2 + (t & 1), c0, x);
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const float u = gn_u<FAST>(k, e, x[e]);
                    o[e] += f.silu ? u * sigmoid_t<FAST>(u) : u;
                }
            }
#pragma unroll
            for (int e = 0; e < 8; ++e) o[e] *= 0.25f;
        }
        if (f.dropout_p > 0.f) {
            const long long opix = (long long)n * OH * OW + op;
            const uint32_t keep = dropout_keep8(f.seed, (unsigned long long)((opix * C + c0) >> 3), f.dropout_p);
#pragma unroll
            for (int e = 0; e < 8; ++e) o[e] = ((keep >> e) & 1u) ? o[e] * inv_keep : 0.f;
            if (f.keep_mask) reinterpret_cast<uint8_t*>(f.keep_mask)[(opix * C + c0) >> 3] = (uint8_t)keep;
        }
        st8(y + (long long)op * C, o);
    }
}

// ---- backward ----
// gradient wrt the (pre-resample) activation output at input pixel (iy, ix): gathers dy through the transpose
// of the forward resample
template <typename T>
__device__ __forceinline__ void gn_gather8(const T* dy, int rs, int n, int H, int W, int r, int C, int c0, float (&g)[8]) {
    if (rs == PU_RS_NONE) {
        ld8(dy + ((long long)n * H * W + r) * C + c0, g);
    } else if (rs == PU_RS_UP) {
        const int iy = r / W, ix = r - iy * W;
        const int OH = 2 * H, OW = 2 * W;
#pragma unroll
        for (int e = 0; e < 8; ++e) g[e] = 0.f;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            float v[8];
            ld8(dy + (((long long)n * OH + 2 * iy + (t >> 1)) * OW + 2 * ix + (t & 1)) * C + c0, v);
#pragma unroll
            for (int e = 0; e < 8; ++e) g[e] += v[e];
        }
    } else {
        const int iy = r / W, ix = r - iy * W;
        const int OH = H / 2, OW = W / 2;
        ld8(dy + (((long long)n * OH + (iy >> 1)) * OW + (ix >> 1)) * C + c0, g);
#pragma unroll
        for (int e = 0; e < 8; ++e) g[e] *= 0.25f;
    }
}

template <bool FAST>
__device__ __forceinline__ void gn_du8_calc(const PuGnArgs& f, const ChanConst& k, const float (&x)[8],
                                            const float (&g)[8], long long pix, int c0, float (&xh)[8], float (&du)[8]);

// du = d loss / d u  where u = xhat * gamma' + beta' and y = resample(dropout(act(u)))
template <typename T, bool FAST>
__device__ __forceinline__ void gn_du8(const PuGnArgs& f, const ChanConst& k, const T* dy, int n, int r, int c0,
                                       float (&xh)[8], float (&du)[8]) {
    const int C = f.C0 + f.C1;
    const long long pix = (long long)n * f.H * f.W + r;
    float x[8], g[8];
    gn_load_x8<T>(f, pix, c0, x);
    gn_gather8<T>(dy, f.resample, n, f.H, f.W, r, C, c0, g);
    gn_du8_calc<FAST>(f, k, x, g, pix, c0, xh, du);
}

// same, from already loaded x (pre-norm activation) and g (gradient wrt the activation output, resample undone)
template <bool FAST>
__device__ __forceinline__ void gn_du8_calc(const PuGnArgs& f, const ChanConst& k, const float (&x)[8],
                                            const float (&g)[8], long long pix, int c0, float (&xh)[8], float (&du)[8]) {
    const int C = f.C0 + f.C1;
    uint32_t keep = 0xffu;
    float inv_keep = 1.f;
    if (f.dropout_p > 0.f) {
        // the keep bits the forward stored (PuGnArgs.keep_mask), else regenerated from (seed, element index)
        keep = f.keep_mask ? (uint32_t) reinterpret_cast<const uint8_t*>(f.keep_mask)[(pix * C + c0) >> 3]
                           : dropout_keep8(f.seed, (unsigned long long)((pix * C + c0) >> 3), f.dropout_p);
        inv_keep = 1.f / (1.f - f.dropout_p);
    }
    if constexpr (FAST) {
        // packed fp32: xhat = x * rstd + nmr, u = x * ag + bg, d silu = s (1 + u (1 - s)) with s = 0.5 tanh(u / 2) + 0.5
        const float2 ik2 = make_float2(inv_keep, inv_keep), one2 = make_float2(1.f, 1.f), half2 = make_float2(0.5f, 0.5f);
#pragma unroll
        for (int e = 0; e < 8; e += 2) {
            const float2 x2 = make_float2(x[e], x[e + 1]);
            const float2 xh2 = ffma2(x2, make_float2(k.rstd[e], k.rstd[e + 1]), make_float2(k.nmr[e], k.nmr[e + 1]));
            xh[e] = xh2.x;
            xh[e + 1] = xh2.y;
            float2 gg = make_float2(((keep >> e) & 1u) ? g[e] : 0.f, ((keep >> (e + 1)) & 1u) ? g[e + 1] : 0.f);
            if (f.silu) {
                const float2 u2 = ffma2(x2, make_float2(k.ag[e], k.ag[e + 1]), make_float2(k.bg[e], k.bg[e + 1]));
                float t0, t1;
                asm("tanh.approx.f32 %0, %1;" : "=f"(t0) : "f"(0.5f * u2.x));
                asm("tanh.approx.f32 %0, %1;" : "=f"(t1) : "f"(0.5f * u2.y));
                const float2 s2 = ffma2(make_float2(t0, t1), half2, half2);
                const float2 oms = ffma2(s2, make_float2(-1.f, -1.f), one2);           // 1 - s
                const float2 d2 = fmul2(fmul2(s2, ik2), ffma2(u2, oms, one2));          // (s / keep) * (1 + u (1 - s))
                gg = fmul2(gg, d2);
            } else {
                gg = fmul2(gg, ik2);
            }
            du[e] = gg.x;
            du[e + 1] = gg.y;
        }
        return;
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        xh[e] = gn_xhat<FAST>(k, e, x[e]);
        float gg = ((keep >> e) & 1u) ? g[e] : 0.f;
        if (f.silu) {
            const float u = gn_u<FAST>(k, e, x[e]);
            const float s = sigmoid_t<FAST>(u);
            gg *= (s * inv_keep) * fmaf(u, 1.f - s, 1.f);
        } else {
            gg *= inv_keep;
        }
        du[e] = gg;
    }
}

template <typename T, bool FAST>
__global__ void __launch_bounds__(GN_THREADS, GN_BWD_BLOCKS) gn_bwd_reduce_kernel(PuGnBwdArgs a, int rows) {
    // [C][2] block partial sums: fp32 shared-memory atomics (native, fast) within the block, fp64 atomics across
    // blocks (these sums cancel heavily -- signed terms -- so the long cross-block accumulation is done in fp64).
    extern __shared__ float sm[];
    const PuGnArgs& f = a.f;
    const int C = f.C0 + f.C1, nvec = C / 8;
    const int n = blockIdx.y;
    const int PL = GN_THREADS / nvec;
    const int v = threadIdx.x % nvec, pl = threadIdx.x / nvec;
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) sm[i] = 0.f;
    __syncthreads();
    const int HW = f.H * f.W;
    const int r0 = blockIdx.x * rows;
    int r1 = r0 + rows;
    if (r1 > HW) r1 = HW;
    const T* dy = reinterpret_cast<const T*>(a.dy);
    if (pl < PL) {
        const int c0 = v * 8;
        ChanConst k;
        gn_load_consts(f, n, c0, k);
        float A[8], B[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) A[e] = B[e] = 0.f;
        if (f.resample == PU_RS_NONE) {
            const long long base = (long long)n * HW;
            const T* xp;
            int stride;
            if (c0 < f.C0) {
                xp = reinterpret_cast<const T*>(f.src0) + base * f.C0 + c0;
                stride = f.C0;
            } else {
                xp = reinterpret_cast<const T*>(f.src1) + base * f.C1 + (c0 - f.C0);
                stride = f.C1;
            }
            const T* gp = dy + base * C + c0;
            // software pipeline: the loads of the next GN_NB rows are in flight while the current ones are processed
            Raw8<T> xr[GN_NB], gr[GN_NB];
#pragma unroll
            for (int j = 0; j < GN_NB; ++j) {
                const int rr = r0 + pl + j * PL;
                if (rr < r1) {
                    xr[j] = ldraw(xp + (long long)rr * stride);
                    gr[j] = ldraw(gp + (long long)rr * C);
                }
            }
            for (int r = r0 + pl; r < r1; r += GN_NB * PL) {
                Raw8<T> xn[GN_NB], gn[GN_NB];
#pragma unroll
                for (int j = 0; j < GN_NB; ++j) {
                    const int rr = r + (GN_NB + j) * PL;
                    if (rr < r1) {
                        xn[j] = ldraw(xp + (long long)rr * stride);
                        gn[j] = ldraw(gp + (long long)rr * C);
                    }
                }
#pragma unroll
                for (int j = 0; j < GN_NB; ++j) {
                    const int rr = r + j * PL;
                    if (rr >= r1) break;
                    float x[8], g[8], xh[8], du[8];
                    unpack(xr[j], x);
                    unpack(gr[j], g);
                    gn_du8_calc<FAST>(f, k, x, g, base + rr, c0, xh, du);
                    // dy is a scratch tensor in this mode: overwrite it with du so that the apply pass does not
                    // have to redo the SiLU derivative and the dropout mask
                    if (!a.dres || a.dres_resample == PU_RS_NONE) st8(const_cast<T*>(gp) + (long long)rr * C, du);
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        A[e] += du[e];
                        B[e] = fmaf(du[e], xh[e], B[e]);
                    }
                }
#pragma unroll
                for (int j = 0; j < GN_NB; ++j) {
                    xr[j] = xn[j];
                    gr[j] = gn[j];
                }
            }
        } else {
            for (int r = r0 + pl; r < r1; r += PL) {
                float xh[8], du[8];
                gn_du8<T, FAST>(f, k, dy, n, r, c0, xh, du);
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    A[e] += du[e];
                    B[e] = fmaf(du[e], xh[e], B[e]);
                }
            }
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            atomicAdd(&sm[2 * (c0 + e)], A[e]);
            atomicAdd(&sm[2 * (c0 + e) + 1], B[e]);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x)
        atomicAdd(a.sums + (long long)n * C * 2 + i, (double)sm[i]);
}

template <typename T, bool FAST>
__device__ __forceinline__ void gn_bwd_apply_rows(const PuGnBwdArgs& a, int rows, int n, int v, int pl, int PL,
                                                  const double* smd, float (&cs)[8]);

template <typename T, bool FAST>
__global__ void __launch_bounds__(GN_THREADS, GN_BWD_BLOCKS) gn_bwd_apply_kernel(PuGnBwdArgs a, int rows) {
    extern __shared__ double smd[];   // [G][2]: sum_c gamma' A, sum_c gamma' B
    const PuGnArgs& f = a.f;
    const int C = f.C0 + f.C1, nvec = C / 8, Cg = C / f.G;
    const int n = blockIdx.y;
    const int PL = GN_THREADS / nvec;
    const int v = threadIdx.x % nvec, pl = threadIdx.x / nvec;
    float* csm = reinterpret_cast<float*>(smd + 2 * f.G);   // [C] column sums of the written gradient (bias grads)
    const bool want_cs = a.colsum0 != nullptr;
    for (int i = threadIdx.x; i < 2 * f.G; i += blockDim.x) smd[i] = 0.0;
    if (want_cs)
        for (int i = threadIdx.x; i < C; i += blockDim.x) csm[i] = 0.f;
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float gam = f.gamma[c];
        if (f.ada) gam *= 1.f + f.ada[c];
        const int g = c / Cg;
        atomicAdd(&smd[2 * g], (double)gam * a.sums[((long long)n * C + c) * 2]);
        atomicAdd(&smd[2 * g + 1], (double)gam * a.sums[((long long)n * C + c) * 2 + 1]);
    }
    __syncthreads();
    float cs[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) cs[e] = 0.f;
    if (pl < PL) gn_bwd_apply_rows<T, FAST>(a, rows, n, v, pl, PL, smd, cs);
    if (want_cs) {
        if (pl < PL) {
#pragma unroll
            for (int e = 0; e < 8; ++e) atomicAdd(&csm[v * 8 + e], cs[e]);
        }
        __syncthreads();
        for (int c = threadIdx.x; c < C; c += blockDim.x) {
            if (c < f.C0)
                atomicAdd(a.colsum0 + c, csm[c]);
            else if (a.colsum1)
                atomicAdd(a.colsum1 + (c - f.C0), csm[c]);
        }
    }
}

// the per-thread pixel loop of gn_bwd_apply_kernel; cs accumulates the column sums of what is written
template <typename T, bool FAST>
__device__ __forceinline__ void gn_bwd_apply_rows(const PuGnBwdArgs& a, int rows, int n, int v, int pl, int PL,
                                                  const double* smd, float (&cs)[8]) {
    const PuGnArgs& f = a.f;
    const int C = f.C0 + f.C1, Cg = C / f.G;
    const int c0 = v * 8;
    ChanConst k;
    gn_load_consts(f, n, c0, k);
    const double inv_m = 1.0 / ((double)Cg * (double)f.H * (double)f.W);
    float s1[8], s2[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        const int g = (c0 + e) / Cg;
        // dx = rstd * (du * gamma' - mean_g(gamma' du) - xhat * mean_g(gamma' du xhat)) = du * ag + s1 + xhat * s2
        s1[e] = -k.rstd[e] * (float)(smd[2 * g] * inv_m);
        s2[e] = -k.rstd[e] * (float)(smd[2 * g + 1] * inv_m);
    }
    const int HW = f.H * f.W;
    const int r0 = blockIdx.x * rows;
    int r1 = r0 + rows;
    if (r1 > HW) r1 = HW;
    const T* dy = reinterpret_cast<const T*>(a.dy);
    const T* dres = reinterpret_cast<const T*>(a.dres);
    T* dst;
    int stride, acc;
    if (c0 < f.C0) {
        dst = reinterpret_cast<T*>(a.dx0) + (long long)n * HW * f.C0 + c0;
        stride = f.C0;
        acc = a.acc0;
    } else {
        dst = reinterpret_cast<T*>(a.dx1) + (long long)n * HW * f.C1 + (c0 - f.C0);
        stride = f.C1;
        acc = a.acc1;
    }
    if (f.resample == PU_RS_NONE && (!dres || a.dres_resample == PU_RS_NONE)) {
        constexpr int NB = 2;
        const long long base = (long long)n * HW;
        const T* xp = (c0 < f.C0) ? reinterpret_cast<const T*>(f.src0) + base * f.C0 + c0
                                  : reinterpret_cast<const T*>(f.src1) + base * f.C1 + (c0 - f.C0);
        const T* gp = dy + base * C + c0;
        const T* rp = dres ? dres + base * C + c0 : nullptr;
        Raw8<T> xr[NB], gr[NB], rr_[NB], od[NB];
        auto load_rows = [&](int r, Raw8<T> (&xx)[NB], Raw8<T> (&gg)[NB], Raw8<T> (&dd)[NB], Raw8<T> (&oo)[NB]) {
#pragma unroll
            for (int j = 0; j < NB; ++j) {
                const int rr = r + j * PL;
                if (rr < r1) {
                    xx[j] = ldraw(xp + (long long)rr * stride);
                    gg[j] = ldraw(gp + (long long)rr * C);
                    if (rp) dd[j] = ldraw(rp + (long long)rr * C);
                    if (acc) oo[j] = ldraw(dst + (long long)rr * stride);
                }
            }
        };
        load_rows(r0 + pl, xr, gr, rr_, od);
        for (int r = r0 + pl; r < r1; r += NB * PL) {
            // software pipeline: next rows' loads are in flight while the current ones are processed
            Raw8<T> xn[NB], gn[NB], rn[NB], on[NB];
            load_rows(r + NB * PL, xn, gn, rn, on);
#pragma unroll
            for (int j = 0; j < NB; ++j) {
                const int rr = r + j * PL;
                if (rr >= r1) break;
                float x[8], du[8], o[8];
                unpack(xr[j], x);
                unpack(gr[j], du);      // the reduce pass left du in the dy buffer
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const float xh = gn_xhat<FAST>(k, e, x[e]);
                    o[e] = fmaf(xh, s2[e], fmaf(du[e], k.ag[e], s1[e]));
                }
                if (rp) {
                    float d[8];
                    unpack(rr_[j], d);
#pragma unroll
                    for (int e = 0; e < 8; ++e) o[e] += d[e];
                }
                if (acc) {
                    float old[8];
                    unpack(od[j], old);
#pragma unroll
                    for (int e = 0; e < 8; ++e) o[e] += old[e];
                }
#pragma unroll
                for (int e = 0; e < 8; ++e) cs[e] += o[e];
                st8(dst + (long long)rr * stride, o);
            }
#pragma unroll
            for (int j = 0; j < NB; ++j) {
                xr[j] = xn[j];
                gr[j] = gn[j];
                rr_[j] = rn[j];
                od[j] = on[j];
            }
        }
        return;
    }
    for (int r = r0 + pl; r < r1; r += PL) {
        float xh[8], du[8], o[8];
        gn_du8<T, FAST>(f, k, dy, n, r, c0, xh, du);
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] = fmaf(xh[e], s2[e], fmaf(du[e], k.ag[e], s1[e]));
        if (dres) {
            float d[8];
            gn_gather8<T>(dres, a.dres_resample, n, f.H, f.W, r, C, c0, d);
#pragma unroll
            for (int e = 0; e < 8; ++e) o[e] += d[e];
        }
        T* p = dst + (long long)r * stride;
        if (acc) {
            float old[8];
            ld8(p, old);
#pragma unroll
            for (int e = 0; e < 8; ++e) o[e] += old[e];
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) cs[e] += o[e];
        st8(p, o);
    }
}

// one warp per channel, lanes over the samples (a serial loop over N = 64 samples per thread made each of the 73 launches
// per step cost 16 us)
__global__ void gn_bwd_params_kernel(PuGnBwdArgs a) {
    const PuGnArgs& f = a.f;
    const int C = f.C0 + f.C1;
    const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (c >= C) return;
    double sad = 0.0, sbd = 0.0;
    for (int n = lane; n < f.N; n += 32) {
        sad += a.sums[((long long)n * C + c) * 2];
        sbd += a.sums[((long long)n * C + c) * 2 + 1];
    }
    sad = warp_sum_d(sad);
    sbd = warp_sum_d(sbd);
    if (lane != 0) return;
    const float sa = (float)sad, sb = (float)sbd;
    const float sc = f.ada ? f.ada[c] : 0.f;
    const float dg = (1.f + sc) * sb, db = (1.f + sc) * sa;
    if (a.acc_params) {
        a.dgamma[c] += dg;
        a.dbeta[c] += db;
    } else {
        a.dgamma[c] = dg;
        a.dbeta[c] = db;
    }
    if (f.ada && a.dada) {
        const float ds = f.gamma[c] * sb + f.beta[c] * sa;
        if (a.acc_params) {
            a.dada[c] += ds;
            a.dada[C + c] += sa;
        } else {
            a.dada[c] = ds;
            a.dada[C + c] = sa;
        }
    }
}

// ======================================================================================================================
// Bulk-copy (TMA) staged variants of the two backward passes, bf16, no resampling.
//
// The register-pipelined kernels above keep ~2 rows x 16 B per array and thread in flight and run at 53 - 64 % of the
// measured copy bandwidth.  Here a producer warp streams the block's pixel range through a ring of shared-memory stages
// with cp.async.bulk (1-D bulk copies: a pixel range of one sample is one contiguous run per source tensor) -- 3 - 4
// stages x 8 KB per input array, 2 blocks per SM -- and the 256 consumer threads keep the same (channel-vector,
// pixel-lane) mapping and math as above, reading 16-byte vectors from shared memory (conflict-free: a warp reads 512
// contiguous bytes) and writing results straight to global memory.  Measured (B200, batch 64): backward 0.445 -> 0.402 ms
// at 128 channels x 128x128, 0.253 -> 0.232 at 256 x 64x64, i.e. ~10 %, not the 40 % a pure bandwidth model promised:
// with 100+ instructions per 8 elements (sigmoid, mask, fp32 <-> bf16) these passes are as much issue- as HBM-bound.
// The forward apply was 15 % SLOWER this way (two blocks of 8 consumer warps hide less latency than three blocks of the
// register-pipelined kernel) and keeps the kernel above.
// ======================================================================================================================
constexpr int GS_CONSUMERS = 256;
constexpr int GS_THREADS = GS_CONSUMERS + 32;     // + one producer warp
constexpr int GS_MAX_ARRAYS = 6;

__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}

// One input array of a streamed kernel: `base` points at row r0 of the block's range, rows are row_bytes apart
// (contiguous), and the array's tile sits at `off` bytes inside every stage.
struct GsArray {
    const uint8_t* base;
    int row_bytes;
    int off;
};
struct GsPlan {
    GsArray arr[GS_MAX_ARRAYS];
    int n_arrays;
    int stage_bytes;      // multiple of 128
    int stages;
    int TR;               // rows per tile
};

// producer warp: fills stage (t % stages) with rows [t*TR, min((t+1)*TR, nrows)) of every array
__device__ __forceinline__ void gs_produce(const GsPlan& pl, int nrows, uint8_t* stage0, uint64_t* full, uint64_t* empty) {
    using namespace ptx;
    if ((threadIdx.x & 31) != 0) return;
    const int ntiles = cdiv(nrows, pl.TR);
    for (int t = 0; t < ntiles; ++t) {
        const int s = t % pl.stages;
        const int use = t / pl.stages;
        if (use > 0) mbar_wait(smem_u32(&empty[s]), (use - 1) & 1);
        const int rows = min(pl.TR, nrows - t * pl.TR);
        uint32_t total = 0;
        for (int k = 0; k < pl.n_arrays; ++k) total += (uint32_t)(rows * pl.arr[k].row_bytes);
        const uint32_t fb = smem_u32(&full[s]);
        mbar_expect_tx(fb, total);
        for (int k = 0; k < pl.n_arrays; ++k) {
            if (pl.arr[k].row_bytes == 0) continue;            // unused slot
            bulk_g2s(smem_u32(stage0 + (size_t)s * pl.stage_bytes + pl.arr[k].off),
                     pl.arr[k].base + (size_t)t * pl.TR * pl.arr[k].row_bytes, (uint32_t)(rows * pl.arr[k].row_bytes), fb);
        }
    }
}

__device__ __forceinline__ Raw8<__nv_bfloat16> lds_raw(const uint8_t* p) {
    Raw8<__nv_bfloat16> r;
    r.v = *reinterpret_cast<const uint4*>(p);
    return r;
}

// shared-memory carve-up common to the three kernels: [head bytes][barriers 2*stages][pad to 128][stages x stage_bytes]
__device__ __forceinline__ void gs_carve(uint8_t* smem, int head_bytes, int stages, uint64_t*& full, uint64_t*& empty,
                                         uint8_t*& stage0) {
    full = reinterpret_cast<uint64_t*>(smem + ((head_bytes + 7) & ~7));
    empty = full + stages;
    const uint32_t bar_end = (uint32_t)(((head_bytes + 7) & ~7) + 2 * stages * 8);
    stage0 = smem + ((bar_end + 127u) & ~127u);
}
__host__ __device__ inline int gs_smem_bytes(int head_bytes, int stages, int stage_bytes) {
    return (((head_bytes + 7) & ~7) + 2 * stages * 8 + 127) / 128 * 128 + stages * stage_bytes + 128;
}

// ---- backward, first pass: du (written over dy) and the per-(sample, channel) sums ----
__global__ void __launch_bounds__(GS_THREADS, 2) gn_bwd_reduce_tma_kernel(PuGnBwdArgs a, int rows_per_block, GsPlan pl) {
    using namespace ptx;
    extern __shared__ uint8_t gs_raw[];
    uint8_t* smem = gs_raw + ((128u - (smem_u32(gs_raw) & 127u)) & 127u);
    const PuGnArgs& f = a.f;
    const int C = f.C0 + f.C1, nvec = C / 8;
    float* sm = reinterpret_cast<float*>(smem);                 // [C][2] block partial sums
    uint64_t *full, *empty;
    uint8_t* stage0;
    gs_carve(smem, 2 * C * 4, pl.stages, full, empty, stage0);
    const int n = blockIdx.y;
    const int HW = f.H * f.W;
    const int r0 = blockIdx.x * rows_per_block;
    const int nrows = min(rows_per_block, HW - r0);
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) sm[i] = 0.f;
    if (threadIdx.x == 0) {
        for (int s = 0; s < pl.stages; ++s) {
            mbar_init(smem_u32(&full[s]), 1);
            mbar_init(smem_u32(&empty[s]), GS_CONSUMERS / 32);
        }
        fence_barrier_init();
    }
    __syncthreads();
    const long long in_base = (long long)n * HW + r0;
    if (threadIdx.x >= GS_CONSUMERS) {
        pl.arr[0].base = reinterpret_cast<const uint8_t*>(f.src0) + in_base * f.C0 * 2;
        pl.arr[1].base = reinterpret_cast<const uint8_t*>(a.dy) + in_base * C * 2;
        if (f.C1 > 0) pl.arr[2].base = reinterpret_cast<const uint8_t*>(f.src1) + in_base * f.C1 * 2;
        gs_produce(pl, nrows, stage0, full, empty);
    } else {
        const int PL = GS_CONSUMERS / nvec;
        const int v = threadIdx.x % nvec, plane = threadIdx.x / nvec;
        const bool active = plane < PL;
        const int c0 = v * 8;
        ChanConst k;
        if (active) gn_load_consts(f, n, c0, k);
        float A[8], B[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) A[e] = B[e] = 0.f;
        __nv_bfloat16* du_out = reinterpret_cast<__nv_bfloat16*>(const_cast<void*>(a.dy)) + in_base * C + c0;
        const int xoff = (c0 < f.C0) ? pl.arr[0].off + c0 * 2 : pl.arr[2].off + (c0 - f.C0) * 2;
        const int xrow = (c0 < f.C0) ? f.C0 * 2 : f.C1 * 2;
        const int goff = pl.arr[1].off + c0 * 2;
        const int ntiles = cdiv(nrows, pl.TR);
        for (int t = 0; t < ntiles; ++t) {
            const int s = t % pl.stages;
            mbar_wait(smem_u32(&full[s]), (t / pl.stages) & 1);
            const uint8_t* st = stage0 + (size_t)s * pl.stage_bytes;
            const int rows = min(pl.TR, nrows - t * pl.TR);
            if (active) {
                for (int rr = plane; rr < rows; rr += PL) {
                    float x[8], g[8], xh[8], du[8];
                    unpack(lds_raw(st + xoff + rr * xrow), x);
                    unpack(lds_raw(st + goff + rr * C * 2), g);
                    const int row = t * pl.TR + rr;
                    gn_du8_calc<true>(f, k, x, g, in_base + row, c0, xh, du);
                    st8(du_out + (long long)row * C, du);
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        A[e] += du[e];
                        B[e] = fmaf(du[e], xh[e], B[e]);
                    }
                }
            }
            __syncwarp();
            if ((threadIdx.x & 31) == 0) mbar_arrive(smem_u32(&empty[s]));
        }
        if (active) {
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                atomicAdd(&sm[2 * (c0 + e)], A[e]);
                atomicAdd(&sm[2 * (c0 + e) + 1], B[e]);
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) atomicAdd(a.sums + (long long)n * C * 2 + i, (double)sm[i]);
}

// ---- backward, second pass: dx = du * ag + s1 + xhat * s2 (+ dres) (+ old dx) ----
__global__ void __launch_bounds__(GS_THREADS, 2) gn_bwd_apply_tma_kernel(PuGnBwdArgs a, int rows_per_block, GsPlan pl) {
    using namespace ptx;
    extern __shared__ uint8_t gs_raw[];
    uint8_t* smem = gs_raw + ((128u - (smem_u32(gs_raw) & 127u)) & 127u);
    const PuGnArgs& f = a.f;
    const int C = f.C0 + f.C1, nvec = C / 8, Cg = C / f.G;
    double* smd = reinterpret_cast<double*>(smem);              // [G][2]: sum_c gamma' A, sum_c gamma' B
    float* csm = reinterpret_cast<float*>(smd + 2 * f.G);       // [C] column sums of the written gradient
    uint64_t *full, *empty;
    uint8_t* stage0;
    gs_carve(smem, 2 * f.G * 8 + C * 4, pl.stages, full, empty, stage0);
    const int n = blockIdx.y;
    const int HW = f.H * f.W;
    const int r0 = blockIdx.x * rows_per_block;
    const int nrows = min(rows_per_block, HW - r0);
    const bool want_cs = a.colsum0 != nullptr;
    for (int i = threadIdx.x; i < 2 * f.G; i += blockDim.x) smd[i] = 0.0;
    for (int i = threadIdx.x; i < C; i += blockDim.x) csm[i] = 0.f;
    if (threadIdx.x == 0) {
        for (int s = 0; s < pl.stages; ++s) {
            mbar_init(smem_u32(&full[s]), 1);
            mbar_init(smem_u32(&empty[s]), GS_CONSUMERS / 32);
        }
        fence_barrier_init();
    }
    __syncthreads();
    const long long in_base = (long long)n * HW + r0;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float gam = f.gamma[c];
        if (f.ada) gam *= 1.f + f.ada[c];
        const int g = c / Cg;
        atomicAdd(&smd[2 * g], (double)gam * a.sums[((long long)n * C + c) * 2]);
        atomicAdd(&smd[2 * g + 1], (double)gam * a.sums[((long long)n * C + c) * 2 + 1]);
    }
    __syncthreads();        // (before the role split: the producer loop blocks on the consumers' progress)
    // array slots: 0 x0, 1 du, 2 x1, 3 dres, 4 old dx0, 5 old dx1 (unused slots have row_bytes 0 and are skipped)
    if (threadIdx.x >= GS_CONSUMERS) {
        pl.arr[0].base = reinterpret_cast<const uint8_t*>(f.src0) + in_base * f.C0 * 2;
        pl.arr[1].base = reinterpret_cast<const uint8_t*>(a.dy) + in_base * C * 2;
        pl.arr[2].base = reinterpret_cast<const uint8_t*>(f.src1) + in_base * f.C1 * 2;
        pl.arr[3].base = reinterpret_cast<const uint8_t*>(a.dres) + in_base * C * 2;
        pl.arr[4].base = reinterpret_cast<const uint8_t*>(a.dx0) + in_base * f.C0 * 2;
        pl.arr[5].base = reinterpret_cast<const uint8_t*>(a.dx1) + in_base * f.C1 * 2;
        gs_produce(pl, nrows, stage0, full, empty);
    } else {
        const int PL = GS_CONSUMERS / nvec;
        const int v = threadIdx.x % nvec, plane = threadIdx.x / nvec;
        const bool active = plane < PL;
        const int c0 = v * 8;
        ChanConst k;
        float s1[8], s2[8], cs[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) s1[e] = s2[e] = cs[e] = 0.f;
        if (active) {
            gn_load_consts(f, n, c0, k);
            const double inv_m = 1.0 / ((double)Cg * (double)f.H * (double)f.W);
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const int g = (c0 + e) / Cg;
                s1[e] = -k.rstd[e] * (float)(smd[2 * g] * inv_m);
                s2[e] = -k.rstd[e] * (float)(smd[2 * g + 1] * inv_m);
            }
        }
        const bool lo = c0 < f.C0;
        const int xoff = lo ? pl.arr[0].off + c0 * 2 : pl.arr[2].off + (c0 - f.C0) * 2;
        const int xrow = lo ? f.C0 * 2 : f.C1 * 2;
        const int goff = pl.arr[1].off + c0 * 2;
        const bool has_res = pl.arr[3].row_bytes != 0;
        const int roff = pl.arr[3].off + c0 * 2;
        const bool acc = lo ? (pl.arr[4].row_bytes != 0) : (pl.arr[5].row_bytes != 0);
        const int ooff = lo ? pl.arr[4].off + c0 * 2 : pl.arr[5].off + (c0 - f.C0) * 2;
        __nv_bfloat16* dst = lo ? reinterpret_cast<__nv_bfloat16*>(a.dx0) + in_base * f.C0 + c0
                                : reinterpret_cast<__nv_bfloat16*>(a.dx1) + in_base * f.C1 + (c0 - f.C0);
        const int dstride = lo ? f.C0 : f.C1;
        const int ntiles = cdiv(nrows, pl.TR);
        for (int t = 0; t < ntiles; ++t) {
            const int s = t % pl.stages;
            mbar_wait(smem_u32(&full[s]), (t / pl.stages) & 1);
            const uint8_t* st = stage0 + (size_t)s * pl.stage_bytes;
            const int rows = min(pl.TR, nrows - t * pl.TR);
            if (active) {
                for (int rr = plane; rr < rows; rr += PL) {
                    float x[8], du[8], o[8];
                    unpack(lds_raw(st + xoff + rr * xrow), x);
                    unpack(lds_raw(st + goff + rr * C * 2), du);
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const float xh = gn_xhat<true>(k, e, x[e]);
                        o[e] = fmaf(xh, s2[e], fmaf(du[e], k.ag[e], s1[e]));
                    }
                    if (has_res) {
                        float d[8];
                        unpack(lds_raw(st + roff + rr * C * 2), d);
#pragma unroll
                        for (int e = 0; e < 8; ++e) o[e] += d[e];
                    }
                    if (acc) {
                        float old[8];
                        unpack(lds_raw(st + ooff + rr * xrow), old);
#pragma unroll
                        for (int e = 0; e < 8; ++e) o[e] += old[e];
                    }
#pragma unroll
                    for (int e = 0; e < 8; ++e) cs[e] += o[e];
                    st8(dst + (long long)(t * pl.TR + rr) * dstride, o);
                }
            }
            __syncwarp();
            if ((threadIdx.x & 31) == 0) mbar_arrive(smem_u32(&empty[s]));
        }
        if (want_cs && active) {
#pragma unroll
            for (int e = 0; e < 8; ++e) atomicAdd(&csm[c0 + e], cs[e]);
        }
    }
    __syncthreads();
    if (want_cs) {
        for (int c = threadIdx.x; c < C; c += blockDim.x) {
            if (c < f.C0)
                atomicAdd(a.colsum0 + c, csm[c]);
            else if (a.colsum1)
                atomicAdd(a.colsum1 + (c - f.C0), csm[c]);
        }
    }
}

// host: tile plan of a streamed launch.  row_bytes[i] == 0 marks an unused slot.
static GsPlan gs_make_plan(int C, const int* row_bytes, int n_slots, int stages) {
    GsPlan pl;
    const int nvec = C / 8;
    const int PL = GS_CONSUMERS / nvec;
    pl.TR = 2 * (PL > 0 ? PL : 1);
    pl.stages = stages;
    pl.n_arrays = n_slots;
    int off = 0;
    for (int i = 0; i < GS_MAX_ARRAYS; ++i) {
        pl.arr[i].base = nullptr;
        pl.arr[i].row_bytes = i < n_slots ? row_bytes[i] : 0;
        pl.arr[i].off = off;
        off += ((pl.arr[i].row_bytes * pl.TR + 127) / 128) * 128;
    }
    pl.stage_bytes = off;
    return pl;
}
static bool gs_enabled() {
    static const bool on = !(getenv("PU_GN_TMA") && getenv("PU_GN_TMA")[0] == '0');
    return on;
}

// the per-(sample, channel) constants of the conv epilogue that takes over the reduce pass (PuConvGnBwd.consts)
__global__ void gn_bwd_consts_kernel(PuGnArgs f, float4* __restrict__ out) {
    const int C = f.C0 + f.C1;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= f.N * C) return;
    const int n = i / C, c = i - n * C;
    const int Cg = C / f.G, g = c / Cg;
    const double m = (double)Cg * f.H * f.W;
    const double mm = f.stats[((long long)n * f.G + g) * 2] / m;
    double var = f.stats[((long long)n * f.G + g) * 2 + 1] / m - mm * mm;
    if (var < 0) var = 0;
    const float mean = (float)mm;
    const float rstd = (float)(1.0 / sqrt(var + (double)f.eps));
    float gam = f.gamma[c], bet = f.beta[c];
    if (f.ada) {
        const float sc = f.ada[c], sh = f.ada[C + c];
        gam = gam * (1.f + sc);
        bet = fmaf(bet, 1.f + sc, sh);
    }
    out[i] = make_float4(rstd * gam, fmaf(-mean * rstd, gam, bet), rstd, -mean * rstd);
}

// pixel rows per block: aim at ~8 blocks per SM over the whole launch, but at least `min_rows` rows of work
// The grid is (chunks per sample, N) with 3 resident blocks per SM (launch bounds): choose the chunk count (up to ~3
// waves) whose last wave is fullest -- e.g. N = 64: 13 chunks -> 832 blocks = 1.87 waves of 444 instead of
// 18 chunks -> 1152 blocks = 2.6 waves.
static int rows_per_block(int HW, int N, int min_rows, int blocks_per_sm = 3) {
    const int per_wave = 148 * blocks_per_sm;
    int max_chunks = cdiv(HW, min_rows);
    if (max_chunks < 1) max_chunks = 1;
    int hi = (3 * per_wave) / (N > 0 ? N : 1);
    if (hi < 1) hi = 1;
    if (hi > max_chunks) hi = max_chunks;
    int best_c = 1;
    double best = -1.0;
    for (int c = 1; c <= hi; ++c) {
        const long long blocks = (long long)c * N;
        const long long waves = cdivll(blocks, per_wave);
        const double eff = (double)blocks / (double)(waves * per_wave);
        if (eff > best + 1e-9 || (eff > best - 0.02 && blocks <= 2 * per_wave)) {   // prefer more, smaller blocks
            if (eff > best) best = eff;
            best_c = c;
        }
    }
    int rows = cdiv(HW, best_c);
    if (rows < min_rows) rows = min_rows;
    return rows;
}

static int check_gn(const PuGnArgs& f, const char* who) {
    const int C = f.C0 + f.C1;
    PU_REQUIRE(f.src0 && f.stats && f.gamma && f.beta, "%s: null pointer", who);
    PU_REQUIRE(f.N > 0 && f.H > 0 && f.W > 0 && f.C0 > 0 && f.C1 >= 0 && f.G > 0, "%s: bad shape", who);
    PU_REQUIRE(f.C0 % 8 == 0 && f.C1 % 8 == 0 && C % f.G == 0 && C / 8 <= GN_THREADS,
               "%s: channels (%d,%d) must be multiples of 8, divisible by G=%d, and <= %d", who, f.C0, f.C1, f.G,
               GN_THREADS * 8);
    PU_REQUIRE(f.C1 == 0 || f.src1, "%s: C1 > 0 needs src1", who);
    PU_REQUIRE(f.resample != PU_RS_DOWN || (f.H % 2 == 0 && f.W % 2 == 0), "%s: downsample needs even H, W", who);
    PU_REQUIRE(f.dropout_p == 0.f || f.resample == PU_RS_NONE, "%s: dropout with resample is not supported", who);
    PU_REQUIRE(f.dropout_p >= 0.f && f.dropout_p < 1.f, "%s: bad dropout p", who);
    PU_REQUIRE(f.dtype == PU_F32 || f.dtype == PU_BF16, "%s: bad dtype", who);
    return PU_OK;
}

}  // namespace pu

extern "C" {

int pu_gn_stats(const void* src0, const void* src1, int C0, int C1, int N, int HW, int G, int dtype, double* stats,
                void* stream) {
    using namespace pu;
    const int C = C0 + C1;
    PU_REQUIRE(src0 && stats && N > 0 && HW > 0 && C0 > 0 && C1 >= 0 && G > 0, "pu_gn_stats: bad arguments");
    PU_REQUIRE(C0 % 8 == 0 && C1 % 8 == 0 && C % G == 0 && C / 8 <= GN_THREADS, "pu_gn_stats: unsupported channels");
    PU_REQUIRE(C1 == 0 || src1, "pu_gn_stats: C1 > 0 needs src1");
    cudaStream_t st = (cudaStream_t)stream;
    PU_CUDA(cudaMemsetAsync(stats, 0, sizeof(double) * 2 * N * G, st));
    const int PL = GN_THREADS / (C / 8);
    const int rows = rows_per_block(HW, N, PL * 8);
    dim3 grid(cdiv(HW, rows), N);
    const size_t smem = sizeof(double) * 2 * G;
    if (dtype == PU_F32)
        gn_stats_kernel<float><<<grid, GN_THREADS, smem, st>>>((const float*)src0, (const float*)src1, C0, C1, HW, G, rows,
                                                               stats);
    else
        gn_stats_kernel<__nv_bfloat16><<<grid, GN_THREADS, smem, st>>>((const __nv_bfloat16*)src0,
                                                                       (const __nv_bfloat16*)src1, C0, C1, HW, G, rows,
                                                                       stats);
    return check_launch("gn_stats");
}

int pu_gn_stats_from_quads(const double* q0, const double* q1, int C0, int C1, int N, int G, double* stats, void* stream) {
    using namespace pu;
    PU_REQUIRE(q0 && stats && N > 0 && G > 0 && C0 > 0 && C1 >= 0, "pu_gn_stats_from_quads: bad arguments");
    PU_REQUIRE(C1 == 0 || q1, "pu_gn_stats_from_quads: C1 > 0 needs q1");
    PU_REQUIRE(C0 % 4 == 0 && C1 % 4 == 0 && (C0 + C1) % G == 0 && ((C0 + C1) / G) % 4 == 0,
               "pu_gn_stats_from_quads: channels (%d,%d) / groups %d must give group sizes that are multiples of 4", C0, C1, G);
    gn_stats_from_quads_kernel<<<cdiv(N * G, 128), 128, 0, (cudaStream_t)stream>>>(q0, q1, C0, C1, N, G, stats);
    return check_launch("gn_stats_from_quads");
}

int pu_gn_apply(const PuGnArgs* a, void* stream) {
    using namespace pu;
    PU_REQUIRE(a && a->y, "pu_gn_apply: null pointer");
    int rc = check_gn(*a, "pu_gn_apply");
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const int C = a->C0 + a->C1;
    const int OHW = a->resample == PU_RS_UP ? a->H * a->W * 4 : (a->resample == PU_RS_DOWN ? a->H * a->W / 4 : a->H * a->W);
    const int PL = GN_THREADS / (C / 8);
    const int rows = rows_per_block(OHW, a->N, PL * 8);
    dim3 grid(cdiv(OHW, rows), a->N);
    if (a->dtype == PU_F32)
        gn_apply_kernel<float, false><<<grid, GN_THREADS, 0, st>>>(*a, rows);
    else
        gn_apply_kernel<__nv_bfloat16, true><<<grid, GN_THREADS, 0, st>>>(*a, rows);
    return check_launch("gn_apply");
}

int pu_gn_bwd_consts(const PuGnArgs* f, float* consts, void* stream) {
    using namespace pu;
    PU_REQUIRE(f && consts, "pu_gn_bwd_consts: null pointer");
    int rc = check_gn(*f, "pu_gn_bwd_consts");
    if (rc) return rc;
    const int total = f->N * (f->C0 + f->C1);
    gn_bwd_consts_kernel<<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(*f, reinterpret_cast<float4*>(consts));
    return check_launch("gn_bwd_consts");
}

int pu_gn_bwd(const PuGnBwdArgs* a, void* stream) {
    using namespace pu;
    PU_REQUIRE(a && a->dy && a->sums && a->dx0 && a->dgamma && a->dbeta, "pu_gn_bwd: null pointer");
    int rc = check_gn(a->f, "pu_gn_bwd");
    if (rc) return rc;
    const PuGnArgs& f = a->f;
    PU_REQUIRE(f.C1 == 0 || a->dx1, "pu_gn_bwd: C1 > 0 needs dx1");
    PU_REQUIRE(!a->du_ready || (f.resample == PU_RS_NONE && (!a->dres || a->dres_resample == PU_RS_NONE)),
               "pu_gn_bwd: du_ready needs resample == dres_resample == PU_RS_NONE");
    cudaStream_t st = (cudaStream_t)stream;
    const int C = f.C0 + f.C1;
    const int HW = f.H * f.W;
    const int PL = GN_THREADS / (C / 8);
    const int rows = rows_per_block(HW, f.N, PL * 8, GN_BWD_BLOCKS);
    dim3 grid(cdiv(HW, rows), f.N);
    const bool streamed = f.dtype == PU_BF16 && f.resample == PU_RS_NONE && (!a->dres || a->dres_resample == PU_RS_NONE) &&
                          gs_enabled();
    if (streamed) {
        // bulk-copy staged kernels (see gn_bwd_reduce_tma_kernel): same math, shared-memory ring fed by a producer warp
        if (!a->du_ready) {
            PU_CUDA(cudaMemsetAsync(a->sums, 0, sizeof(double) * 2 * f.N * C, st));
            const int rb[3] = {f.C0 * 2, C * 2, f.C1 * 2};
            GsPlan pl = gs_make_plan(C, rb, 3, 4);
            const int rows1 = rows_per_block(HW, f.N, pl.TR * 4, 2);
            const int smem = gs_smem_bytes(2 * C * 4, pl.stages, pl.stage_bytes);
            PU_SMEM_ATTR(gn_bwd_reduce_tma_kernel, 110 * 1024);
            PU_REQUIRE(smem <= 110 * 1024, "pu_gn_bwd: reduce tile plan needs %d bytes of shared memory", smem);
            dim3 g1(cdiv(HW, rows1), f.N);
            gn_bwd_reduce_tma_kernel<<<g1, GS_THREADS, smem, st>>>(*a, rows1, pl);
            rc = check_launch("gn_bwd_reduce_tma");
            if (rc) return rc;
        }
        if (a->colsum0) PU_CUDA(cudaMemsetAsync(a->colsum0, 0, sizeof(float) * f.C0, st));
        if (a->colsum1 && f.C1 > 0) PU_CUDA(cudaMemsetAsync(a->colsum1, 0, sizeof(float) * f.C1, st));
        const int rb[6] = {f.C0 * 2, C * 2, f.C1 * 2, a->dres ? C * 2 : 0, a->acc0 ? f.C0 * 2 : 0,
                           (a->acc1 && f.C1 > 0) ? f.C1 * 2 : 0};
        GsPlan pl = gs_make_plan(C, rb, 6, 3);
        const int rows2 = rows_per_block(HW, f.N, pl.TR * 4, 2);
        const int smem = gs_smem_bytes(2 * f.G * 8 + C * 4, pl.stages, pl.stage_bytes);
        PU_SMEM_ATTR(gn_bwd_apply_tma_kernel, 110 * 1024);
        PU_REQUIRE(smem <= 110 * 1024, "pu_gn_bwd: apply tile plan needs %d bytes of shared memory", smem);
        dim3 g2(cdiv(HW, rows2), f.N);
        gn_bwd_apply_tma_kernel<<<g2, GS_THREADS, smem, st>>>(*a, rows2, pl);
        rc = check_launch("gn_bwd_apply_tma");
        if (rc) return rc;
        gn_bwd_params_kernel<<<cdiv(C, 4), 128, 0, st>>>(*a);
        return check_launch("gn_bwd_params");
    }
    if (!a->du_ready) {
        PU_CUDA(cudaMemsetAsync(a->sums, 0, sizeof(double) * 2 * f.N * C, st));
        const size_t smem = sizeof(float) * 2 * C;
        if (f.dtype == PU_F32)
            gn_bwd_reduce_kernel<float, false><<<grid, GN_THREADS, smem, st>>>(*a, rows);
        else
            gn_bwd_reduce_kernel<__nv_bfloat16, true><<<grid, GN_THREADS, smem, st>>>(*a, rows);
        rc = check_launch("gn_bwd_reduce");
        if (rc) return rc;
    }
    {
        if (a->colsum0) PU_CUDA(cudaMemsetAsync(a->colsum0, 0, sizeof(float) * f.C0, st));
        if (a->colsum1 && f.C1 > 0) PU_CUDA(cudaMemsetAsync(a->colsum1, 0, sizeof(float) * f.C1, st));
        const size_t smem = sizeof(double) * 2 * f.G + sizeof(float) * C;
        if (f.dtype == PU_F32)
            gn_bwd_apply_kernel<float, false><<<grid, GN_THREADS, smem, st>>>(*a, rows);
        else
            gn_bwd_apply_kernel<__nv_bfloat16, true><<<grid, GN_THREADS, smem, st>>>(*a, rows);
        rc = check_launch("gn_bwd_apply");
        if (rc) return rc;
    }
    gn_bwd_params_kernel<<<cdiv(C, 4), 128, 0, st>>>(*a);
    return check_launch("gn_bwd_params");
}
}
