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
    extern __shared__ float sm[];   // [G][2]
    const int C = C0 + C1, nvec = C / 8, Cg = C / G;
    const int n = blockIdx.y;
    const int PL = GN_THREADS / nvec;
    const int v = threadIdx.x % nvec, pl = threadIdx.x / nvec;
    for (int i = threadIdx.x; i < 2 * G; i += blockDim.x) sm[i] = 0.f;
    __syncthreads();
    const int r0 = blockIdx.x * rows;
    int r1 = r0 + rows;
    if (r1 > HW) r1 = HW;
    if (pl < PL) {
        float s[8], q[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) s[e] = q[e] = 0.f;
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
#pragma unroll
            for (int j = 0; j < SNB; ++j) {
                if (r + j * PL >= r1) break;
                float x[8];
                unpack(raw[j], x);
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    s[e] += x[e];
                    q[e] = fmaf(x[e], x[e], q[e]);
                }
            }
#pragma unroll
            for (int j = 0; j < SNB; ++j) raw[j] = nxt[j];
        }
        // combine channels of the same group before touching shared memory
        int g = c0 / Cg;
        float gs = 0.f, gq = 0.f;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int ge = (c0 + e) / Cg;
            if (ge != g) {
                atomicAdd(&sm[2 * g], gs);
                atomicAdd(&sm[2 * g + 1], gq);
                g = ge;
                gs = gq = 0.f;
            }
            gs += s[e];
            gq += q[e];
        }
        atomicAdd(&sm[2 * g], gs);
        atomicAdd(&sm[2 * g + 1], gq);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * G; i += blockDim.x) atomicAdd(stats + (long long)n * G * 2 + i, (double)sm[i]);
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
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const float u = gn_u<FAST>(k, e, x[e]);
                    o[e] = f.silu ? u * sigmoid_t<FAST>(u) : u;
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
        keep = dropout_keep8(f.seed, (unsigned long long)((pix * C + c0) >> 3), f.dropout_p);
        inv_keep = 1.f / (1.f - f.dropout_p);
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

__global__ void gn_bwd_params_kernel(PuGnBwdArgs a) {
    const PuGnArgs& f = a.f;
    const int C = f.C0 + f.C1;
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    double sad = 0.0, sbd = 0.0;
    for (int n = 0; n < f.N; ++n) {
        sad += a.sums[((long long)n * C + c) * 2];
        sbd += a.sums[((long long)n * C + c) * 2 + 1];
    }
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
    const size_t smem = sizeof(float) * 2 * G;
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
    gn_bwd_params_kernel<<<cdiv(C, 128), 128, 0, st>>>(*a);
    return check_launch("gn_bwd_params");
}
}
