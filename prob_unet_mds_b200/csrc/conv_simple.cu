// CUDA-core implicit-GEMM convolution (fp32 accumulate).  This is the fp32-mode path (1e-5 parity against
// the reference) and the path for the few layers whose channel counts do not fit the tcgen05 tiles
// (first convs with Cin = 3 / 6).  Replaces F.conv2d at networks.py:87 and nn.Conv2d at prob_unet.py:33.
#include "../../include/probunet_b200.h"
#include "common.cuh"
#include "conv_internal.h"

namespace pu {

constexpr int TS = 64;   // tile size (M and N)
constexpr int TK = 16;   // k chunk

template <typename T>
__global__ void __launch_bounds__(256) conv_simple_kernel(PuConvArgs a) {
    __shared__ float As[TK][TS + 4];
    __shared__ float Bs[TK][TS + 4];
    const int tid = threadIdx.x;
    const int tx = tid % 16, ty = tid / 16;
    const int Ctot = a.C0 + a.C1;
    const int taps = a.ksize * a.ksize;
    const int K = taps * Ctot;
    const long long M = (long long)a.N * a.H * a.W;
    const long long m0 = (long long)blockIdx.x * TS;
    const int n0 = blockIdx.y * TS;
    const T* s0 = reinterpret_cast<const T*>(a.src0);
    const T* s1 = reinterpret_cast<const T*>(a.src1);
    const T* wt = reinterpret_cast<const T*>(a.weight);

    // loader coordinates: each thread loads 4 consecutive k for one row
    const int lrow = tid / 4;
    const int lk = (tid % 4) * 4;
    const long long lm = m0 + lrow;
    int ln = 0, ly = 0, lx = 0;
    const bool lm_ok = lm < M;
    if (lm_ok) {
        ln = (int)(lm / (a.H * a.W));
        int r = (int)(lm % (a.H * a.W));
        ly = r / a.W;
        lx = r % a.W;
    }
    const int pad = a.ksize / 2;

    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int k0 = 0; k0 < K; k0 += TK) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            int kk = k0 + lk + i;
            float v = 0.f;
            if (lm_ok && kk < K) {
                int tap = kk / Ctot;
                int c = kk - tap * Ctot;
                int iy = ly + tap / a.ksize - pad;
                int ix = lx + tap % a.ksize - pad;
                if (iy >= 0 && iy < a.H && ix >= 0 && ix < a.W) {
                    long long pix = ((long long)ln * a.H + iy) * a.W + ix;
                    v = (c < a.C0) ? ldf(s0 + pix * a.C0 + c) : ldf(s1 + pix * a.C1 + (c - a.C0));
                }
            }
            As[lk + i][lrow] = v;
            float w = 0.f;
            int n = n0 + lrow;
            if (n < a.Cout && kk < K) w = ldf(wt + (long long)n * K + kk);
            Bs[lk + i][lrow] = w;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < TK; ++k) {
            float av[4], bv[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) av[i] = As[k][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) bv[j] = Bs[k][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }

    T* out = reinterpret_cast<T*>(a.out);
    const T* res = reinterpret_cast<const T*>(a.residual);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        long long m = m0 + ty * 4 + i;
        if (m >= M) continue;
        int img = (int)(m / (a.H * a.W));
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int n = n0 + tx * 4 + j;
            if (n >= a.Cout) continue;
            float v = acc[i][j];
            if (a.bias) v += a.bias_per_sample ? a.bias[(long long)img * a.Cout + n] : a.bias[n];
            if (res) v += ldf(res + m * a.Cout + n);
            if (a.flags & PU_CONV_RELU) v = relu_f(v);
            stf(out + m * a.Cout + n, v);
        }
    }
}

// dw[co][tap][c] += sum_p dy[p][co] * src[p + tap][c]
template <typename T>
__global__ void __launch_bounds__(256) wgrad_simple_kernel(PuWgradArgs a, int chunk) {
    __shared__ float As[TK][TS + 4];   // [k = pixel][m = co]
    __shared__ float Bs[TK][TS + 4];   // [k = pixel][n = (tap, c)]
    const int tid = threadIdx.x;
    const int tx = tid % 16, ty = tid / 16;
    const int Ctot = a.C0 + a.C1;
    const int taps = a.ksize * a.ksize;
    const int NN = taps * Ctot;
    const long long P = (long long)a.N * a.H * a.W;
    const int m0 = blockIdx.x * TS;
    const int n0 = blockIdx.y * TS;
    const long long p_begin = (long long)blockIdx.z * chunk;
    long long p_end = p_begin + chunk;
    if (p_end > P) p_end = P;
    const T* s0 = reinterpret_cast<const T*>(a.src0);
    const T* s1 = reinterpret_cast<const T*>(a.src1);
    const T* dy = reinterpret_cast<const T*>(a.dy);
    const int pad = a.ksize / 2;

    const int lk = tid / 16;          // pixel within chunk of 16
    const int lc = (tid % 16) * 4;    // 4 consecutive m / n

    // per-thread decode of the 4 n indices it loads
    int tap_dy[4], tap_dx[4], cc[4];
    bool nok[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int nn = n0 + lc + i;
        nok[i] = nn < NN;
        int tap = nok[i] ? nn / Ctot : 0;
        cc[i] = nok[i] ? nn - tap * Ctot : 0;
        tap_dy[i] = tap / a.ksize - pad;
        tap_dx[i] = tap % a.ksize - pad;
    }

    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (long long p0 = p_begin; p0 < p_end; p0 += TK) {
        long long p = p0 + lk;
        bool pok = p < p_end;
        int img = 0, y = 0, x = 0;
        if (pok) {
            img = (int)(p / (a.H * a.W));
            int r = (int)(p % (a.H * a.W));
            y = r / a.W;
            x = r % a.W;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            int m = m0 + lc + i;
            As[lk][lc + i] = (pok && m < a.Cout) ? ldf(dy + p * a.Cout + m) : 0.f;
            float v = 0.f;
            if (pok && nok[i]) {
                int iy = y + tap_dy[i], ix = x + tap_dx[i];
                if (iy >= 0 && iy < a.H && ix >= 0 && ix < a.W) {
                    long long pix = ((long long)img * a.H + iy) * a.W + ix;
                    int c = cc[i];
                    v = (c < a.C0) ? ldf(s0 + pix * a.C0 + c) : ldf(s1 + pix * a.C1 + (c - a.C0));
                }
            }
            Bs[lk][lc + i] = v;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < TK; ++k) {
            float av[4], bv[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) av[i] = As[k][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) bv[j] = Bs[k][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int m = m0 + ty * 4 + i;
        if (m >= a.Cout) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int nn = n0 + tx * 4 + j;
            if (nn >= NN) continue;
            atomicAdd(a.dw + (long long)m * NN + nn, acc[i][j]);
        }
    }
}

template <typename T>
__global__ void bias_grad_kernel(const T* __restrict__ dy, float* __restrict__ db, long long P, int C, int rows) {
    long long p0 = (long long)blockIdx.x * rows;
    long long p1 = p0 + rows;
    if (p1 > P) p1 = P;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float s = 0.f;
        for (long long p = p0; p < p1; ++p) s += ldf(dy + p * C + c);
        atomicAdd(db + c, s);
    }
}

// vectorised variant: thread (v, pl) sums channels [8v, 8v+8) over pixels pl, pl+PL, ... of the block's range
template <typename T>
__global__ void __launch_bounds__(256) bias_grad_vec_kernel(const T* __restrict__ dy, float* __restrict__ db, long long P,
                                                            int C, int rows) {
    extern __shared__ float sm[];   // [C]
    const int nvec = C / 8;
    const int PL = 256 / nvec;
    const int v = threadIdx.x % nvec, pl = threadIdx.x / nvec;
    for (int i = threadIdx.x; i < C; i += blockDim.x) sm[i] = 0.f;
    __syncthreads();
    long long p0 = (long long)blockIdx.x * rows;
    long long p1 = p0 + rows;
    if (p1 > P) p1 = P;
    if (pl < PL) {
        float s[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) s[e] = 0.f;
        for (long long p = p0 + pl; p < p1; p += 4 * PL) {
            float x[4][8];
#pragma unroll
            for (int j = 0; j < 4; ++j) {   // four independent 16-byte loads in flight
                const long long pp = p + (long long)j * PL;
                if (pp < p1) {
                    ld8(dy + pp * C + v * 8, x[j]);
                } else {
#pragma unroll
                    for (int e = 0; e < 8; ++e) x[j][e] = 0.f;
                }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int e = 0; e < 8; ++e) s[e] += x[j][e];
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) atomicAdd(&sm[v * 8 + e], s[e]);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < C; i += blockDim.x) atomicAdd(db + i, sm[i]);
}

int conv_simple_launch(const PuConvArgs* a, cudaStream_t st) {
    long long M = (long long)a->N * a->H * a->W;
    dim3 grid((unsigned)cdivll(M, TS), (unsigned)cdiv(a->Cout, TS));
    if (a->dtype == PU_F32)
        conv_simple_kernel<float><<<grid, 256, 0, st>>>(*a);
    else
        conv_simple_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(*a);
    return check_launch("conv_simple");
}

int wgrad_simple_launch(const PuWgradArgs* a, cudaStream_t st) {
    int Ctot = a->C0 + a->C1;
    int NN = a->ksize * a->ksize * Ctot;
    long long P = (long long)a->N * a->H * a->W;
    int gx = cdiv(a->Cout, TS), gy = cdiv(NN, TS);
    long long want = cdivll(148LL * 6, (long long)gx * gy);
    long long maxsplit = cdivll(P, 256);
    long long split = want < 1 ? 1 : (want > maxsplit ? maxsplit : want);
    int chunk = (int)cdivll(P, split);
    chunk = cdiv(chunk, TK) * TK;
    split = cdivll(P, chunk);
    dim3 grid(gx, gy, (unsigned)split);
    if (a->dtype == PU_F32)
        wgrad_simple_kernel<float><<<grid, 256, 0, st>>>(*a, chunk);
    else
        wgrad_simple_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(*a, chunk);
    return check_launch("wgrad_simple");
}

}  // namespace pu

extern "C" int pu_bias_grad(const void* dy, float* db, long long pixels, int C, int dtype, int accumulate,
                            void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    PU_REQUIRE(dy && db && pixels > 0 && C > 0, "pu_bias_grad: bad arguments");
    if (!accumulate) PU_CUDA(cudaMemsetAsync(db, 0, sizeof(float) * C, st));
    if (C % 8 == 0 && C / 8 <= 256) {
        const int PL = 256 / (C / 8);
        int rows = (int)pu::cdivll(pixels, 148 * 8);
        if (rows < PL * 8) rows = PL * 8;
        unsigned grid = (unsigned)pu::cdivll(pixels, rows);
        if (dtype == PU_F32)
            pu::bias_grad_vec_kernel<float><<<grid, 256, sizeof(float) * C, st>>>((const float*)dy, db, pixels, C, rows);
        else
            pu::bias_grad_vec_kernel<__nv_bfloat16><<<grid, 256, sizeof(float) * C, st>>>((const __nv_bfloat16*)dy, db,
                                                                                         pixels, C, rows);
        return pu::check_launch("bias_grad_vec");
    }
    int rows = (int)pu::cdivll(pixels, 148 * 8);
    if (rows < 32) rows = 32;
    unsigned grid = (unsigned)pu::cdivll(pixels, rows);
    int threads = C >= 256 ? 256 : (C >= 128 ? 128 : 64);
    if (dtype == PU_F32)
        pu::bias_grad_kernel<float><<<grid, threads, 0, st>>>((const float*)dy, db, pixels, C, rows);
    else
        pu::bias_grad_kernel<__nv_bfloat16><<<grid, threads, 0, st>>>((const __nv_bfloat16*)dy, db, pixels, C, rows);
    return pu::check_launch("bias_grad");
}
