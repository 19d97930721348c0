// Prior/posterior encoder glue and the latent kernels: 2x2 average pool, ReLU/pool backward, global mean,
// mu/log_sigma heads, reparameterised sample, analytic KL, sum-MSE.
// Reference: prob_unet.py:32-36,60-77 (AxisAlignedConvGaussian), :188,:193,:221 (rsample), :227 (MSELoss sum),
// :230 (kl_divergence), torch.distributions.kl._kl_normal_normal.
#include "../../include/probunet_b200.h"
#include "common.cuh"

namespace pu {

template <typename T>
__global__ void avgpool2_kernel(const T* __restrict__ x, T* __restrict__ y, int N, int H, int W, int C) {
    const int OH = H / 2, OW = W / 2, nvec = C / 8;
    const long long total = (long long)N * OH * OW * nvec;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int v = (int)(i % nvec);
        long long t = i / nvec;
        const int ox = (int)(t % OW);
        t /= OW;
        const int oy = (int)(t % OH);
        const int n = (int)(t / OH);
        float o[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] = 0.f;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float a[8];
            ld8(x + (((long long)n * H + 2 * oy + (k >> 1)) * W + 2 * ox + (k & 1)) * C + v * 8, a);
#pragma unroll
            for (int e = 0; e < 8; ++e) o[e] += a[e];
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] *= 0.25f;
        st8(y + (((long long)n * OH + oy) * OW + ox) * C + v * 8, o);
    }
}

template <typename T>
__global__ void upsample2_kernel(const T* __restrict__ x, T* __restrict__ y, int N, int H, int W, int C) {
    const int OH = H * 2, OW = W * 2, nvec = C / 8;
    const long long total = (long long)N * OH * OW * nvec;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int v = (int)(i % nvec);
        long long t = i / nvec;
        const int ox = (int)(t % OW);
        t /= OW;
        const int oy = (int)(t % OH);
        const int n = (int)(t / OH);
        const uint4* src = reinterpret_cast<const uint4*>(x + (((long long)n * H + (oy >> 1)) * W + (ox >> 1)) * C + v * 8);
        uint4* dst = reinterpret_cast<uint4*>(y + (((long long)n * OH + oy) * OW + ox) * C + v * 8);
        if (sizeof(T) == 2) {
            dst[0] = src[0];
        } else {
            dst[0] = src[0];
            dst[1] = src[1];
        }
    }
}

// transpose of the 2x resampling that sits between a GroupNorm(+SiLU) and the conv it feeds (networks.py:82-87): the
// gradient wrt the PRE-resample activation [N,H,W,C] from the gradient dy wrt the resampled one.
//   UP   (forward = nearest x2, dy is [N,2H,2W,C]):   g[y][x] = sum of the four children of dy
//   DOWN (forward = 2x2 mean,  dy is [N,H/2,W/2,C]):  g[y][x] = 0.25 * dy[y/2][x/2]
template <typename T, bool UP>
__global__ void resample_grad_kernel(const T* __restrict__ dy, T* __restrict__ g, int N, int H, int W, int C) {
    const int nvec = C / 8;
    const long long total = (long long)N * H * W * nvec;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int v = (int)(i % nvec);
        long long t = i / nvec;
        const int x = (int)(t % W);
        t /= W;
        const int y = (int)(t % H);
        const int n = (int)(t / H);
        float o[8];
        if (UP) {
            const int OH = 2 * H, OW = 2 * W;
#pragma unroll
            for (int e = 0; e < 8; ++e) o[e] = 0.f;
#pragma unroll
            for (int k = 0; k < 4; ++k) {       // same order as the gather it replaces (gn_gather8): bit-identical in fp32
                float a[8];
                ld8(dy + (((long long)n * OH + 2 * y + (k >> 1)) * OW + 2 * x + (k & 1)) * C + v * 8, a);
#pragma unroll
                for (int e = 0; e < 8; ++e) o[e] += a[e];
            }
        } else {
            const int OH = H / 2, OW = W / 2;
            ld8(dy + (((long long)n * OH + (y >> 1)) * OW + (x >> 1)) * C + v * 8, o);
#pragma unroll
            for (int e = 0; e < 8; ++e) o[e] *= 0.25f;
        }
        st8(g + (((long long)n * H + y) * W + x) * C + v * 8, o);
    }
}

// dr[n,y,x,c] = scale * dp[n,y/2,x/2,c] * (r > 0)
template <typename T>
__global__ void relu_pool_bwd_kernel(const T* __restrict__ dp, const T* __restrict__ r, T* __restrict__ dr, int N, int H,
                                     int W, int C) {
    const int OH = H / 2, OW = W / 2, nvec = C / 8;
    const long long total = (long long)N * H * W * nvec;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int v = (int)(i % nvec);
        long long t = i / nvec;
        const int x = (int)(t % W);
        t /= W;
        const int y = (int)(t % H);
        const int n = (int)(t / H);
        float g[8], rv[8], o[8];
        ld8(dp + (((long long)n * OH + (y >> 1)) * OW + (x >> 1)) * C + v * 8, g);
        const long long off = (((long long)n * H + y) * W + x) * C + v * 8;
        ld8(r + off, rv);
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] = rv[e] > 0.f ? 0.25f * g[e] : 0.f;
        st8(dr + off, o);
    }
}

// m[n][c] = mean_p x[n][p][c];  grid (C/8-vec groups.., N)
template <typename T>
__global__ void global_mean_kernel(const T* __restrict__ x, float* __restrict__ m, int HW, int C) {
    const int n = blockIdx.y;
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    float s = 0.f;
    for (int p = 0; p < HW; ++p) s += ldf(x + ((long long)n * HW + p) * C + c);
    m[(long long)n * C + c] = s / (float)HW;
}

// same, pixels split over blocks (grid (chunks, N), 256 threads = C/8 channel vectors x pixel lanes): per-thread partial
// sums -> shared-memory atomics -> one global atomic per (block, channel).  m must be zeroed by the caller.
template <typename T>
__global__ void __launch_bounds__(256) global_mean_split_kernel(const T* __restrict__ x, float* __restrict__ m, int HW,
                                                                 int C, int rows) {
    extern __shared__ float gm_sm[];   // [C]
    const int n = blockIdx.y;
    const int nvec = C / 8, PL = 256 / nvec;
    const int v = threadIdx.x % nvec, pl = threadIdx.x / nvec;
    for (int i = threadIdx.x; i < C; i += 256) gm_sm[i] = 0.f;
    __syncthreads();
    const int r0 = blockIdx.x * rows;
    int r1 = r0 + rows;
    if (r1 > HW) r1 = HW;
    if (pl < PL) {
        float s[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) s[e] = 0.f;
        const T* base = x + (long long)n * HW * C + v * 8;
        int r = r0 + pl;
        for (; r + PL < r1; r += 2 * PL) {
            float a[8], b[8];
            ld8(base + (long long)r * C, a);
            ld8(base + (long long)(r + PL) * C, b);
#pragma unroll
            for (int e = 0; e < 8; ++e) s[e] += a[e] + b[e];
        }
        if (r < r1) {
            float a[8];
            ld8(base + (long long)r * C, a);
#pragma unroll
            for (int e = 0; e < 8; ++e) s[e] += a[e];
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) atomicAdd(&gm_sm[v * 8 + e], s[e]);
    }
    __syncthreads();
    const float inv = 1.f / (float)HW;
    for (int i = threadIdx.x; i < C; i += 256) atomicAdd(m + (long long)n * C + i, gm_sm[i] * inv);
}

template <typename T>
__global__ void relu_mean_bwd_kernel(const float* __restrict__ dm, const T* __restrict__ r, T* __restrict__ dr, int N,
                                     int HW, int C) {
    const long long total = (long long)N * HW * C;
    const float inv = 1.f / (float)HW;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        const int n = (int)(i / ((long long)HW * C));
        stf(dr + i, ldf(r + i) > 0.f ? dm[(long long)n * C + c] * inv : 0.f);
    }
}

// dy masked by y > 0 (ReLU backward), in place allowed
template <typename T>
__global__ void relu_mask_kernel(const T* __restrict__ dy, const T* __restrict__ y, T* __restrict__ out, long long n8) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
        float g[8], v[8];
        ld8(dy + i * 8, g);
        ld8(y + i * 8, v);
#pragma unroll
        for (int e = 0; e < 8; ++e) g[e] = v[e] > 0.f ? g[e] : 0.f;
        st8(out + i * 8, g);
    }
}

// one warp per output (n, l)
__global__ void heads_fwd_kernel(const float* __restrict__ m, const float* __restrict__ w, const float* __restrict__ b,
                                 float* __restrict__ out, int N, int C, int L2) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) / 32, lane = threadIdx.x % 32;
    if (warp >= N * L2) return;
    const int n = warp / L2, l = warp % L2;
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s = fmaf(w[(long long)l * C + c], m[(long long)n * C + c], s);
    s = warp_sum(s);
    if (lane == 0) out[warp] = s + b[l];
}

__global__ void heads_bwd_kernel(const float* __restrict__ m, const float* __restrict__ w, const float* __restrict__ dout,
                                 float* __restrict__ dm, float* __restrict__ dw, float* __restrict__ db, int N, int C,
                                 int L2, int accumulate, int acc_dm) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    // dm[n][c]
    if (i < N * C) {
        const int n = i / C, c = i % C;
        float s = 0.f;
        for (int l = 0; l < L2; ++l) s = fmaf(dout[n * L2 + l], w[(long long)l * C + c], s);
        dm[i] = acc_dm ? dm[i] + s : s;
    }
    // dw[l][c]
    if (i < L2 * C) {
        const int l = i / C, c = i % C;
        float s = 0.f;
        for (int n = 0; n < N; ++n) s = fmaf(dout[n * L2 + l], m[(long long)n * C + c], s);
        dw[i] = accumulate ? dw[i] + s : s;
    }
    if (i < L2) {
        float s = 0.f;
        for (int n = 0; n < N; ++n) s += dout[n * L2 + i];
        db[i] = accumulate ? db[i] + s : s;
    }
}

__global__ void rsample_kernel(const float* __restrict__ mu, const float* __restrict__ ls, const float* __restrict__ eps,
                               float* __restrict__ z, float* __restrict__ sigma_out, int* __restrict__ flag, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float sg = expf(ls[i]);
    // separately rounded multiply and add: bit-identical to torch's loc + eps * scale
    z[i] = __fadd_rn(mu[i], __fmul_rn(eps[i], sg));
    if (sigma_out) sigma_out[i] = sg;
    if (flag && (!isfinite(mu[i]) || !(sg > 0.f) || !isfinite(sg))) atomicOr(flag, 1);
}

__global__ void kl_kernel(const float* __restrict__ mu_q, const float* __restrict__ ls_q, const float* __restrict__ mu_p,
                          const float* __restrict__ ls_p, double* __restrict__ kl, float* dmu_q, float* dls_q,
                          float* dmu_p, float* dls_p, const float* gscale_ptr, int n) {
    __shared__ double part[32];
    const float gscale = gscale_ptr ? *gscale_ptr : 1.f;
    double acc = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const float sq = expf(ls_q[i]), sp = expf(ls_p[i]);
        const float ratio = sq / sp;
        const float vr = ratio * ratio;
        const float dmu = (mu_q[i] - mu_p[i]) / sp;
        const float t1 = dmu * dmu;
        acc += (double)(0.5f * (vr + t1 - 1.f - logf(vr)));
        const float gm = gscale * dmu / sp;
        if (dmu_q) dmu_q[i] = gm;
        if (dmu_p) dmu_p[i] = -gm;
        if (dls_q) dls_q[i] = gscale * (vr - 1.f);
        if (dls_p) dls_p[i] = gscale * (1.f - vr - t1);
    }
    acc = warp_sum_d(acc);
    if (threadIdx.x % 32 == 0) part[threadIdx.x / 32] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        double v = threadIdx.x < blockDim.x / 32 ? part[threadIdx.x] : 0.0;
        v = warp_sum_d(v);
        if (threadIdx.x == 0) kl[0] += v;
    }
}

// recon += sum (out - target)^2 ; dlogits (NHWC, dtype) = gscale * 2 (out - target)
template <typename T>
__global__ void mse_kernel(const float* __restrict__ out_nchw, const float* __restrict__ target, double* __restrict__ recon,
                           T* __restrict__ dlogits, const float* gscale_ptr, int N, int C, int HW, int Cdst) {
    __shared__ double part[32];
    const float gscale = gscale_ptr ? *gscale_ptr : 1.f;
    double acc = 0.0;
    const long long total = (long long)N * C * HW;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const float d = out_nchw[i] - target[i];
        acc += (double)(d * d);
        if (dlogits) {
            const int hw = (int)(i % HW);
            const long long t = i / HW;
            const int c = (int)(t % C);
            const int n = (int)(t / C);
            stf(dlogits + ((long long)n * HW + hw) * Cdst + c, gscale * 2.f * d);
        }
    }
    acc = warp_sum_d(acc);
    if (threadIdx.x % 32 == 0) part[threadIdx.x / 32] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        double v = threadIdx.x < blockDim.x / 32 ? part[threadIdx.x] : 0.0;
        v = warp_sum_d(v);
        if (threadIdx.x == 0) atomicAdd(recon, v);
    }
}

// (total, recon, kl) = (recon + beta * kl, recon, kl) as fp32 scalars
__global__ void loss_finalize_kernel(const double* acc, float beta, float* total, float* recon, float* kl) {
    *total = (float)(acc[0] + (double)beta * acc[1]);
    *recon = (float)acc[0];
    *kl = (float)acc[1];
}
// upstream gradients of (total, recon, kl) -> effective seeds for the recon and kl branches
__global__ void loss_bwd_scales_kernel(const float* g_total, const float* g_recon, const float* g_kl, float beta,
                                       float* out2) {
    const float gt = g_total ? *g_total : 0.f;
    out2[0] = gt + (g_recon ? *g_recon : 0.f);
    out2[1] = beta * gt + (g_kl ? *g_kl : 0.f);
}

static unsigned grid_for(long long total, int threads = 256) {
    long long g = cdivll(total, threads);
    if (g > 148LL * 16) g = 148LL * 16;
    if (g < 1) g = 1;
    return (unsigned)g;
}

}  // namespace pu

extern "C" {
using namespace pu;

int pu_avgpool2(const void* x, void* y, int N, int H, int W, int C, int dtype, void* stream) {
    PU_REQUIRE(x && y && N > 0 && H > 0 && W > 0 && H % 2 == 0 && W % 2 == 0 && C % 8 == 0, "pu_avgpool2: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    long long total = (long long)N * (H / 2) * (W / 2) * (C / 8);
    if (dtype == PU_F32)
        avgpool2_kernel<float><<<grid_for(total), 256, 0, st>>>((const float*)x, (float*)y, N, H, W, C);
    else
        avgpool2_kernel<__nv_bfloat16><<<grid_for(total), 256, 0, st>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)y, N, H, W, C);
    return check_launch("avgpool2");
}

int pu_relu_pool_bwd(const void* dp, const void* r, void* dr, int N, int H, int W, int C, int dtype, void* stream) {
    PU_REQUIRE(dp && r && dr && N > 0 && H % 2 == 0 && W % 2 == 0 && C % 8 == 0, "pu_relu_pool_bwd: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    long long total = (long long)N * H * W * (C / 8);
    if (dtype == PU_F32)
        relu_pool_bwd_kernel<float><<<grid_for(total), 256, 0, st>>>((const float*)dp, (const float*)r, (float*)dr, N, H, W, C);
    else
        relu_pool_bwd_kernel<__nv_bfloat16><<<grid_for(total), 256, 0, st>>>((const __nv_bfloat16*)dp, (const __nv_bfloat16*)r,
                                                                            (__nv_bfloat16*)dr, N, H, W, C);
    return check_launch("relu_pool_bwd");
}

int pu_global_mean(const void* x, float* m, int N, int HW, int C, int dtype, void* stream) {
    PU_REQUIRE(x && m && N > 0 && HW > 0 && C > 0, "pu_global_mean: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == PU_BF16 && C % 8 == 0 && C / 8 <= 256 && HW >= 1024) {
        // large images (the Fcomb backward reduces 128x128 pixels): split the pixels over ~4 blocks per SM.  The fp32
        // mode keeps the one-thread-per-channel kernel below: fixed summation order, bit-reproducible mu / log sigma.
        PU_CUDA(cudaMemsetAsync(m, 0, sizeof(float) * (size_t)N * C, st));
        int chunks = cdiv(148 * 4, N);
        const int PL = 256 / (C / 8);
        if (chunks > HW / (4 * PL)) chunks = HW / (4 * PL);
        if (chunks < 1) chunks = 1;
        const int rows = cdiv(HW, chunks);
        dim3 g2(cdiv(HW, rows), N);
        global_mean_split_kernel<__nv_bfloat16><<<g2, 256, sizeof(float) * C, st>>>((const __nv_bfloat16*)x, m, HW, C, rows);
        return check_launch("global_mean");
    }
    dim3 grid(cdiv(C, 128), N);
    if (dtype == PU_F32)
        global_mean_kernel<float><<<grid, 128, 0, st>>>((const float*)x, m, HW, C);
    else
        global_mean_kernel<__nv_bfloat16><<<grid, 128, 0, st>>>((const __nv_bfloat16*)x, m, HW, C);
    return check_launch("global_mean");
}

int pu_relu_mean_bwd(const float* dm, const void* r, void* dr, int N, int HW, int C, int dtype, void* stream) {
    PU_REQUIRE(dm && r && dr && N > 0 && HW > 0 && C > 0, "pu_relu_mean_bwd: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    long long total = (long long)N * HW * C;
    if (dtype == PU_F32)
        relu_mean_bwd_kernel<float><<<grid_for(total), 256, 0, st>>>(dm, (const float*)r, (float*)dr, N, HW, C);
    else
        relu_mean_bwd_kernel<__nv_bfloat16><<<grid_for(total), 256, 0, st>>>(dm, (const __nv_bfloat16*)r, (__nv_bfloat16*)dr, N, HW, C);
    return check_launch("relu_mean_bwd");
}

int pu_relu_mask(const void* dy, const void* y, void* out, long long n, int dtype, void* stream) {
    PU_REQUIRE(dy && y && out && n > 0 && n % 8 == 0, "pu_relu_mask: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == PU_F32)
        relu_mask_kernel<float><<<grid_for(n / 8), 256, 0, st>>>((const float*)dy, (const float*)y, (float*)out, n / 8);
    else
        relu_mask_kernel<__nv_bfloat16><<<grid_for(n / 8), 256, 0, st>>>((const __nv_bfloat16*)dy, (const __nv_bfloat16*)y,
                                                                        (__nv_bfloat16*)out, n / 8);
    return check_launch("relu_mask");
}

int pu_heads_fwd(const float* m, const float* w, const float* b, float* out, int N, int C, int L2, void* stream) {
    PU_REQUIRE(m && w && b && out && N > 0 && C > 0 && L2 > 0, "pu_heads_fwd: bad arguments");
    heads_fwd_kernel<<<cdiv(N * L2 * 32, 128), 128, 0, (cudaStream_t)stream>>>(m, w, b, out, N, C, L2);
    return check_launch("heads_fwd");
}

int pu_heads_bwd(const float* m, const float* w, const float* dout, float* dm, float* dw, float* db, int N, int C,
                 int L2, int accumulate, int acc_dm, void* stream) {
    PU_REQUIRE(m && w && dout && dm && dw && db && N > 0 && C > 0 && L2 > 0, "pu_heads_bwd: bad arguments");
    int total = (N > L2 ? N : L2) * C;
    heads_bwd_kernel<<<cdiv(total, 128), 128, 0, (cudaStream_t)stream>>>(m, w, dout, dm, dw, db, N, C, L2, accumulate, acc_dm);
    return check_launch("heads_bwd");
}

int pu_rsample(const float* mu, const float* log_sigma, const float* eps, float* z, float* sigma_out, int* flag, int n,
               void* stream) {
    PU_REQUIRE(mu && log_sigma && eps && z && n > 0, "pu_rsample: bad arguments");
    rsample_kernel<<<cdiv(n, 128), 128, 0, (cudaStream_t)stream>>>(mu, log_sigma, eps, z, sigma_out, flag, n);
    return check_launch("rsample");
}

int pu_kl_fwd_bwd(const float* mu_q, const float* ls_q, const float* mu_p, const float* ls_p, double* kl_acc,
                  float* dmu_q, float* dls_q, float* dmu_p, float* dls_p, const float* gscale, int n, void* stream) {
    PU_REQUIRE(mu_q && ls_q && mu_p && ls_p && kl_acc && n > 0, "pu_kl_fwd_bwd: bad arguments");
    kl_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(mu_q, ls_q, mu_p, ls_p, kl_acc, dmu_q, dls_q, dmu_p, dls_p, gscale, n);
    return check_launch("kl");
}

int pu_mse_fwd_bwd(const float* out_nchw, const float* target, double* recon_acc, void* dlogits, const float* gscale,
                   int N, int C, int HW, int Cdst, int dtype, void* stream) {
    PU_REQUIRE(out_nchw && target && recon_acc && N > 0 && C > 0 && HW > 0 && (!dlogits || Cdst >= C),
               "pu_mse_fwd_bwd: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    long long total = (long long)N * C * HW;
    unsigned grid = grid_for(total);
    if (grid > 592) grid = 592;
    if (dtype == PU_F32)
        mse_kernel<float><<<grid, 256, 0, st>>>(out_nchw, target, recon_acc, (float*)dlogits, gscale, N, C, HW, Cdst);
    else
        mse_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(out_nchw, target, recon_acc, (__nv_bfloat16*)dlogits, gscale, N, C, HW,
                                                        Cdst);
    return check_launch("mse");
}

int pu_loss_finalize(const double* acc, float beta, float* total, float* recon, float* kl, void* stream) {
    PU_REQUIRE(acc && total && recon && kl, "pu_loss_finalize: bad arguments");
    loss_finalize_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(acc, beta, total, recon, kl);
    return check_launch("loss_finalize");
}

int pu_loss_bwd_scales(const float* g_total, const float* g_recon, const float* g_kl, float beta, float* out2,
                       void* stream) {
    PU_REQUIRE(out2, "pu_loss_bwd_scales: bad arguments");
    loss_bwd_scales_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(g_total, g_recon, g_kl, beta, out2);
    return check_launch("loss_bwd_scales");
}

int pu_resample_grad(const void* dy, void* g, int N, int H, int W, int C, int resample, int dtype, void* stream) {
    PU_REQUIRE(dy && g && N > 0 && H > 0 && W > 0 && C % 8 == 0, "pu_resample_grad: bad arguments");
    PU_REQUIRE(resample == PU_RS_UP || (resample == PU_RS_DOWN && H % 2 == 0 && W % 2 == 0),
               "pu_resample_grad: resample must be PU_RS_UP or PU_RS_DOWN (even H, W); got %d, %dx%d", resample, H, W);
    PU_REQUIRE(dtype == PU_F32 || dtype == PU_BF16, "pu_resample_grad: bad dtype");
    cudaStream_t st = (cudaStream_t)stream;
    const long long total = (long long)N * H * W * (C / 8);
    const bool up = resample == PU_RS_UP;
    if (dtype == PU_F32) {
        if (up) resample_grad_kernel<float, true><<<grid_for(total), 256, 0, st>>>((const float*)dy, (float*)g, N, H, W, C);
        else resample_grad_kernel<float, false><<<grid_for(total), 256, 0, st>>>((const float*)dy, (float*)g, N, H, W, C);
    } else {
        using B = __nv_bfloat16;
        if (up) resample_grad_kernel<B, true><<<grid_for(total), 256, 0, st>>>((const B*)dy, (B*)g, N, H, W, C);
        else resample_grad_kernel<B, false><<<grid_for(total), 256, 0, st>>>((const B*)dy, (B*)g, N, H, W, C);
    }
    return check_launch("resample_grad");
}

int pu_upsample2(const void* x, void* y, int N, int H, int W, int C, int dtype, void* stream) {
    PU_REQUIRE(x && y && N > 0 && H > 0 && W > 0 && C % 8 == 0, "pu_upsample2: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    long long total = (long long)N * H * W * 4 * (C / 8);
    if (dtype == PU_F32)
        upsample2_kernel<float><<<grid_for(total), 256, 0, st>>>((const float*)x, (float*)y, N, H, W, C);
    else
        upsample2_kernel<__nv_bfloat16><<<grid_for(total), 256, 0, st>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)y, N, H, W, C);
    return check_launch("upsample2");
}
}
