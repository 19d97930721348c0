// Layout conversion and weight packing kernels (pure data movement, HBM bound, tiny next to the convs).
#include "../../include/probunet_b200.h"
#include "common.cuh"

namespace pu {

template <typename T>
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ src, T* __restrict__ dst, int N, int C, int HW, int Cdst,
                                    int c_off) {
    long long total = (long long)N * HW * C;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        int c = (int)(i % C);
        long long t = i / C;
        int hw = (int)(t % HW);
        int n = (int)(t / HW);
        stf(dst + ((long long)n * HW + hw) * Cdst + c_off + c, src[((long long)n * C + c) * HW + hw]);
    }
}

template <typename T>
__global__ void nhwc_to_nchw_kernel(const T* __restrict__ src, float* __restrict__ dst, int N, int C, int HW, int Csrc) {
    long long total = (long long)N * HW * C;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        int hw = (int)(i % HW);
        long long t = i / HW;
        int c = (int)(t % C);
        int n = (int)(t / C);
        dst[i] = ldf(src + ((long long)n * HW + hw) * Csrc + c);
    }
}

template <typename T>
__global__ void pack_weight_kernel(const float* __restrict__ src, T* __restrict__ dst, int Co, int Ci, int k, int Ci_pad,
                                   int mode, const int* __restrict__ perm, long long src_co_stride) {
    const int kk = k * k;
    long long total = (mode == 0) ? (long long)Co * kk * Ci_pad : (long long)Ci * kk * Co;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        float v = 0.f;
        if (mode == 0) {
            int ci = (int)(i % Ci_pad);
            long long t = i / Ci_pad;
            int tap = (int)(t % kk);
            int co = (int)(t / kk);
            if (ci < Ci) {
                int sco = perm ? perm[co] : co;
                v = src[(long long)sco * src_co_stride + (long long)ci * kk + tap];
            }
        } else {
            int co = (int)(i % Co);
            long long t = i / Co;
            int tap = (int)(t % kk);
            int ci = (int)(t / kk);
            int sco = perm ? perm[co] : co;
            v = src[(long long)sco * src_co_stride + (long long)ci * kk + (kk - 1 - tap)];
        }
        stf(dst + i, v);
    }
}

// mode 2: forward packing with the bf16 rounding error of every weight kept as a second K block per tap:
//   dst[co][tap][0 .. Ci_pad)        = bf16(w)
//   dst[co][tap][Ci_pad .. 2 Ci_pad) = bf16(w - float(bf16(w)))
// A convolution over the channel concatenation [x ; x] with this weight computes (w_hi + w_lo) * x, i.e. the weights
// enter with ~16 mantissa bits instead of 8 while the kernel, its operands and its accumulation stay what they are.
__global__ void pack_weight_split_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int Co, int Ci,
                                         int k, int Ci_pad, const int* __restrict__ perm, long long src_co_stride) {
    const int kk = k * k;
    long long total = (long long)Co * kk * Ci_pad;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        int ci = (int)(i % Ci_pad);
        long long t = i / Ci_pad;
        int tap = (int)(t % kk);
        int co = (int)(t / kk);
        float v = 0.f;
        if (ci < Ci) {
            int sco = perm ? perm[co] : co;
            v = src[(long long)sco * src_co_stride + (long long)ci * kk + tap];
        }
        const __nv_bfloat16 hi = __float2bfloat16_rn(v);
        const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
        __nv_bfloat16* d = dst + ((long long)co * kk + tap) * 2 * Ci_pad + ci;
        d[0] = hi;
        d[Ci_pad] = lo;
    }
}

// ---- multi-tensor packing: every conv weight of a model in ONE launch (after each optimizer step) ----
// One block per (item, 32 output channels, 32 input channels) tile: the tile's [32][32*k*k] fp32 source rows are read
// with coalesced loads into shared memory and written back as contiguous runs of the destination layout (32 input
// channels for the forward / split packing, 32 output channels for the data-gradient packing).
__device__ __forceinline__ void pack_store(void* dst, long long i, float v, int dtype) {
    if (dtype == PU_F32)
        reinterpret_cast<float*>(dst)[i] = v;
    else
        reinterpret_cast<__nv_bfloat16*>(dst)[i] = __float2bfloat16_rn(v);
}

__global__ void __launch_bounds__(256) pack_weights_multi_kernel(const PuPackItem* __restrict__ items, int n_items) {
    __shared__ float t[32][32 * 9 + 1];
    // binary search: last item with tile_begin <= blockIdx.x
    int lo = 0, hi = n_items - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (items[mid].tile_begin <= (int)blockIdx.x) lo = mid; else hi = mid - 1;
    }
    const PuPackItem it = items[lo];
    const int kk = it.k * it.k;
    const int ci_tiles = (it.Ci_pad + 31) / 32;
    const int tile = blockIdx.x - it.tile_begin;
    const int co0 = (tile / ci_tiles) * 32, ci0 = (tile % ci_tiles) * 32;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int ncol = it.Ci - ci0;                 // real input channels in this tile (<= 0: padding only)
    ncol = ncol > 32 ? 32 : (ncol < 0 ? 0 : ncol);
    for (int r = warp; r < 32; r += 8) {
        const int co = co0 + r;
        if (co < it.Co) {
            const int sco = it.perm ? it.perm[co] : co;
            const float* s = it.src + (long long)sco * it.src_co_stride + (long long)ci0 * kk;
            for (int j = lane; j < 32 * kk; j += 32) t[r][j] = j < ncol * kk ? s[j] : 0.f;
        }
    }
    __syncthreads();
    if (it.mode != 1) {
        const int cstride = it.mode == 2 ? 2 * it.Ci_pad : it.Ci_pad;
        const int ci = ci0 + lane;
        for (int rt = warp; rt < 32 * kk; rt += 8) {
            const int r = rt / kk, tap = rt - r * kk;
            const int co = co0 + r;
            if (co >= it.Co || ci >= it.Ci_pad) continue;
            const float v = t[r][lane * kk + tap];
            const long long o = ((long long)co * kk + tap) * cstride + ci;
            if (it.mode == 2) {
                const __nv_bfloat16 h = __float2bfloat16_rn(v);
                reinterpret_cast<__nv_bfloat16*>(it.dst)[o] = h;
                reinterpret_cast<__nv_bfloat16*>(it.dst)[o + it.Ci_pad] = __float2bfloat16_rn(v - __bfloat162float(h));
            } else {
                pack_store(it.dst, o, v, it.dtype);
            }
        }
    } else {
        const int co = co0 + lane;
        for (int ct = warp; ct < 32 * kk; ct += 8) {
            const int c = ct / kk, tap = ct - c * kk;
            const int ci = ci0 + c;
            if (ci >= it.Ci || co >= it.Co) continue;
            pack_store(it.dst, ((long long)ci * kk + (kk - 1 - tap)) * it.Co + co, t[lane][c * kk + tap], it.dtype);
        }
    }
}

__global__ void unpack_wgrad_kernel(const float* __restrict__ src, float* __restrict__ dst, int Co, int Ci, int k,
                                    int Ci_pad, const int* __restrict__ perm, int accumulate, long long dst_co_stride) {
    const int kk = k * k;
    long long total = (long long)Co * Ci * kk;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        int tap = (int)(i % kk);
        long long t = i / kk;
        int ci = (int)(t % Ci);
        int co = (int)(t / Ci);
        int dco = perm ? perm[co] : co;
        float v = src[((long long)co * kk + tap) * Ci_pad + ci];
        float* d = dst + (long long)dco * dst_co_stride + (long long)ci * kk + tap;
        *d = accumulate ? (*d + v) : v;
    }
}

// 3x3 variant: one block per (128 input channels, output channel).  The nine [tap][ci] source rows are read with
// coalesced loads into shared memory and written back as one contiguous run of 128*9 floats in [ci][tap] order
// (the element-wise kernel above reads with a 9-way stride).
__global__ void __launch_bounds__(128) unpack_wgrad3_kernel(const float* __restrict__ src, float* __restrict__ dst, int Ci,
                                                            int Ci_pad, const int* __restrict__ perm, int accumulate,
                                                            long long dst_co_stride) {
    __shared__ float t[9][129];
    const int co = blockIdx.y, ci0 = blockIdx.x * 128, tid = threadIdx.x;
    const int nci = min(128, Ci - ci0);
#pragma unroll
    for (int tap = 0; tap < 9; ++tap)
        t[tap][tid] = tid < nci ? src[((long long)co * 9 + tap) * Ci_pad + ci0 + tid] : 0.f;
    __syncthreads();
    const int dco = perm ? perm[co] : co;
    float* d = dst + (long long)dco * dst_co_stride + (long long)ci0 * 9;
    for (int j = tid; j < nci * 9; j += 128) {
        const int ci = j / 9, tap = j - ci * 9;
        const float v = t[tap][ci];
        d[j] = accumulate ? d[j] + v : v;
    }
}

__global__ void gather_kernel(const float* src, const int* perm, float* dst, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[perm[i]];
}
__global__ void scatter_kernel(const float* src, const int* perm, float* dst, int n, int accumulate) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        float* d = dst + perm[i];
        *d = accumulate ? (*d + src[i]) : src[i];
    }
}

static unsigned grid_for(long long total, int threads = 256) {
    long long g = cdivll(total, threads);
    if (g > 148LL * 16) g = 148LL * 16;
    if (g < 1) g = 1;
    return (unsigned)g;
}

}  // namespace pu

extern "C" {

int pu_nchw_to_nhwc(const float* src, void* dst, int N, int C, int H, int W, int Cdst, int c_off, int dst_dtype,
                    void* stream) {
    PU_REQUIRE(src && dst && N > 0 && C > 0 && H > 0 && W > 0 && c_off >= 0 && c_off + C <= Cdst,
               "pu_nchw_to_nhwc: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    long long total = (long long)N * C * H * W;
    if (dst_dtype == PU_F32)
        pu::nchw_to_nhwc_kernel<float><<<pu::grid_for(total), 256, 0, st>>>(src, (float*)dst, N, C, H * W, Cdst, c_off);
    else
        pu::nchw_to_nhwc_kernel<__nv_bfloat16><<<pu::grid_for(total), 256, 0, st>>>(src, (__nv_bfloat16*)dst, N, C,
                                                                                  H * W, Cdst, c_off);
    return pu::check_launch("nchw_to_nhwc");
}

int pu_nhwc_to_nchw(const void* src, float* dst, int N, int C, int H, int W, int Csrc, int src_dtype, void* stream) {
    PU_REQUIRE(src && dst && N > 0 && C > 0 && H > 0 && W > 0 && Csrc >= C, "pu_nhwc_to_nchw: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    long long total = (long long)N * C * H * W;
    if (src_dtype == PU_F32)
        pu::nhwc_to_nchw_kernel<float><<<pu::grid_for(total), 256, 0, st>>>((const float*)src, dst, N, C, H * W, Csrc);
    else
        pu::nhwc_to_nchw_kernel<__nv_bfloat16><<<pu::grid_for(total), 256, 0, st>>>((const __nv_bfloat16*)src, dst, N,
                                                                                  C, H * W, Csrc);
    return pu::check_launch("nhwc_to_nchw");
}

int pu_pack_conv_weight(const float* src, void* dst, int Co, int Ci, int k, int Ci_pad, int mode, const int* out_perm,
                        long long src_co_stride, int dtype, void* stream) {
    PU_REQUIRE(src && dst && Co > 0 && Ci > 0 && (k == 1 || k == 3) && Ci_pad >= Ci && mode >= 0 && mode <= 2,
               "pu_pack_conv_weight: bad arguments");
    PU_REQUIRE(mode != 1 || Ci_pad == Ci, "pu_pack_conv_weight: dgrad packing takes no channel padding");
    PU_REQUIRE(mode != 2 || dtype == PU_BF16, "pu_pack_conv_weight: the hi/lo split packing (mode 2) is bf16 only");
    cudaStream_t st = (cudaStream_t)stream;
    if (src_co_stride <= 0) src_co_stride = (long long)Ci * k * k;
    if (mode == 2) {
        pu::pack_weight_split_kernel<<<pu::grid_for((long long)Co * k * k * Ci_pad), 256, 0, st>>>(
            src, (__nv_bfloat16*)dst, Co, Ci, k, Ci_pad, out_perm, src_co_stride);
        return pu::check_launch("pack_conv_weight");
    }
    long long total = (mode == 0) ? (long long)Co * k * k * Ci_pad : (long long)Ci * k * k * Co;
    if (dtype == PU_F32)
        pu::pack_weight_kernel<float><<<pu::grid_for(total), 256, 0, st>>>(src, (float*)dst, Co, Ci, k, Ci_pad, mode,
                                                                         out_perm, src_co_stride);
    else
        pu::pack_weight_kernel<__nv_bfloat16><<<pu::grid_for(total), 256, 0, st>>>(src, (__nv_bfloat16*)dst, Co, Ci, k,
                                                                                 Ci_pad, mode, out_perm, src_co_stride);
    return pu::check_launch("pack_conv_weight");
}

int pu_pack_conv_weights_multi(const PuPackItem* items, int n_items, int total_tiles, void* stream) {
    PU_REQUIRE(items && n_items > 0 && total_tiles > 0, "pu_pack_conv_weights_multi: bad arguments");
    pu::pack_weights_multi_kernel<<<total_tiles, 256, 0, (cudaStream_t)stream>>>(items, n_items);
    return pu::check_launch("pack_conv_weights_multi");
}

int pu_unpack_conv_wgrad(const float* src, float* dst, int Co, int Ci, int k, int Ci_pad, const int* out_perm,
                         long long dst_co_stride, int accumulate, void* stream) {
    PU_REQUIRE(src && dst && Co > 0 && Ci > 0 && (k == 1 || k == 3) && Ci_pad >= Ci, "pu_unpack_conv_wgrad: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    if (dst_co_stride <= 0) dst_co_stride = (long long)Ci * k * k;
    long long total = (long long)Co * Ci * k * k;
    if (k == 3 && Co <= 65535) {
        dim3 grid(pu::cdiv(Ci, 128), Co);
        pu::unpack_wgrad3_kernel<<<grid, 128, 0, st>>>(src, dst, Ci, Ci_pad, out_perm, accumulate, dst_co_stride);
        return pu::check_launch("unpack_conv_wgrad");
    }
    pu::unpack_wgrad_kernel<<<pu::grid_for(total), 256, 0, st>>>(src, dst, Co, Ci, k, Ci_pad, out_perm, accumulate,
                                                                dst_co_stride);
    return pu::check_launch("unpack_conv_wgrad");
}

int pu_gather_f32(const float* src, const int* perm, float* dst, int n, void* stream) {
    PU_REQUIRE(src && perm && dst && n > 0, "pu_gather_f32: bad arguments");
    pu::gather_kernel<<<pu::cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(src, perm, dst, n);
    return pu::check_launch("gather_f32");
}
int pu_scatter_f32(const float* src, const int* perm, float* dst, int n, int accumulate, void* stream) {
    PU_REQUIRE(src && perm && dst && n > 0, "pu_scatter_f32: bad arguments");
    pu::scatter_kernel<<<pu::cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(src, perm, dst, n, accumulate);
    return pu::check_launch("scatter_f32");
}
}
