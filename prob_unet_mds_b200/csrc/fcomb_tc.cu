// Ensemble Fcomb decode on the 5th-gen tensor cores (prob_unet.py:100-121 evaluated for S latent samples per input,
// train_prob_unet_model.py:179-180).  One CTA owns a 128-pixel tile of one input and walks the members:
//
//   once per tile    h0 = feat_tile[128x64] . W0f^T                 tcgen05.mma (M128 N64 K64) -> TMEM -> registers
//   per member s     A  = relu(h0 + (W0z . z[n,s] + b0)) as bf16    registers -> 128B-swizzled shared memory
//                    D  = A . W1^T                                  tcgen05.mma (M128 N64 K64) -> TMEM
//                    out[c] = b2[c] + sum_j W2[c][j] relu(D[j] + b1[j])   registers, thread == pixel
//
// The tiled z, the concat, and both hidden activations never touch global memory.  Two warpgroups per CTA each run
// their own member pipeline (own A buffer, own accumulator columns, own mbarrier) and two CTAs share an SM, so four
// independent pipelines hide the issue -> commit -> wait round trip of the small MMA.
#include "../../include/probunet_b200.h"
#include "common.cuh"
#include "tc_ptx.cuh"
#include <stdlib.h>

namespace pu {

using namespace ptx;

constexpr int FT_C = 64;               // Fcomb width == unet_output_channels
constexpr int FT_ROWS = 128;           // pixels per CTA == MMA M
constexpr int FT_THREADS = 256;
constexpr int FT_SCHUNK = 128;         // members whose layer-0 bias vectors are staged per CTA
constexpr int FT_A_BYTES = FT_ROWS * 128;
constexpr int FT_W_BYTES = FT_C * 128;
constexpr int FT_SMEM = 1024 + 2 * FT_A_BYTES + 2 * FT_W_BYTES + FT_SCHUNK * FT_C * 4 + FT_C * 16 + 64;

__device__ __forceinline__ float relu_nan(float v) {
    float r;
    asm("max.NaN.f32 %0, %1, 0f00000000;" : "=f"(r) : "f"(v));
    return r;
}

__device__ __forceinline__ void wg_sync(int wg) {
    if (wg == 0)
        asm volatile("bar.sync 1, 128;" ::: "memory");
    else
        asm volatile("bar.sync 2, 128;" ::: "memory");
}

// 8 fp32 -> 8 bf16 into chunk `chunk` of row `r` of a [rows][64] K-major tile with the 128-byte swizzle
__device__ __forceinline__ void store_chunk_sw128(uint8_t* tile, int r, int chunk, const float (&v)[8]) {
    uint4 pk;
    __nv_bfloat162* hp = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
    for (int e = 0; e < 4; ++e) hp[e] = __floats2bfloat162_rn(v[2 * e], v[2 * e + 1]);
    *reinterpret_cast<uint4*>(tile + r * 128 + ((chunk ^ (r & 7)) << 4)) = pk;
}

__global__ void __launch_bounds__(FT_THREADS, 2) fcomb_members_tc_kernel(const __grid_constant__ PuFcombArgs a) {
    extern __shared__ uint8_t ft_smem_raw[];
    uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(ft_smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = base;                                   // warpgroup w at + w*FT_A_BYTES
    uint8_t* sW0 = sA + 2 * FT_A_BYTES;
    uint8_t* sW1 = sW0 + FT_W_BYTES;
    float* zb = reinterpret_cast<float*>(sW1 + FT_W_BYTES);                 // [FT_SCHUNK][64]
    float4* ep = reinterpret_cast<float4*>(zb + FT_SCHUNK * FT_C);           // per hidden unit {b1, w2[0], w2[1], w2[2]}
    uint64_t* bars = reinterpret_cast<uint64_t*>(ep + FT_C);                  // [0] layer 0, [1 + wg] members
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n = blockIdx.y;
    const int p0 = blockIdx.x * FT_ROWS;
    const int s0 = blockIdx.z * FT_SCHUNK;
    const int ns = min(FT_SCHUNK, a.S - s0);
    const int K0 = FT_C + a.L;

    // ---- stage weights (fp32 master -> bf16, swizzled K-major), the feature tile, the per-member bias vectors ----
    for (int i = tid; i < 2 * FT_C * 8; i += FT_THREADS) {
        const int which = i >> 9, o = (i >> 3) & 63, chunk = i & 7;
        const float* src = which ? a.w1 + o * FT_C + chunk * 8 : a.w0 + (long long)o * K0 + chunk * 8;
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = src[e];
        store_chunk_sw128(which ? sW1 : sW0, o, chunk, v);
    }
    {
        const int r = tid >> 1, half = tid & 1;
        const uint4* src = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(a.feat) +
                                                          ((long long)n * a.HW + p0 + r) * FT_C) + half * 4;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const uint4 v = src[c];
            *reinterpret_cast<uint4*>(sA + r * 128 + (((half * 4 + c) ^ (r & 7)) << 4)) = v;
        }
    }
    for (int i = tid; i < ns * FT_C; i += FT_THREADS) {
        const int s = i >> 6, o = i & 63;
        const float* zp = a.z + ((long long)n * a.S + s0 + s) * a.L;
        const float* wp = a.w0 + (long long)o * K0 + FT_C;
        float v = a.b0[o];
        for (int l = 0; l < a.L; ++l) v = fmaf(wp[l], zp[l], v);
        zb[i] = v;
    }
    if (tid < FT_C) {
        float4 e;
        e.x = a.b1[tid];
        e.y = a.w2[tid];
        e.z = a.num_classes > 1 ? a.w2[FT_C + tid] : 0.f;
        e.w = a.num_classes > 2 ? a.w2[2 * FT_C + tid] : 0.f;
        ep[tid] = e;
    }
    if (warp == 1 && lane == 0) {
        for (int b = 0; b < 3; ++b) mbar_init(smem_u32(&bars[b]), 1);
        fence_barrier_init();
    }
    if (warp == 2) {
        tmem_alloc(smem_u32(tmem_slot), 128);
        tmem_relinquish();
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);
    constexpr uint32_t IDESC = idesc_bf16_f32(128, FT_C, 0, 0);

    // ---- layer 0, feature half: shared by every member of this tile ----
    if (warp == 0) {
        const uint32_t a_addr = smem_u32(sA), b_addr = smem_u32(sW0);
#pragma unroll
        for (int k = 0; k < FT_C / 16; ++k)
            mma_f16_ss(tmem_base, smem_desc_sw128(a_addr + k * 32, 16, 1024), smem_desc_sw128(b_addr + k * 32, 16, 1024),
                       IDESC, k ? 1u : 0u);
        mma_commit(smem_u32(&bars[0]));
    }
    mbar_wait(smem_u32(&bars[0]), 0);
    tc_fence_after();
    const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
    float h0[FT_C];
#pragma unroll
    for (int c = 0; c < FT_C; c += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_base + lane_off + c, v);
        tc_wait_ld();
#pragma unroll
        for (int i = 0; i < 32; ++i) h0[c + i] = __uint_as_float(v[i]);
    }
    tc_fence_before();
    __syncthreads();          // accumulator columns and the A buffer are free for the member pipelines

    // ---- members: warpgroup wg handles s = wg, wg + 2, ... ----
    const int wg = warp >> 2;
    const int r = tid & 127;                                  // pixel row == TMEM lane
    uint8_t* sAw = sA + wg * FT_A_BYTES;
    const uint32_t a_addr = smem_u32(sAw), b_addr = smem_u32(sW1);
    const uint32_t tacc = tmem_base + wg * FT_C;
    const uint32_t bar = smem_u32(&bars[1 + wg]);
    const float b2x = a.b2[0], b2y = a.num_classes > 1 ? a.b2[1] : 0.f, b2z = a.num_classes > 2 ? a.b2[2] : 0.f;
    uint32_t phase = 0;
    for (int s = wg; s < ns; s += 2) {
        const float4* zv = reinterpret_cast<const float4*>(zb + s * FT_C);
#pragma unroll
        for (int c8 = 0; c8 < 8; ++c8) {
            const float4 z0 = zv[2 * c8], z1 = zv[2 * c8 + 1];
            float v[8];
            v[0] = relu_nan(h0[c8 * 8 + 0] + z0.x);
            v[1] = relu_nan(h0[c8 * 8 + 1] + z0.y);
            v[2] = relu_nan(h0[c8 * 8 + 2] + z0.z);
            v[3] = relu_nan(h0[c8 * 8 + 3] + z0.w);
            v[4] = relu_nan(h0[c8 * 8 + 4] + z1.x);
            v[5] = relu_nan(h0[c8 * 8 + 5] + z1.y);
            v[6] = relu_nan(h0[c8 * 8 + 6] + z1.z);
            v[7] = relu_nan(h0[c8 * 8 + 7] + z1.w);
            store_chunk_sw128(sAw, r, c8, v);
        }
        fence_proxy_async();
        tc_fence_before();
        wg_sync(wg);
        if ((warp & 3) == 0) {
            tc_fence_after();
#pragma unroll
            for (int k = 0; k < FT_C / 16; ++k)
                mma_f16_ss(tacc, smem_desc_sw128(a_addr + k * 32, 16, 1024), smem_desc_sw128(b_addr + k * 32, 16, 1024),
                           IDESC, k ? 1u : 0u);
            mma_commit(bar);
        }
        mbar_wait(bar, phase);
        phase ^= 1;
        tc_fence_after();
        float o0 = b2x, o1 = b2y, o2 = b2z;
#pragma unroll
        for (int c = 0; c < FT_C; c += 32) {
            uint32_t v[32];
            tmem_ld32(tacc + lane_off + c, v);
            tc_wait_ld();
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const float4 e = ep[c + i];
                const float x = relu_nan(__uint_as_float(v[i]) + e.x);
                o0 = fmaf(x, e.y, o0);
                o1 = fmaf(x, e.z, o1);
                o2 = fmaf(x, e.w, o2);
            }
        }
        tc_fence_before();
        float* op = a.out_nchw + ((long long)n * a.S + s0 + s) * a.num_classes * a.HW + p0 + r;
        op[0] = o0;
        if (a.num_classes > 1) op[a.HW] = o1;
        if (a.num_classes > 2) op[2 * (long long)a.HW] = o2;
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 128);
    }
}

// true when the tensor-core ensemble kernel applies to these arguments
bool fcomb_members_tc_applicable(const PuFcombArgs* a) {
    static int ok = -1;
    if (ok < 0) ok = pu_device_supports_tc();
    const char* e = getenv("PU_FCOMB_TC");      // PU_FCOMB_TC=0 forces the CUDA-core kernel (tests compare the two)
    const bool env = !(e && e[0] == '0');
    return ok && env && a->dtype == PU_BF16 && a->S >= 4 && a->HW % FT_ROWS == 0 && !a->h1_out && !a->h2_out;
}

int fcomb_members_tc_launch(const PuFcombArgs* a, cudaStream_t st) {
    PU_SMEM_ATTR(fcomb_members_tc_kernel, FT_SMEM);
    dim3 grid(a->HW / FT_ROWS, a->N, cdiv(a->S, FT_SCHUNK));
    fcomb_members_tc_kernel<<<grid, FT_THREADS, FT_SMEM, st>>>(*a);
    return check_launch("fcomb_members_tc");
}

}  // namespace pu
