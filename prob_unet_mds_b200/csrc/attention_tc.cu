// tcgen05 flash attention (forward) for the U-Net's self-attention blocks: head dim 64, non-causal, bf16 operands,
// fp32 softmax statistics and accumulation, no T x T matrix in HBM.  Reference: networks.py:112-125,179-184.
//
// One CTA per (q tile of 128 rows, sample*head); it walks the keys in tiles of 128.
//   warp 0   TMA producer : Q tile once, then (K_j, V_j) tiles through a 3-stage ring
//   warp 1   MMA issuer   : S_j = Q K_j^T (K-major operands) into TMEM S[j%2];  O_j = P_j V_j into TMEM O[j%2]
//                           (P_j from shared memory, V_j as an MN-major operand); S_{j+1} is issued before O_j so the
//                           tensor core works while the softmax warps process tile j
//   warp 2   TMEM allocator
//   warps 4-7 softmax     : thread r owns query row r: tcgen05.ld S -> running max / exp2 / row sum, P_j written to
//                           shared memory in the SW128 K-major layout, then acc = acc*corr + O_j read back from TMEM
// Layout: qkv [N*T][3C] with channel = j*C + head*64 + d; out [N*T][C]; lse [N][heads][T] (natural log).
#include <cuda.h>

#include "../../include/probunet_b200.h"
#include "attn_internal.h"
#include "common.cuh"
#include "tc_ptx.cuh"

namespace pu {
using namespace ptx;

int make_mat_tmap(CUtensorMap* m, const void* ptr, long long rows, long long cols, int brows);

constexpr int AT_TQ = 128, AT_TK = 128, AT_D = 64;
constexpr int AT_TILE = 128 * 128;            // one [128 rows][64 bf16] tile = 16 KB
constexpr int AT_KV_STAGES = 3;
constexpr int AT_SMEM = AT_TILE /*Q*/ + AT_KV_STAGES * 2 * AT_TILE /*K,V*/ + 2 * 2 * AT_TILE /*P ping-pong*/ + 1024 + 256;

struct AttnFwdParams {
    int T, heads, C;
    __nv_bfloat16* out;
    float* lse;
};

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__global__ void __launch_bounds__(256, 1)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmQKV, const AttnFwdParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sQ = smem;
    uint8_t* sKV = sQ + AT_TILE;                              // stage s: K at +s*2*TILE, V at +s*2*TILE + TILE
    uint8_t* sP = sKV + AT_KV_STAGES * 2 * AT_TILE;           // buffer b at +b*2*TILE (two 64-key K-blocks)
    uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * 2 * AT_TILE);
    uint64_t* q_full = bars;                 // 1
    uint64_t* kv_full = bars + 1;            // [3]
    uint64_t* kv_empty = kv_full + AT_KV_STAGES;
    uint64_t* s_full = kv_empty + AT_KV_STAGES;   // [2]
    uint64_t* s_empty = s_full + 2;
    uint64_t* p_full = s_empty + 2;
    uint64_t* o_full = p_full + 2;
    uint64_t* o_empty = o_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nh = blockIdx.y, n = nh / p.heads, h = nh % p.heads;
    const int q0 = blockIdx.x * AT_TQ;
    const int ntiles = p.T / AT_TK;
    const int row_base = n * p.T;
    const int colQ = h * AT_D, colK = p.C + h * AT_D, colV = 2 * p.C + h * AT_D;

    if (warp == 0 && lane == 0) prefetch_tmap(&tmQKV);
    if (warp == 1 && lane == 0) {
        mbar_init(smem_u32(q_full), 1);
        for (int s = 0; s < AT_KV_STAGES; ++s) {
            mbar_init(smem_u32(&kv_full[s]), 1);
            mbar_init(smem_u32(&kv_empty[s]), 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(smem_u32(&s_full[b]), 1);
            mbar_init(smem_u32(&s_empty[b]), 4);
            mbar_init(smem_u32(&p_full[b]), 4);
            mbar_init(smem_u32(&o_full[b]), 1);
            mbar_init(smem_u32(&o_empty[b]), 4);
        }
        fence_barrier_init();
    }
    if (warp == 2) {
        tmem_alloc(smem_u32(tmem_slot), 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);
    const uint32_t tS = tmem_base;           // S[b] at + b*128
    const uint32_t tO = tmem_base + 256;     // O[b] at + b*64

    if (warp == 0) {
        if (lane == 0) {
            mbar_expect_tx(smem_u32(q_full), AT_TILE);
            tma_load_2d(smem_u32(sQ), &tmQKV, smem_u32(q_full), colQ, row_base + q0);
            int stage = 0;
            uint32_t phase = 0;
            for (int j = 0; j < ntiles; ++j) {
                mbar_wait(smem_u32(&kv_empty[stage]), phase ^ 1);
                const uint32_t fb = smem_u32(&kv_full[stage]);
                mbar_expect_tx(fb, 2 * AT_TILE);
                tma_load_2d(smem_u32(sKV + stage * 2 * AT_TILE), &tmQKV, fb, colK, row_base + j * AT_TK);
                tma_load_2d(smem_u32(sKV + stage * 2 * AT_TILE + AT_TILE), &tmQKV, fb, colV, row_base + j * AT_TK);
                if (++stage == AT_KV_STAGES) {
                    stage = 0;
                    phase ^= 1;
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t IDESC_S = idesc_bf16_f32(128, 128, 0, 0);
            constexpr uint32_t IDESC_O = idesc_bf16_f32(128, 64, 0, 1);
            mbar_wait(smem_u32(q_full), 0);
            const uint32_t q_addr = smem_u32(sQ);
            auto issue_pv = [&](int jj, int st) {
                const int b = jj & 1;
                const uint32_t ph = (jj >> 1) & 1;
                mbar_wait(smem_u32(&p_full[b]), ph);
                mbar_wait(smem_u32(&o_empty[b]), ph ^ 1);
                tc_fence_after();
                const uint32_t p_addr = smem_u32(sP + b * 2 * AT_TILE);
                const uint32_t v_addr = smem_u32(sKV + st * 2 * AT_TILE + AT_TILE);
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const uint64_t ad = smem_desc_sw128(p_addr + (k >> 2) * AT_TILE + (k & 3) * 32, 16, 1024);
                    const uint64_t bd = smem_desc_sw128(v_addr + k * 2048, 8192, 1024);
                    mma_f16_ss(tO + b * 64, ad, bd, IDESC_O, k ? 1u : 0u);
                }
                mma_commit(smem_u32(&o_full[b]));
                mma_commit(smem_u32(&kv_empty[st]));
            };
            int stage = 0;
            uint32_t phase = 0;
            int prev_stage = 0;
            for (int j = 0; j < ntiles; ++j) {
                const int b = j & 1;
                mbar_wait(smem_u32(&kv_full[stage]), phase);
                mbar_wait(smem_u32(&s_empty[b]), ((j >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t k_addr = smem_u32(sKV + stage * 2 * AT_TILE);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint64_t ad = smem_desc_sw128(q_addr + k * 32, 16, 1024);
                    const uint64_t bd = smem_desc_sw128(k_addr + k * 32, 16, 1024);
                    mma_f16_ss(tS + b * 128, ad, bd, IDESC_S, k ? 1u : 0u);
                }
                mma_commit(smem_u32(&s_full[b]));
                if (j > 0) issue_pv(j - 1, prev_stage);
                prev_stage = stage;
                if (++stage == AT_KV_STAGES) {
                    stage = 0;
                    phase ^= 1;
                }
            }
            issue_pv(ntiles - 1, prev_stage);
        }
    } else if (warp >= 4) {
        const int q = warp - 4;
        const int r = q * 32 + lane;                       // query row within the tile == TMEM lane
        const uint32_t lane_off = (uint32_t)(q * 32) << 16;
        const float sc = 0.125f * 1.4426950408889634f;     // 1/sqrt(64) * log2(e)
        float m = -INFINITY, l = 0.f;
        float acc[AT_D];
#pragma unroll
        for (int d = 0; d < AT_D; ++d) acc[d] = 0.f;
        float corr_prev = 1.f;
        auto accumulate = [&](int jj, float corr) {
            const int b = jj & 1;
            mbar_wait(smem_u32(&o_full[b]), (jj >> 1) & 1);
            tc_fence_after();
#pragma unroll
            for (int c = 0; c < AT_D; c += 32) {
                uint32_t v[32];
                tmem_ld32(tO + lane_off + b * 64 + c, v);
                tc_wait_ld();
#pragma unroll
                for (int i = 0; i < 32; ++i) acc[c + i] = fmaf(acc[c + i], corr, __uint_as_float(v[i]));
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&o_empty[b]));
        };
        for (int j = 0; j < ntiles; ++j) {
            const int b = j & 1;
            mbar_wait(smem_u32(&s_full[b]), (j >> 1) & 1);
            tc_fence_after();
            // pass 1: row max
            float mx = -INFINITY;
#pragma unroll 1
            for (int c = 0; c < AT_TK; c += 32) {
                uint32_t v[32];
                tmem_ld32(tS + lane_off + b * 128 + c, v);
                tc_wait_ld();
#pragma unroll
                for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(v[i]));
            }
            const float m_new = fmaxf(m, mx * sc);
            const float corr = ex2_approx(m - m_new);
            // pass 2: p = exp2(s*sc - m_new) -> bf16 -> swizzled shared memory
            float rs = 0.f;
            uint8_t* pb = sP + b * 2 * AT_TILE;
#pragma unroll 1
            for (int c = 0; c < AT_TK; c += 32) {
                uint32_t v[32];
                tmem_ld32(tS + lane_off + b * 128 + c, v);
                tc_wait_ld();
                float pv[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    pv[i] = ex2_approx(fmaf(__uint_as_float(v[i]), sc, -m_new));
                    rs += pv[i];
                }
                uint8_t* kb_base = pb + (c >> 6) * AT_TILE + r * 128;
                const int chunk0 = (c & 63) >> 3;          // 16-byte chunk index of key c within the 64-key block
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    uint4 pk;
                    __nv_bfloat162* hp = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
                    for (int e = 0; e < 4; ++e) hp[e] = __floats2bfloat162_rn(pv[g * 8 + 2 * e], pv[g * 8 + 2 * e + 1]);
                    *reinterpret_cast<uint4*>(kb_base + (((chunk0 + g) ^ (r & 7)) << 4)) = pk;
                }
            }
            l = l * corr + rs;
            m = m_new;
            // S[b] fully read, P[b] fully written: release S to the MMA warp and publish P to the async proxy
            tc_fence_before();
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(smem_u32(&s_empty[b]));
                mbar_arrive(smem_u32(&p_full[b]));
            }
            if (j > 0) accumulate(j - 1, corr_prev);
            corr_prev = corr;
        }
        accumulate(ntiles - 1, corr_prev);
        const float inv = 1.f / l;
        const long long row = (long long)row_base + q0 + r;
        __nv_bfloat16* op = p.out + row * p.C + h * AT_D;
#pragma unroll
        for (int d = 0; d < AT_D; d += 8) {
            float o8[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) o8[e] = acc[d + e] * inv;
            st8(op + d, o8);
        }
        p.lse[((long long)n * p.heads + h) * p.T + q0 + r] = (m + log2f(l)) * 0.6931471805599453f;
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

bool attention_tc_applicable(int N, int T, int heads, int dtype) {
    static int ok = -1;
    if (ok < 0) ok = pu_device_supports_tc();
    return ok == 1 && dtype == PU_BF16 && T >= 128 && T % 128 == 0 && (long long)N * T < (1LL << 31);
}

int attention_fwd_tc(const void* qkv, void* out, float* lse, int N, int T, int heads, cudaStream_t st) {
    const int C = heads * AT_D;
    CUtensorMap tm;
    int rc = make_mat_tmap(&tm, qkv, (long long)N * T, 3LL * C, 128);
    if (rc) return rc;
    static bool attr = false;
    if (!attr) {
        PU_CUDA(cudaFuncSetAttribute(attn_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM));
        attr = true;
    }
    AttnFwdParams p;
    p.T = T; p.heads = heads; p.C = C;
    p.out = (__nv_bfloat16*)out;
    p.lse = lse;
    dim3 grid(T / AT_TQ, N * heads);
    attn_fwd_tc_kernel<<<grid, 256, AT_SMEM, st>>>(tm, p);
    return check_launch("attn_fwd_tc");
}

bool attention_bwd_tc_applicable(int, int, int, int) { return false; }

int attention_bwd_tc(const void*, const void*, const float*, const float*, void*, int, int, int, cudaStream_t) {
    set_error("attention_bwd_tc: not built");
    return PU_ERR_UNSUPPORTED;
}

}  // namespace pu
