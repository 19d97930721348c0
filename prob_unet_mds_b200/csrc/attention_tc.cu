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
#include <stdlib.h>

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

// 2^x for x <= 0 on the FMA / integer pipes (no MUFU): round-to-nearest split x = n + f with |f| <= 0.5, cubic
// minimax for 2^f (relative error ~1e-4, far below the bf16 rounding of P), exponent added through the bit pattern.
// Used for a fraction of the softmax exponentials so that MUFU.EX2 (16/clk/SM) is not the only path.
__device__ __forceinline__ float ex2_poly(float x) {
    x = fmaxf(x, -120.f);
    const float t = x + 12582912.f;                 // 1.5 * 2^23: integer part lands in the low mantissa bits
    const float f = x - (t - 12582912.f);
    float pz = fmaf(f, 0.0555041f, 0.2402265f);
    pz = fmaf(pz, f, 0.6931472f);
    pz = fmaf(pz, f, 1.0f);
    return __int_as_float(__float_as_int(pz) + (__float_as_int(t) << 23));
}
// two of them at once with packed fp32 instructions (FFMA2 / FADD2): ~5.5 instead of 8 instructions per exponential
__device__ __forceinline__ float2 ex2_poly2(float2 x) {
    x.x = fmaxf(x.x, -120.f);
    x.y = fmaxf(x.y, -120.f);
    const float2 magic = make_float2(12582912.f, 12582912.f);
    const float2 t = fadd2(x, magic);
    const float2 f = fadd2(x, fadd2(magic, make_float2(-t.x, -t.y)));          // x - (t - magic)
    float2 pz = ffma2(f, make_float2(0.0555041f, 0.0555041f), make_float2(0.2402265f, 0.2402265f));
    pz = ffma2(pz, f, make_float2(0.6931472f, 0.6931472f));
    pz = ffma2(pz, f, make_float2(1.0f, 1.0f));
    return make_float2(__int_as_float(__float_as_int(pz.x) + (__float_as_int(t.x) << 23)),
                       __int_as_float(__float_as_int(pz.y) + (__float_as_int(t.y) << 23)));
}
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));      // not volatile: a pure function the scheduler may move
    return y;
}

__global__ void __launch_bounds__(256, 1)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmQKV, const AttnFwdParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // 1024-aligned; an OFFSET from the
    // __shared__ symbol, so that plain C++ accesses below compile to LDS / STS instead of generic LD / ST
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
        {   // whole warp, warp-uniform control flow; mma_f16_ss / mma_commit elect one lane internally
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
                const float2 corr2 = make_float2(corr, corr);
#pragma unroll
                for (int i = 0; i < 32; i += 2) {
                    const float2 a2 = ffma2(make_float2(acc[c + i], acc[c + i + 1]), corr2,
                                            make_float2(__uint_as_float(v[i]), __uint_as_float(v[i + 1])));
                    acc[c + i] = a2.x;
                    acc[c + i + 1] = a2.y;
                }
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
        for (int d = 0; d < AT_D; d += 16) {
            float o16[16];
#pragma unroll
            for (int e = 0; e < 16; ++e) o16[e] = acc[d + e] * inv;
            st16(op + d, o16);
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

// ------------------------------------------------------------------------------------------------
// forward, two query tiles per CTA (256 rows): the K/V tiles are loaded once and used by both q tiles, and two
// softmax warpgroups (warps 4-7 -> q tile 0, warps 8-11 -> q tile 1) keep the MUFU busy while the tensor core
// works for the other tile.  TMEM: S0 | S1 (128 cols each) | O0 | O1 (64 cols each).  Used when T % 256 == 0.
// ------------------------------------------------------------------------------------------------
constexpr int A2_KV_STAGES = 3;
#ifndef A2_POLY_EVERY
#define A2_POLY_EVERY 4               // every 4th exponential of a row goes to the FMA pipe (ex2_poly)
#endif
constexpr int A2_SMEM = 2 * AT_TILE /*Q0,Q1*/ + A2_KV_STAGES * 2 * AT_TILE /*K,V*/ + 1024 + 256;

template <int ABL = 0, int POLY_EVERY = A2_POLY_EVERY>
__global__ void __launch_bounds__(384, 1)
attn_fwd_tc2_kernel(const __grid_constant__ CUtensorMap tmQKV, const AttnFwdParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // 1024-aligned; an OFFSET from the
    // __shared__ symbol, so that plain C++ accesses below compile to LDS / STS instead of generic LD / ST
    uint8_t* sQ = smem;                                        // q tile t at + t*TILE
    uint8_t* sKV = sQ + 2 * AT_TILE;                           // stage s: K at +s*2*TILE, V at +s*2*TILE + TILE
    uint64_t* bars = reinterpret_cast<uint64_t*>(sKV + A2_KV_STAGES * 2 * AT_TILE);
    uint64_t* q_full = bars;                       // 1
    uint64_t* kv_full = bars + 1;                  // [2]
    uint64_t* kv_empty = kv_full + A2_KV_STAGES;   // [2]
    uint64_t* s_full = kv_empty + A2_KV_STAGES;    // [2] per q tile
    uint64_t* s_empty = s_full + 2;
    uint64_t* p_full = s_empty + 2;
    uint64_t* o_full = p_full + 2;
    uint64_t* o_empty = o_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nh = blockIdx.y, n = nh / p.heads, h = nh % p.heads;
    const int q0 = blockIdx.x * 2 * AT_TQ;
    const int ntiles = p.T / AT_TK;
    const int row_base = n * p.T;
    const int colQ = h * AT_D, colK = p.C + h * AT_D, colV = 2 * p.C + h * AT_D;

    if (warp == 0 && lane == 0) prefetch_tmap(&tmQKV);
    if (warp == 1 && lane == 0) {
        mbar_init(smem_u32(q_full), 1);
        for (int s = 0; s < A2_KV_STAGES; ++s) {
            mbar_init(smem_u32(&kv_full[s]), 1);
            mbar_init(smem_u32(&kv_empty[s]), 1);
        }
        for (int t = 0; t < 2; ++t) {
            mbar_init(smem_u32(&s_full[t]), 1);
            mbar_init(smem_u32(&s_empty[t]), 4);
            mbar_init(smem_u32(&p_full[t]), 4);
            mbar_init(smem_u32(&o_full[t]), 1);
            mbar_init(smem_u32(&o_empty[t]), 4);
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
    const uint32_t tS = tmem_base;           // S[t] at + t*128
    const uint32_t tO = tmem_base + 256;     // O[t] at + t*64
    const uint32_t tP = tmem_base + 384;     // P[t] at + t*64: bf16 pairs, the A operand of O = P V (never in smem)

    if (warp == 0) {
        if (lane == 0) {
            mbar_expect_tx(smem_u32(q_full), 2 * AT_TILE);
            tma_load_2d(smem_u32(sQ), &tmQKV, smem_u32(q_full), colQ, row_base + q0);
            tma_load_2d(smem_u32(sQ + AT_TILE), &tmQKV, smem_u32(q_full), colQ, row_base + q0 + AT_TQ);
            int stage = 0;
            uint32_t phase = 0;
            for (int j = 0; j < ntiles; ++j) {
                mbar_wait(smem_u32(&kv_empty[stage]), phase ^ 1);
                const uint32_t fb = smem_u32(&kv_full[stage]);
                mbar_expect_tx(fb, 2 * AT_TILE);
                tma_load_2d(smem_u32(sKV + stage * 2 * AT_TILE), &tmQKV, fb, colK, row_base + j * AT_TK);
                tma_load_2d(smem_u32(sKV + stage * 2 * AT_TILE + AT_TILE), &tmQKV, fb, colV, row_base + j * AT_TK);
                if (++stage == A2_KV_STAGES) {
                    stage = 0;
                    phase ^= 1;
                }
            }
        }
    } else if (warp == 1) {
        {   // whole warp, warp-uniform control flow; mma_f16_ss / mma_commit elect one lane internally
            constexpr uint32_t IDESC_S = idesc_bf16_f32(128, 128, 0, 0);
            constexpr uint32_t IDESC_O = idesc_bf16_f32(128, 64, 0, 1);
            mbar_wait(smem_u32(q_full), 0);
            auto issue_s = [&](int t, uint32_t k_addr) {
                const uint32_t q_addr = smem_u32(sQ + t * AT_TILE);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    mma_f16_ss(tS + t * 128, smem_desc_sw128(q_addr + k * 32, 16, 1024),
                               smem_desc_sw128(k_addr + k * 32, 16, 1024), IDESC_S, k ? 1u : 0u);
                mma_commit(smem_u32(&s_full[t]));
            };
            int stage = 0;
            uint32_t phase = 0;
            mbar_wait(smem_u32(&kv_full[0]), 0);
            tc_fence_after();
            issue_s(0, smem_u32(sKV));
            issue_s(1, smem_u32(sKV));
            for (int j = 0; j < ntiles; ++j) {
                const uint32_t jp = j & 1;
                const uint32_t v_addr = smem_u32(sKV + stage * 2 * AT_TILE + AT_TILE);
                int nstage = stage + 1;
                uint32_t nphase = phase;
                if (nstage == A2_KV_STAGES) {
                    nstage = 0;
                    nphase ^= 1;
                }
                const bool more = j + 1 < ntiles;
                if (more) mbar_wait(smem_u32(&kv_full[nstage]), nphase);
                for (int t = 0; t < 2; ++t) {
                    // P_j written and S_j consumed by warpgroup t: first refill S (next keys), then O_j = P_j V_j
                    mbar_wait(smem_u32(&p_full[t]), jp);
                    mbar_wait(smem_u32(&s_empty[t]), jp);
                    tc_fence_after();
                    if (more) issue_s(t, smem_u32(sKV + nstage * 2 * AT_TILE));
                    mbar_wait(smem_u32(&o_empty[t]), jp ^ 1);
                    tc_fence_after();
#pragma unroll
                    for (int k = 0; k < 8; ++k)
                        mma_f16_ts(tO + t * 64, tP + t * 64 + k * 8, smem_desc_sw128(v_addr + k * 2048, 8192, 1024),
                                   IDESC_O, k ? 1u : 0u);
                    mma_commit(smem_u32(&o_full[t]));
                }
                mma_commit(smem_u32(&kv_empty[stage]));
                stage = nstage;
                phase = nphase;
            }
        }
    } else if (warp >= 4) {
        const int q = warp & 3;
        const int t = (warp - 4) >> 2;                     // q tile of this warpgroup
        const int r = q * 32 + lane;
        const uint32_t lane_off = (uint32_t)(q * 32) << 16;
        const float sc = 0.125f * 1.4426950408889634f;
        float m = -INFINITY, l = 0.f;
        float acc[AT_D];
#pragma unroll
        for (int d = 0; d < AT_D; ++d) acc[d] = 0.f;
        float corr_prev = 1.f;
        // acc = acc * corr + O_jj, with O_jj read back from TMEM (also means P_jj has been consumed by the MMA)
        auto accumulate = [&](int jj, float corr) {
            mbar_wait(smem_u32(&o_full[t]), jj & 1);
            tc_fence_after();
#pragma unroll
            for (int c = 0; c < AT_D; c += 32) {
                if ((ABL == 1 || ABL == 5) && jj != ntiles - 1) continue;      // ablation: no per-tile read-back of O
                uint32_t v[32];
                tmem_ld32(tO + lane_off + t * 64 + c, v);
                tc_wait_ld();
                const float2 corr2 = make_float2(corr, corr);
#pragma unroll
                for (int i = 0; i < 32; i += 2) {
                    const float2 a2 = ffma2(make_float2(acc[c + i], acc[c + i + 1]), corr2,
                                            make_float2(__uint_as_float(v[i]), __uint_as_float(v[i + 1])));
                    acc[c + i] = a2.x;
                    acc[c + i + 1] = a2.y;
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&o_empty[t]));
        };
        // Online softmax with a lagging reference.  The exact scheme reads S twice per key tile (row max, then exp); here
        // tile j > 0 uses the running maximum of tiles < j as its reference, so one pass over S suffices (P may exceed 1,
        // harmless in bf16 / fp32; O and l are rescaled when the reference moves, as before).  Exactness is kept by a
        // fallback: if any row of the warp sees a score more than 2^64 above its reference, the warp redoes the tile
        // with the exact two-pass scheme (S is still in TMEM, P has not been handed to the MMA yet).
        float m_ref = 0.f;                 // reference of the previous tile's P (== reference of acc and l)
        auto exp_chunk = [&](int c, float ref, float& rs, float& mxr, uint32_t (&pk)[16]) {
            uint32_t v[32];
            tmem_ld32(tS + lane_off + t * 128 + c, v);
            tc_wait_ld();
            // pairs of scores with packed fp32 instructions (the softmax warps are issue-bound: 57 % issue-slot utilisation with
            // eight warps, ncu): x = s * sc - ref and the row sum as FFMA2 / FADD2, the row maximum as a 3-input max, every
            // A2_POLY_EVERY-th PAIR of exponentials on the FMA pipe
            const float2 sc2 = make_float2(sc, sc), nref2 = make_float2(-ref, -ref);
            float2 rs2 = make_float2(0.f, 0.f);
#pragma unroll
            for (int e = 0; e < 16; ++e) {
                const float2 s2 = make_float2(__uint_as_float(v[2 * e]), __uint_as_float(v[2 * e + 1]));
                mxr = fmaxf(fmaxf(s2.x, s2.y), mxr);
                const float2 x2 = ffma2(s2, sc2, nref2);
                float2 p2;
                if (ABL == 2) {
                    p2 = x2;
                } else if (e % POLY_EVERY == POLY_EVERY - 1) {
                    p2 = ex2_poly2(x2);
                } else {
                    p2.x = ex2_approx(x2.x);
                    p2.y = ex2_approx(x2.y);
                }
                rs2 = fadd2(rs2, p2);
                const __nv_bfloat162 h2 = __floats2bfloat162_rn(p2.x, p2.y);
                pk[e] = *reinterpret_cast<const uint32_t*>(&h2);
            }
            rs += rs2.x + rs2.y;
        };
        for (int j = 0; j < ntiles; ++j) {
            const uint32_t jp = j & 1;
            mbar_wait(smem_u32(&s_full[t]), jp);
            tc_fence_after();
            float ref = m, rs = 0.f, mxr = -INFINITY;
            bool exact = (j == 0);
            uint32_t pkA[16];
            if (j > 0) accumulate(j - 1, corr_prev);          // P_{j-1} consumed: the P columns may be overwritten
            if (ABL == 5) {                                    // ablation: synchronisation chain + MMAs only
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive(smem_u32(&s_empty[t]));
                    mbar_arrive(smem_u32(&p_full[t]));
                }
                continue;
            }
            if (!exact) {
#pragma unroll 1
                for (int c = 0; c < AT_TK; c += 32) {
                    exp_chunk(c, ref, rs, mxr, pkA);
                    tmem_st16(tP + lane_off + t * 64 + (c >> 1), pkA);
                }
                exact = __any_sync(0xffffffffu, fmaf(mxr, sc, -ref) > 64.f);
            }
            if (exact) {
                // exact two-pass scheme (first tile, or a score far above the lagging reference)
                mxr = -INFINITY;
#pragma unroll 1
                for (int c = 0; c < AT_TK; c += 32) {
                    uint32_t v[32];
                    tmem_ld32(tS + lane_off + t * 128 + c, v);
                    tc_wait_ld();
#pragma unroll
                    for (int i = 0; i < 32; ++i) mxr = fmaxf(mxr, __uint_as_float(v[i]));
                }
                ref = fmaxf(m, mxr * sc);
                rs = 0.f;
                float dummy = -INFINITY;
#pragma unroll 1
                for (int c = 0; c < AT_TK; c += 32) {
                    exp_chunk(c, ref, rs, dummy, pkA);
                    tmem_st16(tP + lane_off + t * 64 + (c >> 1), pkA);
                }
            }
            const float corr = (j == 0) ? 1.f : ex2_approx(m_ref - ref);
            corr_prev = corr;
            l = l * corr + rs;
            m_ref = ref;
            m = fmaxf(m, mxr * sc);        // running maximum over the tiles seen so far
            tc_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(smem_u32(&s_empty[t]));
                mbar_arrive(smem_u32(&p_full[t]));
            }
        }
        accumulate(ntiles - 1, corr_prev);
        const float inv = 1.f / l;
        const long long row = (long long)row_base + q0 + t * AT_TQ + r;
        __nv_bfloat16* op = p.out + row * p.C + h * AT_D;
#pragma unroll
        for (int d = 0; d < AT_D; d += 16) {
            float o16[16];
#pragma unroll
            for (int e = 0; e < 16; ++e) o16[e] = acc[d + e] * inv;
            st16(op + d, o16);
        }
        p.lse[((long long)n * p.heads + h) * p.T + q0 + t * AT_TQ + r] = (m_ref + log2f(l)) * 0.6931471805599453f;
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

// ------------------------------------------------------------------------------------------------
// forward, two query tiles per CTA, O ACCUMULATED IN TENSOR MEMORY (the default; PU_ATTN_FWD=2 selects the kernel above).  attn_fwd_tc2_kernel reads O_j back
// after every key tile and keeps the running output in 64 registers per thread (6 % of its time by ablation, and the
// reason it has no registers left to prefetch S).  Here O = sum_j P_j V_j stays in TMEM for the whole key walk (PV MMAs
// with the accumulate flag) and the softmax reference of a row only moves when the running maximum has drifted more than
// 2^8 above it (P <= 256: harmless in bf16 / fp32): then -- rarely after the first few tiles -- the warp waits for the
// pending PV MMA, rescales its 32 x 64 slice of O in place (tcgen05.ld / st) and its row sums.  The exact two-pass
// fallback for scores more than 2^64 above the reference is kept.  With the registers that frees, the next 32-column
// chunk of S is already in flight while the current one is exponentiated.
// ------------------------------------------------------------------------------------------------
template <int POLY_EVERY = A2_POLY_EVERY>
__global__ void __launch_bounds__(384, 1)
attn_fwd_tc3_kernel(const __grid_constant__ CUtensorMap tmQKV, const AttnFwdParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* sQ = smem;                                        // q tile t at + t*TILE
    uint8_t* sKV = sQ + 2 * AT_TILE;                           // stage s: K at +s*2*TILE, V at +s*2*TILE + TILE
    uint64_t* bars = reinterpret_cast<uint64_t*>(sKV + A2_KV_STAGES * 2 * AT_TILE);
    uint64_t* q_full = bars;
    uint64_t* kv_full = bars + 1;
    uint64_t* kv_empty = kv_full + A2_KV_STAGES;
    uint64_t* s_full = kv_empty + A2_KV_STAGES;    // [2] per q tile
    uint64_t* s_empty = s_full + 2;
    uint64_t* p_full = s_empty + 2;
    uint64_t* o_full = p_full + 2;                 // PV_j of q tile t has completed (P_j consumed, O up to date)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nh = blockIdx.y, n = nh / p.heads, h = nh % p.heads;
    const int q0 = blockIdx.x * 2 * AT_TQ;
    const int ntiles = p.T / AT_TK;
    const int row_base = n * p.T;
    const int colQ = h * AT_D, colK = p.C + h * AT_D, colV = 2 * p.C + h * AT_D;

    if (warp == 0 && lane == 0) prefetch_tmap(&tmQKV);
    if (warp == 1 && lane == 0) {
        mbar_init(smem_u32(q_full), 1);
        for (int s = 0; s < A2_KV_STAGES; ++s) {
            mbar_init(smem_u32(&kv_full[s]), 1);
            mbar_init(smem_u32(&kv_empty[s]), 1);
        }
        for (int t = 0; t < 2; ++t) {
            mbar_init(smem_u32(&s_full[t]), 1);
            mbar_init(smem_u32(&s_empty[t]), 4);
            mbar_init(smem_u32(&p_full[t]), 4);
            mbar_init(smem_u32(&o_full[t]), 1);
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
    const uint32_t tS = tmem_base;           // S[t] at + t*128
    const uint32_t tO = tmem_base + 256;     // O[t] at + t*64
    const uint32_t tP = tmem_base + 384;     // P[t] at + t*64 (bf16 pairs)

    if (warp == 0) {
        if (lane == 0) {
            mbar_expect_tx(smem_u32(q_full), 2 * AT_TILE);
            tma_load_2d(smem_u32(sQ), &tmQKV, smem_u32(q_full), colQ, row_base + q0);
            tma_load_2d(smem_u32(sQ + AT_TILE), &tmQKV, smem_u32(q_full), colQ, row_base + q0 + AT_TQ);
            int stage = 0;
            uint32_t phase = 0;
            for (int j = 0; j < ntiles; ++j) {
                mbar_wait(smem_u32(&kv_empty[stage]), phase ^ 1);
                const uint32_t fb = smem_u32(&kv_full[stage]);
                mbar_expect_tx(fb, 2 * AT_TILE);
                tma_load_2d(smem_u32(sKV + stage * 2 * AT_TILE), &tmQKV, fb, colK, row_base + j * AT_TK);
                tma_load_2d(smem_u32(sKV + stage * 2 * AT_TILE + AT_TILE), &tmQKV, fb, colV, row_base + j * AT_TK);
                if (++stage == A2_KV_STAGES) {
                    stage = 0;
                    phase ^= 1;
                }
            }
        }
    } else if (warp == 1) {
        constexpr uint32_t IDESC_S = idesc_bf16_f32(128, 128, 0, 0);
        constexpr uint32_t IDESC_O = idesc_bf16_f32(128, 64, 0, 1);
        mbar_wait(smem_u32(q_full), 0);
        auto issue_s = [&](int t, uint32_t k_addr) {
            const uint32_t q_addr = smem_u32(sQ + t * AT_TILE);
#pragma unroll
            for (int k = 0; k < 4; ++k)
                mma_f16_ss(tS + t * 128, smem_desc_sw128(q_addr + k * 32, 16, 1024),
                           smem_desc_sw128(k_addr + k * 32, 16, 1024), IDESC_S, k ? 1u : 0u);
            mma_commit(smem_u32(&s_full[t]));
        };
        int stage = 0;
        uint32_t phase = 0;
        mbar_wait(smem_u32(&kv_full[0]), 0);
        tc_fence_after();
        issue_s(0, smem_u32(sKV));
        issue_s(1, smem_u32(sKV));
        for (int j = 0; j < ntiles; ++j) {
            const uint32_t jp = j & 1;
            const uint32_t v_addr = smem_u32(sKV + stage * 2 * AT_TILE + AT_TILE);
            int nstage = stage + 1;
            uint32_t nphase = phase;
            if (nstage == A2_KV_STAGES) {
                nstage = 0;
                nphase ^= 1;
            }
            const bool more = j + 1 < ntiles;
            if (more) mbar_wait(smem_u32(&kv_full[nstage]), nphase);
            for (int t = 0; t < 2; ++t) {
                // P_j written (and O rescaled if the reference moved), S_j consumed: refill S, then O += P_j V_j
                mbar_wait(smem_u32(&p_full[t]), jp);
                mbar_wait(smem_u32(&s_empty[t]), jp);
                tc_fence_after();
                if (more) issue_s(t, smem_u32(sKV + nstage * 2 * AT_TILE));
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    mma_f16_ts(tO + t * 64, tP + t * 64 + k * 8, smem_desc_sw128(v_addr + k * 2048, 8192, 1024),
                               IDESC_O, (j | k) ? 1u : 0u);
                mma_commit(smem_u32(&o_full[t]));
            }
            mma_commit(smem_u32(&kv_empty[stage]));
            stage = nstage;
            phase = nphase;
        }
    } else if (warp >= 4) {
        const int q = warp & 3;
        const int t = (warp - 4) >> 2;                     // q tile of this warpgroup
        const int r = q * 32 + lane;
        const uint32_t lane_off = (uint32_t)(q * 32) << 16;
        const float sc = 0.125f * 1.4426950408889634f;
        float m = -INFINITY, l = 0.f;      // running maximum (scaled, log2 domain) and row sum relative to m_ref
        float m_ref = 0.f;                 // reference of O, l and of the P tiles written so far
        float pend_ref = 0.f;              // reference to move to once the pending PV MMA has completed
        bool pending = false;              // warp-uniform
        // exponentials of one preloaded 32-column chunk: bf16 pairs into pk, row sum / row maximum updated
        auto exp_chunk = [&](const uint32_t (&v)[32], float ref, float& rs, float& mxr, uint32_t (&pk)[16]) {
            const float2 sc2 = make_float2(sc, sc), nref2 = make_float2(-ref, -ref);
            float2 rs2 = make_float2(0.f, 0.f);
#pragma unroll
            for (int e = 0; e < 16; ++e) {
                const float2 s2 = make_float2(__uint_as_float(v[2 * e]), __uint_as_float(v[2 * e + 1]));
                mxr = fmaxf(fmaxf(s2.x, s2.y), mxr);
                const float2 x2 = ffma2(s2, sc2, nref2);
                float2 p2;
                if (POLY_EVERY > 0 && e % POLY_EVERY == POLY_EVERY - 1) {
                    p2 = ex2_poly2(x2);
                } else {
                    p2.x = ex2_approx(x2.x);
                    p2.y = ex2_approx(x2.y);
                }
                rs2 = fadd2(rs2, p2);
                const __nv_bfloat162 h2 = __floats2bfloat162_rn(p2.x, p2.y);
                pk[e] = *reinterpret_cast<const uint32_t*>(&h2);
            }
            rs += rs2.x + rs2.y;
        };
        // O[32 x 64 slice of this warp] *= corr (per row), in tensor memory
        auto rescale_o = [&](float corr) {
#pragma unroll
            for (int c = 0; c < AT_D; c += 32) {
                uint32_t v[32];
                tmem_ld32(tO + lane_off + t * 64 + c, v);
                tc_wait_ld();
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * corr);
                tmem_st32(tO + lane_off + t * 64 + c, v);
            }
        };
        // one pass over S_j with reference `ref`: P_j -> TMEM; the next chunk's tcgen05.ld is in flight during the math
        auto pass = [&](int j, float ref, float& rs, float& mxr, bool& o_waited) {
            uint32_t va[32], vb[32];
            tmem_ld32(tS + lane_off + t * 128, va);
            tc_wait_ld();
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                uint32_t pk[16];
                if (c + 1 < 4) {
                    if (c & 1) tmem_ld32(tS + lane_off + t * 128 + (c + 1) * 32, va);
                    else tmem_ld32(tS + lane_off + t * 128 + (c + 1) * 32, vb);
                }
                if (c & 1) exp_chunk(vb, ref, rs, mxr, pk);
                else exp_chunk(va, ref, rs, mxr, pk);
                if (c == 0 && j > 0 && !o_waited) {
                    // PV_{j-1} has consumed P_{j-1}: the P columns may be overwritten
                    mbar_wait(smem_u32(&o_full[t]), (j - 1) & 1);
                    tc_fence_after();
                    o_waited = true;
                }
                tmem_st16(tP + lane_off + t * 64 + c * 16, pk);
                if (c + 1 < 4) tc_wait_ld();
            }
        };
        for (int j = 0; j < ntiles; ++j) {
            const uint32_t jp = j & 1;
            mbar_wait(smem_u32(&s_full[t]), jp);
            tc_fence_after();
            bool o_waited = false;
            if (pending) {
                // the reference moves now: O (which includes PV_{j-1} once o_full fires) and l are rescaled first
                mbar_wait(smem_u32(&o_full[t]), (j - 1) & 1);
                tc_fence_after();
                o_waited = true;
                const float corr = ex2_approx(m_ref - pend_ref);
                rescale_o(corr);
                l *= corr;
                m_ref = pend_ref;
                pending = false;
            }
            float rs = 0.f, mxr = -INFINITY;
            bool exact = (j == 0);
            if (!exact) {
                pass(j, m_ref, rs, mxr, o_waited);
                exact = __any_sync(0xffffffffu, fmaf(mxr, sc, -m_ref) > 64.f);
            }
            if (exact) {
                // exact two-pass scheme (first tile, or a score far above the reference): row maximum first
                mxr = -INFINITY;
#pragma unroll 1
                for (int c = 0; c < AT_TK; c += 32) {
                    uint32_t v[32];
                    tmem_ld32(tS + lane_off + t * 128 + c, v);
                    tc_wait_ld();
#pragma unroll
                    for (int i = 0; i < 32; ++i) mxr = fmaxf(mxr, __uint_as_float(v[i]));
                }
                const float ref = fmaxf(m, mxr * sc);
                if (j > 0) {
                    if (!o_waited) {
                        mbar_wait(smem_u32(&o_full[t]), (j - 1) & 1);
                        tc_fence_after();
                        o_waited = true;
                    }
                    const float corr = ex2_approx(m_ref - ref);
                    rescale_o(corr);
                    l *= corr;
                }
                m_ref = ref;
                rs = 0.f;
                float dummy = -INFINITY;
                pass(j, m_ref, rs, dummy, o_waited);
            }
            l += rs;
            m = fmaxf(m, mxr * sc);
            // lazy reference: move it (before the next tile) only when the maximum has drifted 2^8 above it
            if (__any_sync(0xffffffffu, m - m_ref > 8.f)) {
                pending = true;
                pend_ref = m;
            }
            tc_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(smem_u32(&s_empty[t]));
                mbar_arrive(smem_u32(&p_full[t]));
            }
        }
        // final O: wait for the last PV MMA, normalise, store
        mbar_wait(smem_u32(&o_full[t]), (ntiles - 1) & 1);
        tc_fence_after();
        const float inv = 1.f / l;
        const long long row = (long long)row_base + q0 + t * AT_TQ + r;
        __nv_bfloat16* op = p.out + row * p.C + h * AT_D;
#pragma unroll
        for (int c = 0; c < AT_D; c += 32) {
            uint32_t v[32];
            tmem_ld32(tO + lane_off + t * 64 + c, v);
            tc_wait_ld();
#pragma unroll
            for (int d = 0; d < 32; d += 16) {
                float o16[16];
#pragma unroll
                for (int e = 0; e < 16; ++e) o16[e] = __uint_as_float(v[d + e]) * inv;
                st16(op + c + d, o16);
            }
        }
        p.lse[((long long)n * p.heads + h) * p.T + q0 + t * AT_TQ + r] = (m_ref + log2f(l)) * 0.6931471805599453f;
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
    PU_SMEM_ATTR(attn_fwd_tc_kernel, AT_SMEM);
    AttnFwdParams p;
    p.T = T; p.heads = heads; p.C = C;
    p.out = (__nv_bfloat16*)out;
    p.lse = lse;
    if (T % (2 * AT_TQ) == 0) {
        dim3 grid2(T / (2 * AT_TQ), N * heads);
        // PU_ATTN_FWD: 3 (default) = O accumulated in tensor memory, 1.46 ms at T = 4096, heads 4, batch 64; 2 = the kernel
        // that reads O back per key tile (1.71 ms; kept for A/B runs and the ablation switches)
        static const int fwd_variant = getenv("PU_ATTN_FWD") ? atoi(getenv("PU_ATTN_FWD")) : 3;
        if (fwd_variant == 3 && !getenv("PU_ATTN_FWD_ABL") && !getenv("PU_ATTN_FWD_POLY")) {
            // every POLY-th pair of exponentials on the FMA pipe (ex2_poly2); measured at T = 4096: none 1.574 ms, every 2nd
            // 1.467, 3rd 1.424, 4th 1.454, 6th 1.436, 8th 1.466
            static const int poly3 = getenv("PU_ATTN_FWD3_POLY") ? atoi(getenv("PU_ATTN_FWD3_POLY")) : 3;
            if (poly3 == 0) {
                PU_SMEM_ATTR(attn_fwd_tc3_kernel<0>, A2_SMEM);
                attn_fwd_tc3_kernel<0><<<grid2, 384, A2_SMEM, st>>>(tm, p);
            } else if (poly3 == 4) {
                PU_SMEM_ATTR(attn_fwd_tc3_kernel<4>, A2_SMEM);
                attn_fwd_tc3_kernel<4><<<grid2, 384, A2_SMEM, st>>>(tm, p);
            } else {
                PU_SMEM_ATTR(attn_fwd_tc3_kernel<3>, A2_SMEM);
                attn_fwd_tc3_kernel<3><<<grid2, 384, A2_SMEM, st>>>(tm, p);
            }
            return check_launch("attn_fwd_tc3");
        }
        static const int abl = getenv("PU_ATTN_FWD_ABL") ? atoi(getenv("PU_ATTN_FWD_ABL")) : 0;
        if (abl == 1) {
            PU_SMEM_ATTR(attn_fwd_tc2_kernel<1>, A2_SMEM);
            attn_fwd_tc2_kernel<1><<<grid2, 384, A2_SMEM, st>>>(tm, p);
        } else if (abl == 2) {
            PU_SMEM_ATTR(attn_fwd_tc2_kernel<2>, A2_SMEM);
            attn_fwd_tc2_kernel<2><<<grid2, 384, A2_SMEM, st>>>(tm, p);
        } else if (abl == 5) {
            PU_SMEM_ATTR(attn_fwd_tc2_kernel<5>, A2_SMEM);
            attn_fwd_tc2_kernel<5><<<grid2, 384, A2_SMEM, st>>>(tm, p);
        } else {
            static const int poly = getenv("PU_ATTN_FWD_POLY") ? atoi(getenv("PU_ATTN_FWD_POLY")) : A2_POLY_EVERY;
            if (poly == 1) {
                PU_SMEM_ATTR((attn_fwd_tc2_kernel<0, 1>), A2_SMEM);
                attn_fwd_tc2_kernel<0, 1><<<grid2, 384, A2_SMEM, st>>>(tm, p);
            } else if (poly == 2) {
                PU_SMEM_ATTR((attn_fwd_tc2_kernel<0, 2>), A2_SMEM);
                attn_fwd_tc2_kernel<0, 2><<<grid2, 384, A2_SMEM, st>>>(tm, p);
            } else if (poly == 3) {
                PU_SMEM_ATTR((attn_fwd_tc2_kernel<0, 3>), A2_SMEM);
                attn_fwd_tc2_kernel<0, 3><<<grid2, 384, A2_SMEM, st>>>(tm, p);
            } else {
                PU_SMEM_ATTR((attn_fwd_tc2_kernel<0, 4>), A2_SMEM);
                attn_fwd_tc2_kernel<0, 4><<<grid2, 384, A2_SMEM, st>>>(tm, p);
            }
        }
        return check_launch("attn_fwd_tc2");
    }
    dim3 grid(T / AT_TQ, N * heads);
    attn_fwd_tc_kernel<<<grid, 256, AT_SMEM, st>>>(tm, p);
    return check_launch("attn_fwd_tc");
}

// ------------------------------------------------------------------------------------------------
// backward: one CTA per (key tile of 128, sample*head), walking the query tiles.
//   S  = Q_i K_j^T, dP = dO_i V_j^T                      (K-major operands)            -> TMEM
//   P  = exp2(S*sc - lse_i), dS = P (dP - delta_i) / 8   (softmax warps)               -> shared (SW128, [q][k])
//   dV += P^T dO_i, dK += dS^T Q_i                       (P/dS and dO/Q as MN-major operands, accumulate in TMEM)
//   dQ_i = dS K_j                                        (dS K-major, K_j MN-major)    -> TMEM -> red.add.v4 (fp32)
// dQ partial sums of the different key tiles are combined with vectorised fp32 reductions into dq_acc [N*T][C].
// ------------------------------------------------------------------------------------------------
struct AttnBwdParams {
    int T, heads, C;
    const float* lse;
    const float* delta;
    float* dq_acc;            // [N*T][C] fp32, zeroed by the caller
    __nv_bfloat16* dqkv;      // [N*T][3C]
    float* dbias;             // optional [3C], zeroed by the caller: column sums of dqkv (= gradient of the qkv conv's bias)
    float* dbias_part;        // optional [N*heads][T/128][192]: per-CTA column sums (dq | dk | dv) instead of atomics on `dbias`
};

__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// ------------------------------------------------------------------------------------------------
// backward kernel.  S / dP are produced as two 64-key halves with their own barriers and P / dS live in two
// shared-memory buffers (tile parity), so the tensor pipe always runs one unit ahead of the softmax warps:
//   MMA warp, tile i :  [P/dS half 0 ready] dQ_i  = dS_h0 K_h0 ; S/dP half 0 of tile i+1
//                       [P/dS half 1 ready] dQ_i += dS_h1 K_h1 ; dV += P^T dO ; dK += dS^T Q ; S/dP half 1 of tile i+1
//   softmax warps    :  half 0 of tile i ; half 1 of tile i            (warps 4-11, nothing else)
//   dQ drain warps   :  dQ_i : TMEM -> registers -> red.global.add.v4.f32   (warps 12-15, double-buffered accumulator)
// The softmax warps never wait for an MMA that was not issued at least half a tile earlier.  The dQ partials used to be
// drained by the softmax warps themselves between the two halves; the 32 KB of fp32 reductions per tile (bound by the
// L2 atomic rate, ~1500 clk per tile and SM) then sat on the critical path next to the ~2000 clk of softmax work.  With
// their own warpgroup and two dQ accumulators in the 64 spare TMEM columns, reductions, softmax and MMAs overlap.
// TMEM map (512 columns): S 0-127 | dP 128-255 | dV 256-319 | dK 320-383 | dQ[0] 384-447 | dQ[1] 448-511.
// ------------------------------------------------------------------------------------------------
constexpr int AB2_SMEM = 2 * AT_TILE /*K,V*/ + 2 * 2 * AT_TILE /*Q,dO x2 stages*/ + 2 * 2 * AT_TILE /*P x2*/ +
                         2 * 2 * AT_TILE /*dS x2*/ + 1024 + 256;

constexpr int AB2_THREADS = 512;

// register reallocation between warpgroups (sm_90+): a warpgroup gives registers back / takes more than the launch gave it
template <int N>
__device__ __forceinline__ void setmaxnreg_dec() {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N));
}
template <int N>
__device__ __forceinline__ void setmaxnreg_inc() {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N));
}

// POLY: every POLY-th exponential of a thread is evaluated on the FMA pipe (ex2_poly) instead of MUFU.EX2; 0 = none.
// MUFU instructions share the MIO queue with the STS.128 of P / dS, which is where the softmax warps stall.
// SW: softmax warps.  8: both 64-key halves of a tile are processed one after the other by the same two warpgroups
// (512 threads).  16: each half has its own two warpgroups (768 threads; the control and drain warpgroups hand registers
// to the softmax warpgroups with setmaxnreg) -- the softmax unit is latency-bound (28 % of the issue slots used with two
// warps per scheduler), so twice the warps in flight overlap two units instead of serialising them.
template <int POLY, int SW = 8, int ABL = 0>
__global__ void __launch_bounds__(256 + SW * 32, 1)
attn_bwd_tc2_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmDO,
                    const AttnBwdParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // 1024-aligned; an OFFSET from the
    // __shared__ symbol, so that plain C++ accesses below compile to LDS / STS instead of generic LD / ST
    uint8_t* sK = smem;
    uint8_t* sV = sK + AT_TILE;
    uint8_t* sQD = sV + AT_TILE;                           // stage s: Q at +s*2*TILE, dO at +s*2*TILE + TILE
    uint8_t* sP = sQD + 2 * 2 * AT_TILE;                   // buffer b at +b*2*TILE: two 64-key blocks of [128 q][128 B]
    uint8_t* sDS = sP + 2 * 2 * AT_TILE;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sDS + 2 * 2 * AT_TILE);
    uint64_t* kv_full = bars;
    uint64_t* qd_full = bars + 1;                          // [2]
    uint64_t* qd_empty = qd_full + 2;                      // [2]
    uint64_t* sdp_full = qd_empty + 2;                     // [2] per key half
    uint64_t* pds_full = sdp_full + 2;                     // [2] per key half
    uint64_t* pds_free = pds_full + 2;                     // [2] per P/dS buffer
    uint64_t* dq_full = pds_free + 2;                      // [2] per dQ accumulator
    uint64_t* dq_empty = dq_full + 2;                      // [2]
    uint64_t* fin = dq_empty + 2;                          // every MMA of the CTA has completed (dK / dV final)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(fin + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nh = blockIdx.y, n = nh / p.heads, h = nh % p.heads;
    const int k0 = blockIdx.x * AT_TK;
    const int nq = p.T / AT_TQ;
    const int row_base = n * p.T;
    const int colQ = h * AT_D, colK = p.C + h * AT_D, colV = 2 * p.C + h * AT_D;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmQKV);
        prefetch_tmap(&tmDO);
    }
    if (warp == 1 && lane == 0) {
        mbar_init(smem_u32(kv_full), 1);
        for (int s = 0; s < 2; ++s) {
            mbar_init(smem_u32(&qd_full[s]), 1);
            mbar_init(smem_u32(&qd_empty[s]), 1);
            mbar_init(smem_u32(&sdp_full[s]), 1);
            mbar_init(smem_u32(&pds_full[s]), 8);    // one arrive per softmax warp of the half (SW = 16: the half's own 8)
            mbar_init(smem_u32(&pds_free[s]), 1);
            mbar_init(smem_u32(&dq_full[s]), 1);
            mbar_init(smem_u32(&dq_empty[s]), 4);    // one arrive per drain warp
        }
        mbar_init(smem_u32(fin), 1);
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
    const uint32_t tS = tmem_base, tDP = tmem_base + 128, tDV = tmem_base + 256, tDK = tmem_base + 320,
                   tDQ = tmem_base + 384;

    // SW = 16: 768 threads x 80 registers at launch; the control warpgroup (warps 0-3) keeps 56, the drain warpgroup 40, the
    // four softmax warpgroups take 96 each (128 x 56 + 128 x 40 + 512 x 96 = 61440).  setmaxnreg is the first
    // statement of every role branch so that each role's code is compiled against its own register budget.
    if (warp < 4) {
        if constexpr (SW == 16) setmaxnreg_dec<56>();
    }
    if (warp == 0) {
        if (lane == 0) {
            mbar_expect_tx(smem_u32(kv_full), 2 * AT_TILE);
            tma_load_2d(smem_u32(sK), &tmQKV, smem_u32(kv_full), colK, row_base + k0);
            tma_load_2d(smem_u32(sV), &tmQKV, smem_u32(kv_full), colV, row_base + k0);
            for (int i = 0; i < nq; ++i) {
                const int stage = i & 1;
                mbar_wait(smem_u32(&qd_empty[stage]), ((i >> 1) & 1) ^ 1);
                const uint32_t fb = smem_u32(&qd_full[stage]);
                mbar_expect_tx(fb, 2 * AT_TILE);
                // q tiles are walked starting at this CTA's own key-tile index: concurrently running CTAs of one
                // (sample, head) then reduce into different dQ tiles instead of contending for the same addresses
                const int qi = (i + (int)blockIdx.x) % nq;
                tma_load_2d(smem_u32(sQD + stage * 2 * AT_TILE), &tmQKV, fb, colQ, row_base + qi * AT_TQ);
                tma_load_2d(smem_u32(sQD + stage * 2 * AT_TILE + AT_TILE), &tmDO, fb, colQ, row_base + qi * AT_TQ);
            }
        }
    } else if (warp == 1) {
        // whole warp, warp-uniform control flow; mma_f16_ss / mma_commit elect one lane internally
        constexpr uint32_t IDESC_H = idesc_bf16_f32(128, 64, 0, 0);      // S, dP halves (64 keys)
        constexpr uint32_t IDESC_T = idesc_bf16_f32(128, 64, 1, 1);      // dV, dK (both operands MN-major)
        constexpr uint32_t IDESC_Q = idesc_bf16_f32(128, 64, 0, 1);      // dQ
        const uint32_t k_addr = smem_u32(sK), v_addr = smem_u32(sV);
        mbar_wait(smem_u32(kv_full), 0);
        auto issue_sdp = [&](int stage, int half) {
            const uint32_t q_addr = smem_u32(sQD + stage * 2 * AT_TILE);
            const uint32_t do_addr = q_addr + AT_TILE;
            const uint32_t kh = k_addr + half * (AT_TILE / 2), vh = v_addr + half * (AT_TILE / 2);
#pragma unroll
            for (int k = 0; k < 4; ++k)
                mma_f16_ss(tS + half * 64, smem_desc_sw128(q_addr + k * 32, 16, 1024),
                           smem_desc_sw128(kh + k * 32, 16, 1024), IDESC_H, k ? 1u : 0u);
#pragma unroll
            for (int k = 0; k < 4; ++k)
                mma_f16_ss(tDP + half * 64, smem_desc_sw128(do_addr + k * 32, 16, 1024),
                           smem_desc_sw128(vh + k * 32, 16, 1024), IDESC_H, k ? 1u : 0u);
            mma_commit(smem_u32(&sdp_full[half]));
        };
        mbar_wait(smem_u32(&qd_full[0]), 0);
        tc_fence_after();
        issue_sdp(0, 0);
        issue_sdp(0, 1);
        for (int i = 0; i < nq; ++i) {
            const int stage = i & 1;
            const uint32_t q_addr = smem_u32(sQD + stage * 2 * AT_TILE);
            const uint32_t do_addr = q_addr + AT_TILE;
            const uint32_t p_addr = smem_u32(sP + stage * 2 * AT_TILE), ds_addr = smem_u32(sDS + stage * 2 * AT_TILE);
            const uint32_t tDQb = tDQ + (uint32_t)(i & 1) * 64;      // dQ accumulator of this tile
            // ---- key half 0 ----
            mbar_wait(smem_u32(&pds_full[0]), i & 1);
            if (i >= 2) mbar_wait(smem_u32(&dq_empty[i & 1]), ((i >> 1) - 1) & 1);   // tile i-2 drained from this buffer
            tc_fence_after();
#pragma unroll
            for (int k = 0; k < 4; ++k)
                mma_f16_ss(tDQb, smem_desc_sw128(ds_addr + k * 32, 16, 1024),
                           smem_desc_sw128(k_addr + k * 2048, AT_TILE, 1024), IDESC_Q, k ? 1u : 0u);
            if (i + 1 < nq) {
                mbar_wait(smem_u32(&qd_full[stage ^ 1]), ((i + 1) >> 1) & 1);
                tc_fence_after();
                issue_sdp(stage ^ 1, 0);
            }
            // ---- key half 1 ----
            mbar_wait(smem_u32(&pds_full[1]), i & 1);
            tc_fence_after();
#pragma unroll
            for (int k = 0; k < 4; ++k)
                mma_f16_ss(tDQb, smem_desc_sw128(ds_addr + AT_TILE + k * 32, 16, 1024),
                           smem_desc_sw128(k_addr + (4 + k) * 2048, AT_TILE, 1024), IDESC_Q, 1u);
#pragma unroll
            for (int k = 0; k < 8; ++k)     // K = 128 query rows, 16 per step
                mma_f16_ss(tDV, smem_desc_sw128(p_addr + k * 2048, AT_TILE, 1024),
                           smem_desc_sw128(do_addr + k * 2048, AT_TILE, 1024), IDESC_T, (i | k) ? 1u : 0u);
#pragma unroll
            for (int k = 0; k < 8; ++k)
                mma_f16_ss(tDK, smem_desc_sw128(ds_addr + k * 2048, AT_TILE, 1024),
                           smem_desc_sw128(q_addr + k * 2048, AT_TILE, 1024), IDESC_T, (i | k) ? 1u : 0u);
            mma_commit(smem_u32(&dq_full[i & 1]));
            mma_commit(smem_u32(&qd_empty[stage]));
            mma_commit(smem_u32(&pds_free[stage]));
            if (i + 1 < nq) issue_sdp(stage ^ 1, 1);
        }
        mma_commit(smem_u32(fin));
    } else if (warp >= 4 && warp < 8) {
        if constexpr (SW == 16) setmaxnreg_dec<40>();
        // dQ drain warpgroup: warp quadrant q owns query rows (= TMEM lanes) [32q, 32q+32), all 64 feature columns.
        // dQ partial of tile i: TMEM -> registers -> vectorised fp32 reductions at L2 (bulk reductions by the TMA engine
        // and fragment-layout loads were measured and did not help: the L2 atomic rate is the limit either way).
        const int q = warp & 3;
        const int r = q * 32 + lane;
        const uint32_t lane_off = (uint32_t)(q * 32) << 16;
        for (int i = 0; i < nq; ++i) {
            const int b = i & 1;
            const int qi = (i + (int)blockIdx.x) % nq;     // same rotation as the TMA producer
            mbar_wait(smem_u32(&dq_full[b]), (i >> 1) & 1);
            tc_fence_after();
            float* dst = p.dq_acc + ((long long)row_base + qi * AT_TQ + r) * p.C + h * AT_D;
#pragma unroll 1
            for (int c = 0; c < AT_D; c += 16) {           // 16 columns at a time: the warpgroup runs on 40 registers
                uint32_t v0[16];
                tmem_ld16(tDQ + b * 64 + lane_off + c, v0);
                tc_wait_ld();
                if (c == AT_D - 16) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(smem_u32(&dq_empty[b]));
                }
                if (ABL & 4) {
                    if (__uint_as_float(v0[0]) == 1.2345e-30f) red_add_v4(dst + c, 0.f, 0.f, 0.f, 0.f);     // ablation: keep the loads alive
                    continue;
                }
#pragma unroll
                for (int e = 0; e < 16; e += 4)
                    red_add_v4(dst + c + e, __uint_as_float(v0[e]), __uint_as_float(v0[e + 1]), __uint_as_float(v0[e + 2]),
                               __uint_as_float(v0[e + 3]));
            }
        }
    } else if (warp >= 8) {
        if constexpr (SW == 16) setmaxnreg_inc<96>();
        // softmax warps: warp quadrant q owns query rows (= TMEM lanes) [32q, 32q+32); within a 64-key half warpgroup wg
        // handles key columns [32wg, 32wg+32) and, for dK / dV, feature columns [32wg, 32wg+32).  SW = 16: warpgroups 0-1
        // own key half 0, warpgroups 2-3 key half 1.
        const int q = warp & 3;
        const int wgi = (warp - 8) >> 2;
        const int wg = wgi & 1;
        const int own_half = wgi >> 1;               // SW = 16 only
        const int r = q * 32 + lane;
        const uint32_t lane_off = (uint32_t)(q * 32) << 16;
        const float sc = 0.125f * 1.4426950408889634f;
        const float* lse = p.lse + ((long long)n * p.heads + h) * p.T;
        const float* delta = p.delta + ((long long)n * p.heads + h) * p.T;
        float l2 = 0.f, dl = 0.f;                 // lse * log2(e) and delta / sqrt(d) of this thread's query row
        auto unit = [&](int i, int half) {
            const int b = i & 1;
            mbar_wait(smem_u32(&sdp_full[half]), i & 1);
            tc_fence_after();
            const int c = half * 64 + wg * 32;
            uint8_t* pb = sP + b * 2 * AT_TILE + half * AT_TILE + r * 128;
            uint8_t* db = sDS + b * 2 * AT_TILE + half * AT_TILE + r * 128;
            const int chunk0 = wg * 4;
            if constexpr (SW == 16) {
                // 16 columns at a time: this warpgroup runs on 104 registers
                mbar_wait(smem_u32(&pds_free[b]), ((i >> 1) & 1) ^ 1);
#pragma unroll
                for (int sub = 0; sub < 2; ++sub) {
                    uint32_t sv[16], dv[16];
                    tmem_ld16(tS + lane_off + c + sub * 16, sv);
                    tmem_ld16(tDP + lane_off + c + sub * 16, dv);
                    tc_wait_ld();
#pragma unroll
                    for (int g = 0; g < 2; ++g) {
                        uint4 pk, dk;
                        __nv_bfloat162* hp = reinterpret_cast<__nv_bfloat162*>(&pk);
                        __nv_bfloat162* hd = reinterpret_cast<__nv_bfloat162*>(&dk);
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const int i0 = g * 8 + 2 * e;
                            const float p0 = ex2_approx(fmaf(__uint_as_float(sv[i0]), sc, -l2));
                            const float p1 = ex2_approx(fmaf(__uint_as_float(sv[i0 + 1]), sc, -l2));
                            hp[e] = __floats2bfloat162_rn(p0, p1);
                            hd[e] = __floats2bfloat162_rn(p0 * fmaf(__uint_as_float(dv[i0]), 0.125f, -dl),
                                                          p1 * fmaf(__uint_as_float(dv[i0 + 1]), 0.125f, -dl));
                        }
                        const int off = ((chunk0 + sub * 2 + g) ^ (r & 7)) << 4;
                        *reinterpret_cast<uint4*>(pb + off) = pk;
                        *reinterpret_cast<uint4*>(db + off) = dk;
                    }
                }
                tc_fence_before();
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(&pds_full[half]));
                return;
            }
            if (ABL & 8) {     // ablation: synchronisation chain + MMAs only
                if (half == 0) mbar_wait(smem_u32(&pds_free[b]), ((i >> 1) & 1) ^ 1);
                tc_fence_before();
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(&pds_full[half]));
                return;
            }
            uint32_t sv[32], dv[32];
            tmem_ld32(tS + lane_off + c, sv);
            tmem_ld32(tDP + lane_off + c, dv);
            // the MMAs of tile i-2 must have finished reading this P / dS buffer
            if (half == 0) mbar_wait(smem_u32(&pds_free[b]), ((i >> 1) & 1) ^ 1);
            tc_wait_ld();
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                uint4 pk, dk;
                __nv_bfloat162* hp = reinterpret_cast<__nv_bfloat162*>(&pk);
                __nv_bfloat162* hd = reinterpret_cast<__nv_bfloat162*>(&dk);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int i0 = g * 8 + 2 * e;
                    const float x0 = fmaf(__uint_as_float(sv[i0]), sc, -l2), x1 = fmaf(__uint_as_float(sv[i0 + 1]), sc, -l2);
                    const float p0 = (ABL & 2) ? x0 : (POLY > 0 && (i0 % POLY) == POLY - 1) ? ex2_poly(x0) : ex2_approx(x0);
                    const float p1 = (ABL & 2) ? x1 : (POLY > 0 && ((i0 + 1) % POLY) == POLY - 1) ? ex2_poly(x1) : ex2_approx(x1);
                    const float d0 = p0 * fmaf(__uint_as_float(dv[i0]), 0.125f, -dl);
                    const float d1 = p1 * fmaf(__uint_as_float(dv[i0 + 1]), 0.125f, -dl);
                    hp[e] = __floats2bfloat162_rn(p0, p1);
                    hd[e] = __floats2bfloat162_rn(d0, d1);
                }
                const int off = ((chunk0 + g) ^ (r & 7)) << 4;
                if (ABL & 1) {     // ablation: no STS (keep the math alive with a never-true store)
                    if (pk.x == 0x12345678u && dk.y == 0x9abcdef0u) *reinterpret_cast<uint4*>(pb + off) = pk;
                    continue;
                }
                *reinterpret_cast<uint4*>(pb + off) = pk;
                *reinterpret_cast<uint4*>(db + off) = dk;
            }
            tc_fence_before();
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&pds_full[half]));
        };
        for (int i = 0; i < nq; ++i) {
            const int qi = (i + (int)blockIdx.x) % nq;     // same rotation as the TMA producer
            l2 = lse[qi * AT_TQ + r] * 1.4426950408889634f;
            dl = 0.125f * delta[qi * AT_TQ + r];
            if constexpr (SW == 16) {
                unit(i, own_half);
            } else {
                unit(i, 0);
                unit(i, 1);
            }
        }
        // final dV / dK: wait until every MMA of this CTA has completed
        mbar_wait(smem_u32(fin), 0);
        tc_fence_after();
        __nv_bfloat16* kp = p.dqkv + ((long long)row_base + k0 + r) * 3 * p.C + colK;
        __nv_bfloat16* vp = p.dqkv + ((long long)row_base + k0 + r) * 3 * p.C + colV;
#pragma unroll 1
        for (int which = 0; which < 2; ++which) {          // 0: dK, 1: dV  (SW = 16: the warpgroups of half `which` only)
            if (SW == 16 && which != own_half) continue;
            const int c = wg * 32;
            uint32_t a[32];
            tmem_ld32((which == 0 ? tDK : tDV) + lane_off + c, a);
            tc_wait_ld();
            __nv_bfloat16* op = which == 0 ? kp : vp;
#pragma unroll
            for (int e = 0; e < 32; e += 16) {
                float ka[16];
#pragma unroll
                for (int u = 0; u < 16; ++u) ka[u] = __uint_as_float(a[e + u]);
                st16(op + c + e, ka);
            }
            if (p.dbias) {
                // bias gradient of the qkv conv: column sums over this warp's 32 keys, one fp32 atomic per lane
                const float tot = warp_reduce_scatter32([&](int j) { return __uint_as_float(a[j]); }, lane);
                atomicAdd(p.dbias + (which == 0 ? colK : colV) + c + lane, tot);
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

// ------------------------------------------------------------------------------------------------
// backward kernel, transposed scores (round 2; with BULK = 1 the default).  ncu on attn_bwd_tc2_kernel
// (profiles/r2_ncu_attention_tc2_top_stalls.tsv): the eight softmax warps spend 31 % of their time in MIO throttle, 19 %
// waiting for MUFU results behind the same queue, and the tensor core pulls 224 KB of operands per (q tile, key tile) out
// of shared memory.  (What filled that queue were the drain warps' per-lane reductions: see BULK below.)
// Here the CTA computes the TRANSPOSED tiles
//     S^T = K_j Q_i^T ,  dP^T = V_j dO_i^T            (lanes = keys, columns = queries; two 64-query halves)
// so that P^T and dS^T are born in the layout the dV / dK accumulations want as their A operand:
//     dV += P^T dO_i ,  dK += dS^T Q_i                 A straight from TENSOR MEMORY (bf16 pairs written over the first
//                                                     half of the S^T / dP^T columns each warp has just read)
//     dQ_i = dS K_j                                    A = dS^T from shared memory, read as an MN-major operand
// Only dS^T goes through shared memory (4 instead of 8 STS.128 per thread and unit) and the tensor core reads 144 KB
// per tile pair.  lse_i / delta_i are per-COLUMN constants now: the TMA producer brings the tile's 2 x 512 bytes along
// with Q_i / dO_i and the softmax threads read them with broadcast LDS.128.  The 1/sqrt(d) of dS is applied once to the
// dQ partial in the drain warps and to dK at the end (both are linear in dS) instead of per element.
// Pipeline, barriers and TMEM map follow attn_bwd_tc2_kernel: S^T 0-127 | dP^T 128-255 | dV 256-319 | dK 320-383 |
// dQ[0] 384-447 | dQ[1] 448-511;  P^T half h, query block w (32 queries) sits in S^T columns [64h + 32w, +16),
// dS^T in the same columns of dP^T.
// ------------------------------------------------------------------------------------------------
constexpr int AB3_LD_BYTES = 2 * AT_TQ * 4;                     // lse_i + delta_i of one q tile
constexpr int AB3_SMEM = 2 * AT_TILE /*K,V*/ + 2 * 2 * AT_TILE /*Q,dO x2 stages*/ + 2 * 2 * AT_TILE /*dS^T x2*/ +
                         2 * AB3_LD_BYTES + 1024 + 256;
constexpr int AB3_STG_BYTES = AT_TQ * AT_D * 4;                 // dQ staging tile: [16 column vectors][128 rows] x 16 B
constexpr int AB3_SMEM_BULK = AB3_SMEM + AB3_STG_BYTES;

__device__ __forceinline__ void bulk_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}

// BULK = 1: the dQ partial of a tile leaves the SM as ONE 32 KB cp.reduce.async.bulk (add.f32) issued by the TMA engine from a
// shared-memory staging buffer instead of 2048 per-lane red.global.add.v4 -- the reductions then do not queue in the LSU
// / MIO path that the softmax warps' STS and MUFU instructions share.  The dQ workspace is tile-major for this:
// [sample*head][q tile][16 column vectors][128 rows] x 16 B (the staging layout: conflict-free STS.128), un-permuted by
// attn_dq_convert_tiles_kernel.
template <int ABL = 0, int BULK = 0>
__global__ void __launch_bounds__(AB2_THREADS, 1)
attn_bwd_tc3_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmDO,
                    const AttnBwdParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* sK = smem;
    uint8_t* sV = sK + AT_TILE;
    uint8_t* sQD = sV + AT_TILE;                           // stage s: Q at +s*2*TILE, dO at +s*2*TILE + TILE
    uint8_t* sDS = sQD + 2 * 2 * AT_TILE;                  // buffer b at +b*2*TILE: query half h at +h*TILE, [128 keys][128 B]
    float* sLD = reinterpret_cast<float*>(sDS + 2 * 2 * AT_TILE);   // stage s: lse[128] then delta[128]
    uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(sLD) + 2 * AB3_LD_BYTES);
    uint64_t* kv_full = bars;
    uint64_t* qd_full = bars + 1;                          // [2]
    uint64_t* qd_empty = qd_full + 2;                      // [2]  (count 2: the MMA commit and the softmax warps' release of lse/delta)
    uint64_t* sdp_full = qd_empty + 2;                     // [2] per query half
    uint64_t* pds_full = sdp_full + 2;                     // [2] per query half
    uint64_t* ds_free = pds_full + 2;                      // [2] per dS^T buffer
    uint64_t* dq_full = ds_free + 2;                       // [2] per dQ accumulator
    uint64_t* dq_empty = dq_full + 2;                      // [2]
    uint64_t* fin = dq_empty + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(fin + 1);
    uint8_t* sSTG = reinterpret_cast<uint8_t*>(bars) + 256;   // BULK: staging tile of the dQ partial

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nh = blockIdx.y, n = nh / p.heads, h = nh % p.heads;
    const int k0 = blockIdx.x * AT_TK;
    const int nq = p.T / AT_TQ;
    const int row_base = n * p.T;
    const int colQ = h * AT_D, colK = p.C + h * AT_D, colV = 2 * p.C + h * AT_D;
    const int rot = (int)blockIdx.x;                       // rotated start of the query-tile walk (see attn_bwd_tc2_kernel)
    const float* lse_g = p.lse + ((long long)n * p.heads + h) * p.T;
    const float* delta_g = p.delta + ((long long)n * p.heads + h) * p.T;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmQKV);
        prefetch_tmap(&tmDO);
    }
    if (warp == 1 && lane == 0) {
        mbar_init(smem_u32(kv_full), 1);
        for (int s = 0; s < 2; ++s) {
            mbar_init(smem_u32(&qd_full[s]), 1);
            mbar_init(smem_u32(&qd_empty[s]), 1 + 8);  // tcgen05.commit (Q / dO consumed) + 8 softmax warps (lse / delta consumed)
            mbar_init(smem_u32(&sdp_full[s]), 1);
            mbar_init(smem_u32(&pds_full[s]), 8);      // one arrive per softmax warp
            mbar_init(smem_u32(&ds_free[s]), 1);
            mbar_init(smem_u32(&dq_full[s]), 1);
            mbar_init(smem_u32(&dq_empty[s]), 4);      // one arrive per drain warp
        }
        mbar_init(smem_u32(fin), 1);
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
    const uint32_t tS = tmem_base, tDP = tmem_base + 128, tDV = tmem_base + 256, tDK = tmem_base + 320,
                   tDQ = tmem_base + 384;

    if (warp == 0) {
        if (lane == 0) {
            mbar_expect_tx(smem_u32(kv_full), 2 * AT_TILE);
            tma_load_2d(smem_u32(sK), &tmQKV, smem_u32(kv_full), colK, row_base + k0);
            tma_load_2d(smem_u32(sV), &tmQKV, smem_u32(kv_full), colV, row_base + k0);
            for (int i = 0; i < nq; ++i) {
                const int stage = i & 1;
                mbar_wait(smem_u32(&qd_empty[stage]), ((i >> 1) & 1) ^ 1);
                const uint32_t fb = smem_u32(&qd_full[stage]);
                mbar_expect_tx(fb, 2 * AT_TILE + AB3_LD_BYTES);
                const int qi = (i + rot) % nq;
                tma_load_2d(smem_u32(sQD + stage * 2 * AT_TILE), &tmQKV, fb, colQ, row_base + qi * AT_TQ);
                tma_load_2d(smem_u32(sQD + stage * 2 * AT_TILE + AT_TILE), &tmDO, fb, colQ, row_base + qi * AT_TQ);
                bulk_load_1d(smem_u32(sLD + stage * 2 * AT_TQ), lse_g + qi * AT_TQ, AT_TQ * 4, fb);
                bulk_load_1d(smem_u32(sLD + stage * 2 * AT_TQ + AT_TQ), delta_g + qi * AT_TQ, AT_TQ * 4, fb);
            }
        }
    } else if (warp == 1) {
        // whole warp, warp-uniform control flow; the mma helpers elect one lane internally
        constexpr uint32_t IDESC_H = idesc_bf16_f32(128, 64, 0, 0);      // S^T, dP^T halves: A = K / V, B = 64 rows of Q / dO
        constexpr uint32_t IDESC_A = idesc_bf16_f32(128, 64, 0, 1);      // dV, dK: A in TMEM, B = dO / Q rows (MN-major)
        constexpr uint32_t IDESC_Q = idesc_bf16_f32(128, 64, 1, 1);      // dQ: A = dS^T (MN-major), B = K (MN-major)
        const uint32_t k_addr = smem_u32(sK), v_addr = smem_u32(sV);
        mbar_wait(smem_u32(kv_full), 0);
        auto issue_sdp = [&](int stage, int half) {
            const uint32_t qh = smem_u32(sQD + stage * 2 * AT_TILE) + half * (AT_TILE / 2);   // rows [64 half, +64) of Q_i
            const uint32_t doh = qh + AT_TILE;
#pragma unroll
            for (int k = 0; k < 4; ++k)
                mma_f16_ss(tS + half * 64, smem_desc_sw128(k_addr + k * 32, 16, 1024), smem_desc_sw128(qh + k * 32, 16, 1024),
                           IDESC_H, k ? 1u : 0u);
#pragma unroll
            for (int k = 0; k < 4; ++k)
                mma_f16_ss(tDP + half * 64, smem_desc_sw128(v_addr + k * 32, 16, 1024),
                           smem_desc_sw128(doh + k * 32, 16, 1024), IDESC_H, k ? 1u : 0u);
            mma_commit(smem_u32(&sdp_full[half]));
        };
        // dV += P^T_half dO_half, dK += dS^T_half Q_half: K = the half's 64 queries, 16 per step; the A operand of step k
        // (queries [16k, +16) of the half) is the 8 packed columns at [64 half + 32 (k >> 1) + 8 (k & 1)]
        auto issue_dvdk = [&](int stage, int half, bool first) {
            const uint32_t q_addr = smem_u32(sQD + stage * 2 * AT_TILE);
            const uint32_t do_addr = q_addr + AT_TILE;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint32_t col = half * 64 + (k >> 1) * 32 + (k & 1) * 8;
                mma_f16_ts(tDV, tS + col, smem_desc_sw128(do_addr + (half * 4 + k) * 2048, AT_TILE, 1024), IDESC_A,
                           (first && k == 0) ? 0u : 1u);
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint32_t col = half * 64 + (k >> 1) * 32 + (k & 1) * 8;
                mma_f16_ts(tDK, tDP + col, smem_desc_sw128(q_addr + (half * 4 + k) * 2048, AT_TILE, 1024), IDESC_A,
                           (first && k == 0) ? 0u : 1u);
            }
        };
        mbar_wait(smem_u32(&qd_full[0]), 0);
        tc_fence_after();
        issue_sdp(0, 0);
        issue_sdp(0, 1);
        for (int i = 0; i < nq; ++i) {
            const int stage = i & 1;
            const uint32_t ds_addr = smem_u32(sDS + stage * 2 * AT_TILE);
            const uint32_t tDQb = tDQ + (uint32_t)(i & 1) * 64;
            // ---- query half 0 ----
            mbar_wait(smem_u32(&pds_full[0]), i & 1);
            tc_fence_after();
            issue_dvdk(stage, 0, i == 0);
            if (i + 1 < nq) {
                mbar_wait(smem_u32(&qd_full[stage ^ 1]), ((i + 1) >> 1) & 1);
                tc_fence_after();
                issue_sdp(stage ^ 1, 0);
            }
            // ---- query half 1 ----
            mbar_wait(smem_u32(&pds_full[1]), i & 1);
            if (i >= 2) mbar_wait(smem_u32(&dq_empty[i & 1]), ((i >> 1) - 1) & 1);   // tile i-2 drained from this accumulator
            tc_fence_after();
            issue_dvdk(stage, 1, false);
            // dQ_i [128 q x 64] = dS [q x k] K [k x d]: A = dS^T buffer read MN-major (M = queries contiguous, two 64-query
            // blocks AT_TILE apart), 16 keys (= 16 rows of 128 B) per step; B = K rows (MN-major)
#pragma unroll
            for (int k = 0; k < 8; ++k)
                mma_f16_ss(tDQb, smem_desc_sw128(ds_addr + k * 2048, AT_TILE, 1024),
                           smem_desc_sw128(k_addr + k * 2048, AT_TILE, 1024), IDESC_Q, k ? 1u : 0u);
            mma_commit(smem_u32(&dq_full[i & 1]));
            mma_commit(smem_u32(&qd_empty[stage]));
            mma_commit(smem_u32(&ds_free[stage]));
            if (i + 1 < nq) issue_sdp(stage ^ 1, 1);
        }
        mma_commit(smem_u32(fin));
    } else if (warp >= 12) {
        // dQ drain warpgroup (as in attn_bwd_tc2_kernel), with the 1/sqrt(d) of dS applied here
        const int q = warp & 3;
        const int r = q * 32 + lane;
        const uint32_t lane_off = (uint32_t)(q * 32) << 16;
        for (int i = 0; i < nq; ++i) {
            const int b = i & 1;
            const int qi = (i + rot) % nq;
            mbar_wait(smem_u32(&dq_full[b]), (i >> 1) & 1);
            tc_fence_after();
            uint32_t v0[32], v1[32];
            tmem_ld32(tDQ + b * 64 + lane_off, v0);
            tmem_ld32(tDQ + b * 64 + lane_off + 32, v1);
            tc_wait_ld();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&dq_empty[b]));
            float* dst = p.dq_acc + ((long long)row_base + qi * AT_TQ + r) * p.C + h * AT_D;
            if constexpr (BULK) {
                const bool issuer = (warp == 12 && lane == 0);
                // the previous tile's bulk reduction has finished READING the staging buffer
                if (issuer) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                asm volatile("bar.sync 1, 128;" ::: "memory");
                float4* stg = reinterpret_cast<float4*>(sSTG) + r;
#pragma unroll
                for (int v = 0; v < 8; ++v)
                    stg[v * 128] = make_float4(0.125f * __uint_as_float(v0[4 * v]), 0.125f * __uint_as_float(v0[4 * v + 1]),
                                               0.125f * __uint_as_float(v0[4 * v + 2]), 0.125f * __uint_as_float(v0[4 * v + 3]));
#pragma unroll
                for (int v = 0; v < 8; ++v)
                    stg[(8 + v) * 128] = make_float4(0.125f * __uint_as_float(v1[4 * v]), 0.125f * __uint_as_float(v1[4 * v + 1]),
                                                     0.125f * __uint_as_float(v1[4 * v + 2]), 0.125f * __uint_as_float(v1[4 * v + 3]));
                fence_proxy_async();
                asm volatile("bar.sync 1, 128;" ::: "memory");
                if (issuer && !(ABL & 4)) {
                    float* tile = p.dq_acc + (((long long)n * p.heads + h) * nq + qi) * (AT_TQ * AT_D);
                    asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(tile),
                                 "r"(smem_u32(sSTG)), "r"(AT_TQ * AT_D * 4)
                                 : "memory");
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
                continue;
            }
            if (ABL & 4) {      // ablation: no reductions
                if (__uint_as_float(v0[0]) == 1.2345e-30f && __uint_as_float(v1[0]) == 1.2345e-30f) red_add_v4(dst, 0.f, 0.f, 0.f, 0.f);
                continue;
            }
#pragma unroll
            for (int e = 0; e < 32; e += 4)
                red_add_v4(dst + e, 0.125f * __uint_as_float(v0[e]), 0.125f * __uint_as_float(v0[e + 1]),
                           0.125f * __uint_as_float(v0[e + 2]), 0.125f * __uint_as_float(v0[e + 3]));
#pragma unroll
            for (int e = 0; e < 32; e += 4)
                red_add_v4(dst + 32 + e, 0.125f * __uint_as_float(v1[e]), 0.125f * __uint_as_float(v1[e + 1]),
                           0.125f * __uint_as_float(v1[e + 2]), 0.125f * __uint_as_float(v1[e + 3]));
        }
        if constexpr (BULK) {
            if (warp == 12 && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        }
    } else if (warp >= 4) {
        // eight softmax warps: warp quadrant q owns KEY rows (= TMEM lanes) [32q, 32q+32); within a 64-query half warpgroup
        // wg handles the query columns [32wg, 32wg+32)
        const int q = warp & 3;
        const int wg = (warp - 4) >> 2;
        const int r = q * 32 + lane;                       // key row of this thread
        const uint32_t lane_off = (uint32_t)(q * 32) << 16;
        const float sc = 0.125f * 1.4426950408889634f;
        auto unit = [&](int i, int half) {
            const int b = i & 1;
            const int c = half * 64 + wg * 32;             // first query column of this thread's chunk
            if (half == 0) mbar_wait(smem_u32(&qd_full[b]), (i >> 1) & 1);   // lse / delta of tile i are in shared memory
            mbar_wait(smem_u32(&sdp_full[half]), i & 1);
            tc_fence_after();
            if (ABL & 8) {     // ablation: synchronisation chain + MMAs only
                if (half == 0) mbar_wait(smem_u32(&ds_free[b]), ((i >> 1) & 1) ^ 1);
                tc_fence_before();
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive(smem_u32(&pds_full[half]));
                    if (half == 1) mbar_arrive(smem_u32(&qd_empty[b]));
                }
                return;
            }
            uint32_t sv[32], dv[32];
            tmem_ld32(tS + lane_off + c, sv);
            tmem_ld32(tDP + lane_off + c, dv);
            // the dQ MMAs of tile i-2 must have finished reading this dS^T buffer
            if (half == 0) mbar_wait(smem_u32(&ds_free[b]), ((i >> 1) & 1) ^ 1);
            const float4* l4 = reinterpret_cast<const float4*>(sLD + b * 2 * AT_TQ + c);
            const float4* d4 = reinterpret_cast<const float4*>(sLD + b * 2 * AT_TQ + AT_TQ + c);
            tc_wait_ld();
            uint32_t pk[16], dk[16];
#pragma unroll
            for (int g = 0; g < 8; ++g) {
                float4 l, d;
                if (ABL & 1) {                // ablation: no LDS of the per-column constants
                    l = make_float4(3.f, 3.f, 3.f, 3.f);
                    d = make_float4(0.1f, 0.1f, 0.1f, 0.1f);
                } else {
                    l = l4[g];                             // broadcast reads: every lane of the warp wants the same columns
                    d = d4[g];
                }
                const float lv[4] = {l.x, l.y, l.z, l.w}, dl[4] = {d.x, d.y, d.z, d.w};
                float pv[4], dsv[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int j = g * 4 + e;
                    pv[e] = ex2_approx(fmaf(__uint_as_float(sv[j]), sc, -1.4426950408889634f * lv[e]));
                    dsv[e] = pv[e] * (__uint_as_float(dv[j]) - dl[e]);       // unscaled: 1/sqrt(d) is applied to dQ and dK
                }
                const __nv_bfloat162 p0 = __floats2bfloat162_rn(pv[0], pv[1]), p1 = __floats2bfloat162_rn(pv[2], pv[3]);
                const __nv_bfloat162 s0 = __floats2bfloat162_rn(dsv[0], dsv[1]), s1 = __floats2bfloat162_rn(dsv[2], dsv[3]);
                pk[2 * g] = *reinterpret_cast<const uint32_t*>(&p0);
                pk[2 * g + 1] = *reinterpret_cast<const uint32_t*>(&p1);
                dk[2 * g] = *reinterpret_cast<const uint32_t*>(&s0);
                dk[2 * g + 1] = *reinterpret_cast<const uint32_t*>(&s1);
            }
            // P^T / dS^T as TMEM A operands, over the first 16 of the 32 columns this warp has just read
            tmem_st16(tS + lane_off + c, pk);
            tmem_st16(tDP + lane_off + c, dk);
            // dS^T row `r` (this key), queries [c, c+32) -> 64 bytes = chunks [4 wg, +4) of the half's SW128 row
            uint8_t* db = sDS + b * 2 * AT_TILE + half * AT_TILE + r * 128;
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                const int off = ((wg * 4 + g) ^ (r & 7)) << 4;
                if (ABL & 2) continue;                    // ablation: no STS of dS^T
                *reinterpret_cast<uint4*>(db + off) = make_uint4(dk[4 * g], dk[4 * g + 1], dk[4 * g + 2], dk[4 * g + 3]);
            }
            tc_wait_st();
            tc_fence_before();
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(smem_u32(&pds_full[half]));
                if (half == 1) mbar_arrive(smem_u32(&qd_empty[b]));   // lse / delta of this stage no longer needed
            }
        };
        for (int i = 0; i < nq; ++i) {
            unit(i, 0);
            unit(i, 1);
        }
        // final dV / dK: wait until every MMA of this CTA has completed
        mbar_wait(smem_u32(fin), 0);
        tc_fence_after();
        __nv_bfloat16* kp = p.dqkv + ((long long)row_base + k0 + r) * 3 * p.C + colK;
        __nv_bfloat16* vp = p.dqkv + ((long long)row_base + k0 + r) * 3 * p.C + colV;
        {
            const int c = wg * 32;
            uint32_t a[32], b[32];
            tmem_ld32(tDK + lane_off + c, a);
            tmem_ld32(tDV + lane_off + c, b);
            tc_wait_ld();
#pragma unroll
            for (int e = 0; e < 32; e += 16) {
                float ka[16], va[16];
#pragma unroll
                for (int u = 0; u < 16; ++u) {
                    ka[u] = 0.125f * __uint_as_float(a[e + u]);
                    va[u] = __uint_as_float(b[e + u]);
                }
                st16(kp + c + e, ka);
                st16(vp + c + e, va);
            }
            if (p.dbias_part) {
                // bias gradient of the qkv conv (k and v parts): column sums over this CTA's 128 keys, taken from the fp32
                // accumulators.  Warp -> lane-owned columns (reduce-scatter), the four key quadrants meet in shared memory
                // (the lse / delta stage buffers are dead by now), one plain store per column into the per-CTA slot of
                // the partial-sum workspace: global atomics onto the 2C addresses cost 1.3 ms per call at T = 4096
                // (8192 CTAs x 8 warps hitting the same 512 words)
                float* cs = sLD;                       // [0,64) dk, [64,128) dv
                const int st = threadIdx.x - 128;      // 0..255 among the softmax warps
                if (st < 128) cs[st] = 0.f;
                asm volatile("bar.sync 2, 256;" ::: "memory");
                const float tk = warp_reduce_scatter32([&](int j) { return 0.125f * __uint_as_float(a[j]); }, lane);
                const float tv = warp_reduce_scatter32([&](int j) { return __uint_as_float(b[j]); }, lane);
                atomicAdd(cs + c + lane, tk);
                atomicAdd(cs + 64 + c + lane, tv);
                asm volatile("bar.sync 2, 256;" ::: "memory");
                if (st < 128)
                    p.dbias_part[((long long)nh * gridDim.x + blockIdx.x) * 192 + 64 + st] = cs[st];
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

// dqkv[row][0:C] = bf16(dq_acc[row][:]); optionally dbias[0:C] += column sums (the q part of the qkv bias gradient).
// Thread (v, pl): channel vector v = tid % nvec, row lane pl = tid / nvec; a block walks `rows_per_block` rows.
__global__ void __launch_bounds__(256)
attn_dq_convert_kernel(const float* __restrict__ acc, __nv_bfloat16* __restrict__ dqkv, long long rows, int C,
                       int rows_per_block, float* __restrict__ dbias) {
    __shared__ float cs[2048];     // C = 64 heads <= 2048 (nvec = C / 8 <= 256 threads)
    const int nvec = C / 8, PL = 256 / nvec;
    const int v = threadIdx.x % nvec, pl = threadIdx.x / nvec;
    const long long r0 = (long long)blockIdx.x * rows_per_block;
    const long long r1 = min(rows, r0 + rows_per_block);
    for (int i = threadIdx.x; i < C; i += blockDim.x) cs[i] = 0.f;
    __syncthreads();
    if (pl < PL) {
        float s[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) s[e] = 0.f;
        for (long long row = r0 + pl; row < r1; row += PL) {
            float x[8];
            ld8(acc + row * C + v * 8, x);
            st8(dqkv + row * 3 * C + v * 8, x);
#pragma unroll
            for (int e = 0; e < 8; ++e) s[e] += x[e];
        }
        if (dbias) {
#pragma unroll
            for (int e = 0; e < 8; ++e) atomicAdd(&cs[v * 8 + e], s[e]);
        }
    }
    if (dbias) {
        __syncthreads();
        for (int i = threadIdx.x; i < C; i += blockDim.x) atomicAdd(dbias + i, cs[i]);
    }
}

// the same for the tile-major workspace of the bulk-reduction variant: acc[(sample*head)][q tile][16 vec][128 rows] x 16 B.
// One block per (q tile, sample*head): coalesced 16-byte loads (rows fastest), a shared-memory transposition, then one
// 128-byte row segment (64 bf16) per 8 threads.
__global__ void __launch_bounds__(256)
attn_dq_convert_tiles_kernel(const float* __restrict__ acc, __nv_bfloat16* __restrict__ dqkv, int T, int heads, int C,
                             float* __restrict__ dbias_part) {
    __shared__ float tile[128][65];
    const int nq = T / AT_TQ;
    const int qi = blockIdx.x, nh = blockIdx.y, n = nh / heads, h = nh % heads;
    const float4* src = reinterpret_cast<const float4*>(acc + ((long long)nh * nq + qi) * (AT_TQ * AT_D));
    for (int idx = threadIdx.x; idx < 16 * 128; idx += 256) {
        const int v = idx >> 7, r = idx & 127;
        const float4 x = src[idx];
        tile[r][4 * v] = x.x;
        tile[r][4 * v + 1] = x.y;
        tile[r][4 * v + 2] = x.z;
        tile[r][4 * v + 3] = x.w;
    }
    __syncthreads();
    const int seg = threadIdx.x & 7;              // 8 columns of the row
    float cs[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) cs[e] = 0.f;
    for (int r = threadIdx.x >> 3; r < 128; r += 32) {
        float x[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            x[e] = tile[r][seg * 8 + e];
            cs[e] += x[e];
        }
        st8(dqkv + ((long long)n * T + qi * AT_TQ + r) * 3 * C + h * AT_D + seg * 8, x);
    }
    if (dbias_part) {
        // column sums of the tile (the q part of the qkv bias gradient) into this block's slot of the partial-sum
        // workspace: lanes with equal `seg` (stride 8 in the warp) hold partial sums of the same columns
        __shared__ float part[8][64];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            cs[e] += __shfl_xor_sync(0xffffffffu, cs[e], 8);
            cs[e] += __shfl_xor_sync(0xffffffffu, cs[e], 16);
        }
        if ((threadIdx.x & 31) < 8) {
#pragma unroll
            for (int e = 0; e < 8; ++e) part[threadIdx.x >> 5][seg * 8 + e] = cs[e];
        }
        __syncthreads();
        if (threadIdx.x < 64) {
            float tot = 0.f;
#pragma unroll
            for (int w = 0; w < 8; ++w) tot += part[w][threadIdx.x];
            dbias_part[((long long)nh * nq + qi) * 192 + threadIdx.x] = tot;
        }
    }
}

// dbias[part*C + h*64 + d] += sum over samples and tiles of the per-CTA partial sums ws[(n*heads + h)][tile][part*64 + d]:
// block (h, slice) adds up every gridDim.y-th (sample, tile) slot and issues one atomic per column (dbias zeroed by the
// caller; a few dozen atomics per address instead of the thousands of the per-warp scheme)
__global__ void __launch_bounds__(192)
attn_dbias_finish_kernel(const float* __restrict__ ws, float* __restrict__ dbias, int N, int tiles, int heads, int C) {
    const int h = blockIdx.x, c = threadIdx.x;
    float tot = 0.f;
    for (int it = blockIdx.y; it < N * tiles; it += gridDim.y) {
        const int n = it / tiles, t = it % tiles;
        tot += ws[((long long)(n * heads + h) * tiles + t) * 192 + c];
    }
    atomicAdd(dbias + (c >> 6) * C + h * AT_D + (c & 63), tot);
}

bool attention_bwd_tc_applicable(int N, int T, int heads, int dtype) { return attention_tc_applicable(N, T, heads, dtype); }

int attention_bwd_tc(const void* qkv, const void* dout, const float* lse, const float* delta, void* dqkv,
                     float* dq_acc, float* dbias, bool* dbias_done, int N, int T, int heads, cudaStream_t st) {
    const int C = heads * AT_D;
    CUtensorMap tm, tmdo;
    int rc = make_mat_tmap(&tm, qkv, (long long)N * T, 3LL * C, 128);
    if (rc) return rc;
    rc = make_mat_tmap(&tmdo, dout, (long long)N * T, (long long)C, 128);
    if (rc) return rc;
    PU_CUDA(cudaMemsetAsync(dq_acc, 0, sizeof(float) * (size_t)N * T * C, st));
    AttnBwdParams p;
    p.T = T; p.heads = heads; p.C = C;
    p.lse = lse; p.delta = delta; p.dq_acc = dq_acc;
    p.dqkv = (__nv_bfloat16*)dqkv;
    dim3 grid(T / AT_TK, N * heads);
    // PU_ATTN_BWD (A/B runs; ms at T = 4096, heads = 4, batch 64).  Default 33 = attn_bwd_tc3_kernel<0, 1>: transposed scores
    // and ONE TMA bulk reduction per dQ tile, 3.45 ms.  What the ablation runs (scripts/ablate_attn.py, profiles/
    // r2_attention_ablation.md) showed: the per-lane red.global.add.v4 of the dQ partials cost 1.4 ms of the 4.61 ms of
    // variant 20 and 1.95 ms of the 5.00 ms of variant 3 -- they queue in the LSU / MIO path that the softmax warps' STS and
    // MUFU instructions share, which is why every earlier attempt (a dedicated drain warpgroup, 16 softmax warps = variant
    // 4: 5.15 ms, exponentials on the FMA pipe = 24: 4.72 ms, transposed scores alone = 3: 5.00 ms) changed nothing.
    // 20 = attn_bwd_tc2_kernel<0>: scores [q][k], P / dS through shared memory, per-lane reductions (round 1), 4.61 ms.
    // 2xx / 3xx = ablations of 20 / 3 (wrong results on purpose): bit 1 no STS (2xx) / no LDS of lse, delta (3xx), bit 4 no
    // dQ reductions, bit 8 no softmax work; 334 = 33 without issuing the bulk reductions.
    static const int variant = getenv("PU_ATTN_BWD") ? atoi(getenv("PU_ATTN_BWD")) : 33;
    // qkv bias gradient (column sums of dqkv): variant 33 collects per-CTA partial sums in the tail of the dQ workspace
    // (no atomics) and adds them up in attn_dbias_finish_kernel; the attn_bwd_tc2_kernel forms use atomics in their
    // epilogue and in the dQ convert kernel; variant 3 leaves it to the caller's separate pass
    const bool tc3 = variant == 3 || variant == 33 || variant >= 300;
    const bool part_bias = dbias != nullptr && (variant == 33 || variant == 334);
    const bool fuse_bias = dbias != nullptr && C <= 2048 && (!tc3 || part_bias);
    p.dbias = (fuse_bias && !part_bias) ? dbias : nullptr;
    p.dbias_part = part_bias ? dq_acc + (size_t)N * T * C : nullptr;
    if (p.dbias || part_bias) PU_CUDA(cudaMemsetAsync(dbias, 0, sizeof(float) * 3 * C, st));
    if (dbias_done) *dbias_done = fuse_bias;
    if (variant == 20) {
        PU_SMEM_ATTR(attn_bwd_tc2_kernel<0>, AB2_SMEM);
        attn_bwd_tc2_kernel<0><<<grid, AB2_THREADS, AB2_SMEM, st>>>(tm, tmdo, p);
    } else if (variant == 4) {
        PU_SMEM_ATTR((attn_bwd_tc2_kernel<0, 16>), AB2_SMEM);
        attn_bwd_tc2_kernel<0, 16><<<grid, 256 + 16 * 32, AB2_SMEM, st>>>(tm, tmdo, p);
    } else if (variant == 24) {
        PU_SMEM_ATTR(attn_bwd_tc2_kernel<4>, AB2_SMEM);
        attn_bwd_tc2_kernel<4><<<grid, AB2_THREADS, AB2_SMEM, st>>>(tm, tmdo, p);
    }
#define PU_ABL2(V, A)                                                             \
    else if (variant == V) {                                                      \
        PU_SMEM_ATTR((attn_bwd_tc2_kernel<0, 8, A>), AB2_SMEM);                    \
        attn_bwd_tc2_kernel<0, 8, A><<<grid, AB2_THREADS, AB2_SMEM, st>>>(tm, tmdo, p); \
    }
#define PU_ABL3(V, A)                                                             \
    else if (variant == V) {                                                      \
        PU_SMEM_ATTR(attn_bwd_tc3_kernel<A>, AB3_SMEM);                            \
        attn_bwd_tc3_kernel<A><<<grid, AB2_THREADS, AB3_SMEM, st>>>(tm, tmdo, p);   \
    }
    PU_ABL2(201, 1) PU_ABL2(204, 4) PU_ABL2(208, 8) PU_ABL2(212, 12)
    PU_ABL3(301, 1) PU_ABL3(304, 4) PU_ABL3(305, 5) PU_ABL3(308, 8) PU_ABL3(312, 12)
    else if (variant == 33 || variant == 334) {
        // dQ partials as one TMA bulk reduction per tile; tile-major workspace
        if (variant == 33) {
            PU_SMEM_ATTR((attn_bwd_tc3_kernel<0, 1>), AB3_SMEM_BULK);
            attn_bwd_tc3_kernel<0, 1><<<grid, AB2_THREADS, AB3_SMEM_BULK, st>>>(tm, tmdo, p);
        } else {
            PU_SMEM_ATTR((attn_bwd_tc3_kernel<4, 1>), AB3_SMEM_BULK);
            attn_bwd_tc3_kernel<4, 1><<<grid, AB2_THREADS, AB3_SMEM_BULK, st>>>(tm, tmdo, p);
        }
        rc = check_launch("attn_bwd_tc3_bulk");
        if (rc) return rc;
        attn_dq_convert_tiles_kernel<<<dim3(T / AT_TQ, N * heads), 256, 0, st>>>(dq_acc, (__nv_bfloat16*)dqkv, T, heads, C,
                                                                                   p.dbias_part);
        rc = check_launch("attn_dq_convert_tiles");
        if (rc || !part_bias) return rc;
        int slices = N * (T / AT_TQ) / 8;
        slices = slices < 1 ? 1 : (slices > 64 ? 64 : slices);
        attn_dbias_finish_kernel<<<dim3(heads, slices), 192, 0, st>>>(p.dbias_part, dbias, N, T / AT_TQ, heads, C);
        return check_launch("attn_dbias_finish");
    }
    else {
        PU_SMEM_ATTR(attn_bwd_tc3_kernel<0>, AB3_SMEM);
        attn_bwd_tc3_kernel<0><<<grid, AB2_THREADS, AB3_SMEM, st>>>(tm, tmdo, p);
    }
    rc = check_launch("attn_bwd_tc");
    if (rc) return rc;
    const long long rows = (long long)N * T;
    int rpb = (int)cdivll(rows, 148 * 8);
    if (rpb < 32) rpb = 32;
    attn_dq_convert_kernel<<<(unsigned)cdivll(rows, rpb), 256, 0, st>>>(dq_acc, (__nv_bfloat16*)dqkv, rows, C, rpb, p.dbias);
    return check_launch("attn_dq_convert");
}

}  // namespace pu
