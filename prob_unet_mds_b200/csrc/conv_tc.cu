// tcgen05 / TMEM / TMA implicit-GEMM convolution for sm_100a (bf16 operands, fp32 accumulation in TMEM).
//
// Forward and data-gradient (one kernel; dgrad uses the mode-1 packed weight):
//   GEMM view  M = N*H*W output pixels (CTA tile = 8x16 pixel rectangle = 128 rows),
//              N = Cout (BN in {64,128,192,256}),  K = taps * (C0+C1), walked in blocks of 64 channels of one tap.
//   A tile: one 4-D TMA box {64 ch, 16 x, 8 y, 1 n} of the NHWC activation, shifted by the tap offset; rows that
//           fall outside the image are zero-filled by TMA, which is exactly the conv padding.  The box lands in
//           shared memory as 128 rows x 128 B with the 128-byte swizzle = the UMMA K-major SW128 canonical layout.
//   B tile: 2-D TMA box {64 k, BN rows} of the packed weight [Cout][taps*(C0+C1)].
//   Warp roles: warp0 TMA producer, warp1 MMA issuer (one elected thread), warp2 TMEM allocator,
//               warps4-7 epilogue (tcgen05.ld -> +bias (+residual) (ReLU) -> bf16 -> global NHWC).
//   Persistent CTAs (grid = min(tiles, #SM)), STAGES-deep smem ring, two TMEM accumulators so that the epilogue
//   of tile i overlaps the MMAs of tile i+1.
//
// Weight gradient:
//   GEMM view  M = Cout (128 rows), N = Cin tile (BN), K = pixels, walked in blocks of 64 pixels (4x16 rectangle).
//   Both operands are "MN-major" (the contiguous NHWC channel dim is M resp. N): dy tile and shifted-x tile are
//   loaded as [64 px][64 ch] SW128 boxes and described with MN-major UMMA descriptors.  Split-K over CTAs, partial
//   sums are added to the fp32 packed gradient with red.global.add.
//
// Replaces cuDNN behind F.conv2d / convolution_backward (networks.py:87, prob_unet.py:33).
#include <cuda.h>
#include <stdlib.h>

#include <mutex>
#include <unordered_map>

#include "../../include/probunet_b200.h"
#include "common.cuh"
#include "conv_internal.h"
#include "tc_ptx.cuh"

namespace pu {

using namespace ptx;

// ------------------------------------------------------------------------------------------------
// tensor-map encoding through the driver entry point (no link-time dependency on libcuda)
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    });
    return fn;
}

// Descriptor cache ("cached TMA descriptors keyed by (ptr, shape)").  A tensor map is a pure function of
// (base pointer, extents, box); PyTorch's caching allocator hands the same blocks back step after step, so in steady state
// every launch finds its three maps here instead of paying three cuTensorMapEncodeTiled driver calls.  The cache holds
// no reference to the memory: a stale entry for a freed-and-reallocated pointer encodes exactly the same bytes.
struct TmapKey {
    const void* ptr;
    long long a, b;     // packed extents / box
    bool operator==(const TmapKey& o) const { return ptr == o.ptr && a == o.a && b == o.b; }
};
struct TmapKeyHash {
    size_t operator()(const TmapKey& k) const {
        size_t h = std::hash<const void*>()(k.ptr);
        h ^= std::hash<long long>()(k.a) + 0x9e3779b97f4a7c15ULL + (h << 6) + (h >> 2);
        h ^= std::hash<long long>()(k.b) + 0x9e3779b97f4a7c15ULL + (h << 6) + (h >> 2);
        return h;
    }
};
static std::mutex g_tmap_mu;
static std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> g_tmaps;
static bool tmap_lookup(const TmapKey& k, CUtensorMap* m) {
    std::lock_guard<std::mutex> lock(g_tmap_mu);
    auto it = g_tmaps.find(k);
    if (it == g_tmaps.end()) return false;
    *m = it->second;
    return true;
}
static void tmap_store(const TmapKey& k, const CUtensorMap& m) {
    std::lock_guard<std::mutex> lock(g_tmap_mu);
    if (g_tmaps.size() > 16384) g_tmaps.clear();      // bounded: shapes and pointers of a training loop are few
    g_tmaps[k] = m;
}

// NHWC bf16 activation [N][H][W][C] -> 4-D map, box {64, bw, bh, 1}, 128B swizzle, zero OOB fill
int make_act_tmap(CUtensorMap* m, const void* ptr, int N, int H, int W, int C, int bw, int bh) {
    const TmapKey key{ptr, ((long long)N << 40) | ((long long)H << 20) | (long long)W,
                      ((long long)C << 24) | ((long long)bw << 12) | (long long)bh | (1LL << 62)};
    if (tmap_lookup(key, m)) return PU_OK;
    EncodeTiledFn enc = get_encode();
    if (!enc) {
        set_error("cuTensorMapEncodeTiled entry point not available");
        return PU_ERR_CUDA;
    }
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)bw, (cuuint32_t)bh, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled(activation N=%d H=%d W=%d C=%d) failed: %d", N, H, W, C, (int)r);
        return PU_ERR_CUDA;
    }
    tmap_store(key, *m);
    return PU_OK;
}

// row-major bf16 matrix [rows][cols] -> 2-D map, box {64, brows}
int make_mat_tmap(CUtensorMap* m, const void* ptr, long long rows, long long cols, int brows) {
    const TmapKey key{ptr, rows, (cols << 12) | (long long)brows};
    if (tmap_lookup(key, m)) return PU_OK;
    EncodeTiledFn enc = get_encode();
    if (!enc) {
        set_error("cuTensorMapEncodeTiled entry point not available");
        return PU_ERR_CUDA;
    }
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)brows};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled(matrix %lld x %lld) failed: %d", rows, cols, (int)r);
        return PU_ERR_CUDA;
    }
    tmap_store(key, *m);
    return PU_OK;
}

// fp32 packed weight gradient [rows = Cout][cols = taps * (C0 + C1)] -> 2-D map, box {32 floats = 128 B, 128 rows}, 128B
// swizzle: the destination of the weight-gradient kernel's cp.reduce.async.bulk.tensor epilogue
int make_dw_tmap(CUtensorMap* m, const void* ptr, long long rows, long long cols) {
    const TmapKey key{ptr, rows, (cols << 12) | 0x7f1LL};
    if (tmap_lookup(key, m)) return PU_OK;
    EncodeTiledFn enc = get_encode();
    if (!enc) {
        set_error("cuTensorMapEncodeTiled entry point not available");
        return PU_ERR_CUDA;
    }
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)cols * 4};
    cuuint32_t box[2] = {32, 128};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(ptr), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled(weight gradient %lld x %lld) failed: %d", rows, cols, (int)r);
        return PU_ERR_CUDA;
    }
    tmap_store(key, *m);
    return PU_OK;
}

static int num_sms() {
    static int n = 0;
    if (!n) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

// ------------------------------------------------------------------------------------------------
// forward / dgrad kernel
// ------------------------------------------------------------------------------------------------
struct ConvTcParams {
    int N, H, W, Cout, ksize;
    int cblk0, cblks;       // 64-channel blocks in src0 / in src0||src1
    int nkb;                // taps * cblks
    int tiles_x, tiles_y, n_tiles, total_tiles;
    int relu, bias_per_sample;
    const float* bias;
    const __nv_bfloat16* residual;
    __nv_bfloat16* out;
    double* qstats;         // [N][Cout/4][2] or nullptr
    // GroupNorm-backward epilogue (PuConvGnBwd), x0 == nullptr: off
    const __nv_bfloat16* gx0;
    const __nv_bfloat16* gx1;
    int gC0, gC1;
    const float4* gconsts;
    double* gsums;
    int gsilu;
    float gdrop;
    unsigned long long gseed;
    const uint32_t* gmask;  // stored keep bits (one word per pixel and 32 channels) or nullptr
};

// Sums each of 16 per-lane values over the 32 lanes of the warp with 16 shuffles (recursive halving: at every step a
// lane keeps one half of its values and hands the other half to its partner).  Returns the warp total of v[lane >> 1].
__device__ __forceinline__ float warp_reduce_scatter16(float (&v)[16], int lane) {
#pragma unroll
    for (int h = 8; h >= 1; h >>= 1) {
        const int bit = 2 * h;                     // 16, 8, 4, 2
        const bool up = (lane & bit) != 0;
#pragma unroll
        for (int i = 0; i < h; ++i) {
            const float send = up ? v[i] : v[i + h];
            const float keep = up ? v[i + h] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, bit);
        }
    }
    return v[0] + __shfl_xor_sync(0xffffffffu, v[0], 1);
}
__device__ __forceinline__ float sigmoid_fast(float u) {       // 0.5 tanh(u/2) + 0.5, one MUFU (as gn.cu's bf16 kernels)
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * u));
    return fmaf(t, 0.5f, 0.5f);
}

constexpr int TILE_W = 16, TILE_H = 8;      // 128 output pixels per CTA tile
constexpr int A_BYTES = 128 * 128;          // 128 rows x 64 bf16

// MT = number of 128-pixel accumulators per CTA tile.  MT = 2 (a 16x16-pixel tile, two TMA boxes sharing one weight
// tile) is used for Cout tiles of <= 128 channels: those layers are bound by the L2 -> shared-memory operand traffic
// (32 KB per 128x128x64 MMA block), and sharing B between two accumulators cuts it by a quarter per FLOP.
// HALO = true (3x3 only): the three taps of a filter COLUMN (dy = -1, 0, 1 at one dx) read windows of one
// (TH + 2) x 16-pixel activation tile that start 2048 bytes (one pixel row) apart -- multiples of the 1024-byte swizzle
// atom, so each window is an ordinary K-major SW128 operand.  One TMA box per (dx, 64-channel block) instead of three:
// the activation bytes that cross L2 -> shared memory drop from 48 to 20 KB (MT = 1) / 96 to 36 KB (MT = 2) per three
// 64-deep K blocks.  That feed (45 - 48 B/clk/SM with every SM streaming) is what bounds the layers whose
// Cout tile is <= 192 channels.  Activations and weights then live in two rings with their own barriers.
template <int BN, int MT = 1, bool GNB = false, bool HALO = false>
struct ConvTcCfg {
    static constexpr int B_BYTES = BN * 128;
    static constexpr int A_STAGE = HALO ? (TILE_H * MT + 2) * TILE_W * 128 : MT * A_BYTES;
    // GroupNorm-backward epilogue: per accumulator stage the tile's per-channel constants (float4) and sums (2 floats)
    static constexpr int GNB_BYTES = GNB ? 2 * BN * (16 + 8) : 0;
    static constexpr int MAX_STAGES = (227 * 1024 - 1280 - GNB_BYTES) / (A_STAGE + B_BYTES);
    static constexpr int STAGES = MAX_STAGES > 8 ? 8 : MAX_STAGES;      // !HALO: one ring of (A, B) stages
    static constexpr int SA = 3;                                        // HALO: activation ring ...
    static constexpr int MAX_SB = (227 * 1024 - 1280 - GNB_BYTES - SA * A_STAGE) / B_BYTES;
    static constexpr int SB = MAX_SB > 9 ? 9 : MAX_SB;                  // ... and weight ring (three tiles per A tile)
    static_assert(!HALO || SB >= 4, "the weight ring needs at least four stages");
    static constexpr int ACC1 = (BN <= 64) ? 64 : (BN <= 128) ? 128 : 256;
    static constexpr int ACC_STRIDE = MT * ACC1;          // one accumulator stage
    static constexpr int TMEM_COLS = 2 * ACC_STRIDE;
    static_assert(TMEM_COLS <= 512, "two accumulator stages must fit the 512 TMEM columns");
    static constexpr int SMEM = (HALO ? SA * A_STAGE + SB * B_BYTES : STAGES * (A_STAGE + B_BYTES)) + 1024 /*align*/ +
                                256 /*barriers*/ + GNB_BYTES;
};

// GNB = true: the epilogue is the GroupNorm-backward one (PuConvGnBwd) and nothing else (no bias / residual / ReLU)
template <int BN, int MT, bool GNB = false, bool HALO = false>
__global__ void __launch_bounds__(384, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
               const __grid_constant__ CUtensorMap tmB, const ConvTcParams p) {
    using Cfg = ConvTcCfg<BN, MT, GNB, HALO>;
    constexpr int STAGES = HALO ? Cfg::SA : Cfg::STAGES;      // stages of the activation ring (= the only ring if !HALO)
    constexpr int BSTAGES = HALO ? Cfg::SB : Cfg::STAGES;     // stages of the weight ring
    constexpr int A_STAGE = Cfg::A_STAGE;
    constexpr int TH = TILE_H * MT;              // tile height in pixels
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // 1024-aligned; an OFFSET from the
    // __shared__ symbol, so that plain C++ accesses below compile to LDS / STS instead of generic LD / ST
    uint8_t* sA = smem;
    uint8_t* sB = smem + STAGES * A_STAGE;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sB + BSTAGES * Cfg::B_BYTES);
    uint64_t* full = bars;                        // activation ring (and weights, if !HALO)
    uint64_t* empty = bars + STAGES;
    uint64_t* tfull = bars + 2 * STAGES;
    uint64_t* tempty = tfull + 2;
    uint64_t* fullB = tempty + 2;                 // HALO: weight ring
    uint64_t* emptyB = fullB + BSTAGES;
    static_assert(2 * STAGES + 4 + (HALO ? 2 * BSTAGES : 0) <= 30, "barriers must fit the 256-byte block");
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 30);
    float4* gnb_consts = reinterpret_cast<float4*>(reinterpret_cast<uint8_t*>(bars) + 256);   // [2][BN]   (GNB only)
    float* gnb_sums = reinterpret_cast<float*>(gnb_consts + 2 * BN);                          // [2][BN][2]

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmA0);
        prefetch_tmap(&tmA1);
        prefetch_tmap(&tmB);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(smem_u32(&full[s]), 1);
            mbar_init(smem_u32(&empty[s]), 1);
        }
        if constexpr (HALO) {
            for (int s = 0; s < BSTAGES; ++s) {
                mbar_init(smem_u32(&fullB[s]), 1);
                mbar_init(smem_u32(&emptyB[s]), 1);
            }
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(smem_u32(&tfull[s]), 1);
            mbar_init(smem_u32(&tempty[s]), 8);   // eight epilogue warps
        }
        fence_barrier_init();
    }
    if (warp == 2) {
        tmem_alloc(smem_u32(tmem_slot), Cfg::TMEM_COLS);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0, bstage = 0;
            uint32_t phase = 0, bphase = 0;
            for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
                const int n_tile = tile % p.n_tiles;
                int t = tile / p.n_tiles;
                const int tx = t % p.tiles_x;
                t /= p.tiles_x;
                const int ty = t % p.tiles_y;
                const int img = t / p.tiles_y;
                const int x0 = tx * TILE_W, y0 = ty * TH, n0 = n_tile * BN;
                if constexpr (HALO) {
                    // per (dx, 64-channel block): one (TH + 2) x 16-pixel activation box, then the weight tiles of its
                    // three taps (dy = -1, 0, 1)
                    for (int g = 0; g < 3 * p.cblks; ++g) {
                        const int dxi = g / p.cblks;
                        const int cb = g - dxi * p.cblks;
                        mbar_wait(smem_u32(&empty[stage]), phase ^ 1);
                        const uint32_t fa = smem_u32(&full[stage]);
                        mbar_expect_tx(fa, A_STAGE);
                        const uint32_t dstA = smem_u32(sA + stage * A_STAGE);
                        if (cb < p.cblk0)
                            tma_load_4d(dstA, &tmA0, fa, cb * 64, x0 + dxi - 1, y0 - 1, img);
                        else
                            tma_load_4d(dstA, &tmA1, fa, (cb - p.cblk0) * 64, x0 + dxi - 1, y0 - 1, img);
                        if (++stage == STAGES) {
                            stage = 0;
                            phase ^= 1;
                        }
#pragma unroll 1
                        for (int t3 = 0; t3 < 3; ++t3) {
                            const int kbw = (t3 * 3 + dxi) * p.cblks + cb;      // K block of tap (dy = t3 - 1, dx) in the weights
                            mbar_wait(smem_u32(&emptyB[bstage]), bphase ^ 1);
                            const uint32_t fbb = smem_u32(&fullB[bstage]);
                            mbar_expect_tx(fbb, Cfg::B_BYTES);
                            tma_load_2d(smem_u32(sB + bstage * Cfg::B_BYTES), &tmB, fbb, kbw * 64, n0);
                            if (++bstage == BSTAGES) {
                                bstage = 0;
                                bphase ^= 1;
                            }
                        }
                    }
                    continue;
                }
                for (int kb = 0; kb < p.nkb; ++kb) {
                    const int tap = kb / p.cblks;
                    const int cb = kb - tap * p.cblks;
                    const int dy = (p.ksize == 3) ? tap / 3 - 1 : 0;
                    const int dx = (p.ksize == 3) ? tap % 3 - 1 : 0;
                    mbar_wait(smem_u32(&empty[stage]), phase ^ 1);
                    const uint32_t fb = smem_u32(&full[stage]);
                    mbar_expect_tx(fb, A_STAGE + Cfg::B_BYTES);
#pragma unroll
                    for (int m = 0; m < MT; ++m) {
                        const uint32_t dst = smem_u32(sA + stage * A_STAGE + m * A_BYTES);
                        if (cb < p.cblk0)
                            tma_load_4d(dst, &tmA0, fb, cb * 64, x0 + dx, y0 + m * TILE_H + dy, img);
                        else
                            tma_load_4d(dst, &tmA1, fb, (cb - p.cblk0) * 64, x0 + dx, y0 + m * TILE_H + dy, img);
                    }
                    tma_load_2d(smem_u32(sB + stage * Cfg::B_BYTES), &tmB, fb, kb * 64, n0);
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == 1) {
        {   // the whole warp runs the issue loop (warp-uniform operands stay in uniform registers; a lane-0 branch made
            // ptxas emit an ELECT/R2UR waterfall per MMA); mma_f16_ss / mma_commit elect one lane internally
            constexpr uint32_t IDESC = idesc_bf16_f32(128, BN, 0, 0);
            int stage = 0, bstage = 0;
            uint32_t phase = 0, bphase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
                mbar_wait(smem_u32(&tempty[acc]), acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * Cfg::ACC_STRIDE;
                if constexpr (HALO) {
                    for (int g = 0; g < 3 * p.cblks; ++g) {
                        mbar_wait(smem_u32(&full[stage]), phase);
                        tc_fence_after();
                        const uint32_t a_addr = smem_u32(sA + stage * A_STAGE);
#pragma unroll 1
                        for (int t3 = 0; t3 < 3; ++t3) {
                            mbar_wait(smem_u32(&fullB[bstage]), bphase);
                            tc_fence_after();
                            const uint32_t b_addr = smem_u32(sB + bstage * Cfg::B_BYTES);
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                const uint64_t bd = smem_desc_sw128(b_addr + k * 32, 16, 1024);
#pragma unroll
                                for (int m = 0; m < MT; ++m) {
                                    // window of accumulator m, tap dy = t3 - 1: pixel rows [8 m + t3, +8) of the halo tile
                                    const uint64_t ad =
                                        smem_desc_sw128(a_addr + (m * TILE_H + t3) * (TILE_W * 128) + k * 32, 16, 1024);
                                    mma_f16_ss(d_tmem + m * Cfg::ACC1, ad, bd, IDESC, (g | t3 | k) ? 1u : 0u);
                                }
                            }
                            mma_commit(smem_u32(&emptyB[bstage]));
                            if (++bstage == BSTAGES) {
                                bstage = 0;
                                bphase ^= 1;
                            }
                        }
                        mma_commit(smem_u32(&empty[stage]));
                        if (++stage == STAGES) {
                            stage = 0;
                            phase ^= 1;
                        }
                    }
                    mma_commit(smem_u32(&tfull[acc]));
                    acc ^= 1;
                    if (acc == 0) acc_phase ^= 1;
                    continue;
                }
                for (int kb = 0; kb < p.nkb; ++kb) {
                    mbar_wait(smem_u32(&full[stage]), phase);
                    tc_fence_after();
                    const uint32_t a_addr = smem_u32(sA + stage * A_STAGE);
                    const uint32_t b_addr = smem_u32(sB + stage * Cfg::B_BYTES);
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint64_t bd = smem_desc_sw128(b_addr + k * 32, 16, 1024);
#pragma unroll
                        for (int m = 0; m < MT; ++m) {
                            const uint64_t ad = smem_desc_sw128(a_addr + m * A_BYTES + k * 32, 16, 1024);
                            mma_f16_ss(d_tmem + m * Cfg::ACC1, ad, bd, IDESC, (kb | k) ? 1u : 0u);
                        }
                    }
                    mma_commit(smem_u32(&empty[stage]));
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                mma_commit(smem_u32(&tfull[acc]));
                acc ^= 1;
                if (acc == 0) acc_phase ^= 1;
            }
        }
    } else if (warp >= 4) {
        // eight epilogue warps: warps 4-7 and 8-11 both cover TMEM lanes 32*(warp%4)..+31 and take alternate
        // 32-column chunks, which doubles the drain rate of the small-K (1x1) layers whose epilogue is the bottleneck
        const int q = warp & 3;
        const int ehalf = (warp - 4) >> 2;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
            const int n_tile = tile % p.n_tiles;
            int t = tile / p.n_tiles;
            const int tx = t % p.tiles_x;
            t /= p.tiles_x;
            const int ty = t % p.tiles_y;
            const int img = t / p.tiles_y;
            const int n0 = n_tile * BN;
            const int r = q * 32 + lane;
            const float* bias = p.bias ? (p.bias + (p.bias_per_sample ? (long long)img * p.Cout : 0) + n0) : nullptr;

            if constexpr (GNB) {
                // the tile's per-channel constants -> shared memory, its channel sums zeroed (overlaps the tile's MMAs)
                const int et = threadIdx.x - 128;
                for (int cc = et; cc < BN; cc += 256) {
                    gnb_consts[acc * BN + cc] = __ldg(p.gconsts + (long long)img * (p.gC0 + p.gC1) + n0 + cc);
                    gnb_sums[(acc * BN + cc) * 2] = 0.f;
                    gnb_sums[(acc * BN + cc) * 2 + 1] = 0.f;
                }
                asm volatile("bar.sync 1, 256;" ::: "memory");
            }
            mbar_wait(smem_u32(&tfull[acc]), acc_phase);
            tc_fence_after();
#pragma unroll 1
            for (int mc = ehalf; mc < MT * (BN / 32); mc += 2) {
                const int m = mc / (BN / 32);
                const int c = (mc % (BN / 32)) * 32;
                const int py = ty * TH + m * TILE_H + r / TILE_W;
                const int px = tx * TILE_W + r % TILE_W;
                const bool valid = (py < p.H) && (px < p.W);
                const long long pix = ((long long)img * p.H + py) * p.W + px;
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * Cfg::ACC_STRIDE + m * Cfg::ACC1;
                uint32_t v[32];
                tmem_ld32(taddr + c, v);
                tc_wait_ld();
                float f[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
                if constexpr (GNB) {
                    // ---- GroupNorm-backward epilogue: g = dL/dy (accumulator) -> du = dL/du, stored instead of g;
                    // per-(sample, channel) sums of du and du * xhat over the tile's pixels -> p.gsums (fp64 atomics)
                    const int Cn = p.gC0 + p.gC1;
                    const int ch0 = n0 + c;                              // first channel of this chunk in x
                    uint32_t xr[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) xr[j] = 0u;
                    uint32_t keep = 0xffffffffu;
                    if (valid) {
                        const __nv_bfloat16* xp = ch0 < p.gC0 ? p.gx0 + pix * p.gC0 + ch0 : p.gx1 + pix * p.gC1 + (ch0 - p.gC0);
                        asm volatile("ld.global.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                                     : "=r"(xr[0]), "=r"(xr[1]), "=r"(xr[2]), "=r"(xr[3]), "=r"(xr[4]), "=r"(xr[5]),
                                       "=r"(xr[6]), "=r"(xr[7])
                                     : "l"(xp));
                        asm volatile("ld.global.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                                     : "=r"(xr[8]), "=r"(xr[9]), "=r"(xr[10]), "=r"(xr[11]), "=r"(xr[12]), "=r"(xr[13]),
                                       "=r"(xr[14]), "=r"(xr[15])
                                     : "l"(xp + 16));
                        if (p.gdrop > 0.f) {
                            if (p.gmask) {
                                keep = __ldg(p.gmask + ((pix * Cn + ch0) >> 5));      // the bits gn_apply stored
                            } else {
                                keep = 0u;
                                const unsigned long long e8 = (unsigned long long)((pix * Cn + ch0) >> 3);
#pragma unroll 1
                                for (int i = 0; i < 4; ++i) keep |= dropout_keep8(p.gseed, e8 + i, p.gdrop) << (8 * i);
                            }
                        }
                    } else {
                        keep = 0u;
                    }
                    const float inv_keep = p.gdrop > 0.f ? 1.f / (1.f - p.gdrop) : 1.f;
                    const float4* kc = gnb_consts + acc * BN + c;        // (ag, bg, rstd, -mean * rstd) per channel
                    auto xval = [&](int j) {
                        const __nv_bfloat162 xx = *reinterpret_cast<const __nv_bfloat162*>(&xr[j >> 1]);
                        return (j & 1) ? __high2float(xx) : __low2float(xx);
                    };
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        float gg = ((keep >> j) & 1u) ? f[j] : 0.f;
                        if (p.gsilu) {
                            const float2 k2 = *reinterpret_cast<const float2*>(kc + j);         // (ag, bg) of channel ch0 + j
                            const float u = fmaf(xval(j), k2.x, k2.y);
                            const float s = sigmoid_fast(u);
                            gg *= (s * inv_keep) * fmaf(u, 1.f - s, 1.f);
                        } else {
                            gg *= inv_keep;
                        }
                        f[j] = gg;
                    }
                    if (valid) {
                        __nv_bfloat16* op = p.out + pix * p.Cout + n0 + c;
#pragma unroll
                        for (int j = 0; j < 32; j += 16) {
                            float ov[16];
#pragma unroll
                            for (int e = 0; e < 16; ++e) ov[e] = f[j + e];
                            st16(op + j, ov);
                        }
                    }
                    // sum over the tile's pixels (lanes) of du * xhat and of du; lane l ends up with channel ch0 + l.
                    // (Measured alternative: a [32][17] shared-memory transposition per warp instead of the shuffle
                    // butterflies -- same instruction count, 25 % slower.)
                    const float sb = warp_reduce_scatter32([&](int j) {
                        const float2 k2 = *(reinterpret_cast<const float2*>(kc + j) + 1);          // (rstd, -mean * rstd)
                        return f[j] * fmaf(xval(j), k2.x, k2.y);
                    }, lane);
                    const float sa = warp_reduce_scatter32([&](int j) { return f[j]; }, lane);
                    float* sp = gnb_sums + (acc * BN + c + lane) * 2;    // tile-level sums in shared memory (fp32) ...
                    atomicAdd(sp, sa);
                    atomicAdd(sp + 1, sb);
                } else {
                if (bias) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) f[j] += __ldg(bias + c + j);
                }
                if (valid) {
                    if (p.residual) {
                        const __nv_bfloat16* rp = p.residual + pix * p.Cout + n0 + c;
#pragma unroll
                        for (int j = 0; j < 32; j += 16) {
                            float rv[16];
                            ld16(rp + j, rv);
#pragma unroll
                            for (int e = 0; e < 16; ++e) f[j + e] += rv[e];
                        }
                    }
                    if (p.relu) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) f[j] = relu_f(f[j]);
                    }
                    __nv_bfloat16* op = p.out + pix * p.Cout + n0 + c;
#pragma unroll
                    for (int j = 0; j < 32; j += 16) {
                        float ov[16];
#pragma unroll
                        for (int e = 0; e < 16; ++e) ov[e] = f[j + e];
                        st16(op + j, ov);
                    }
                }
                if (p.qstats) {
                    // GroupNorm statistics of the consumer, from the values as stored (bf16-rounded): per lane (= pixel)
                    // the sum and sum of squares of each quad of channels, reduce-scattered over the 32 lanes (16
                    // shuffles), then one fp64 atomic per lane pair into qstats[img][quad][2]
                    float sv[16];
#pragma unroll
                    for (int qd = 0; qd < 8; ++qd) {
                        float s = 0.f, ss = 0.f;
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const float rv = valid ? __bfloat162float(__float2bfloat16_rn(f[qd * 4 + e])) : 0.f;
                            s += rv;
                            ss = fmaf(rv, rv, ss);
                        }
                        sv[2 * qd] = s;
                        sv[2 * qd + 1] = ss;
                    }
                    const float tot = warp_reduce_scatter16(sv, lane);      // total of sv[lane >> 1] over the warp
                    if ((lane & 1) == 0) {
                        const int idx = lane >> 1;                           // quad = idx >> 1, kind = idx & 1
                        atomicAdd(p.qstats + ((long long)img * (p.Cout >> 2) + ((n0 + c) >> 2)) * 2 + idx, (double)tot);
                    }
                }
                }   // !GNB
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&tempty[acc]));
            if constexpr (GNB) {
                // ... flushed once per tile and channel with fp64 atomics (the sums over the whole image cancel heavily)
                asm volatile("bar.sync 1, 256;" ::: "memory");
                const int et = threadIdx.x - 128;
                for (int cc = et; cc < BN; cc += 256) {
                    double* gp = p.gsums + ((long long)img * (p.gC0 + p.gC1) + n0 + cc) * 2;
                    atomicAdd(gp, (double)gnb_sums[(acc * BN + cc) * 2]);
                    atomicAdd(gp + 1, (double)gnb_sums[(acc * BN + cc) * 2 + 1]);
                }
            }
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
    }
}

// ------------------------------------------------------------------------------------------------
// weight-gradient kernel
// ------------------------------------------------------------------------------------------------
struct WgradTcParams {
    int N, H, W, C0, C1, Cout, ksize;
    int px_tiles_x, px_tiles_y;    // 4x16 pixel blocks per image
    long long px_blocks;           // N * px_tiles_y * px_tiles_x
    int co_tiles, ci_tiles, taps, splits;
    long long blocks_per_split;
    int total_items;
    float* dw;
};

constexpr int WG_TW = 16, WG_TH = 4;        // 64 pixels per K block
constexpr int WG_SUB = 64 * 128;            // one [64 px][64 ch] sub-tile = 8 KB

template <int BN>
struct WgradTcCfg {
    static constexpr int A_BYTES_ = 2 * WG_SUB;
    static constexpr int B_BYTES_ = (BN / 64) * WG_SUB;
    static constexpr int STAGES = (BN == 256) ? 4 : (BN == 192) ? 5 : (BN == 128) ? 6 : 8;
    static constexpr int ACC_STRIDE = (BN <= 64) ? 64 : (BN <= 128) ? 128 : 256;
    static constexpr int TMEM_COLS = 2 * ACC_STRIDE;
    static constexpr int SMEM = STAGES * (A_BYTES_ + B_BYTES_) + 1024 + 256;
};

template <int BN>
__global__ void __launch_bounds__(256, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmDY, const __grid_constant__ CUtensorMap tmX0,
                const __grid_constant__ CUtensorMap tmX1, const WgradTcParams p) {
    using Cfg = WgradTcCfg<BN>;
    constexpr int STAGES = Cfg::STAGES;
    constexpr int STAGE_BYTES = Cfg::A_BYTES_ + Cfg::B_BYTES_;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // 1024-aligned; an OFFSET from the
    // __shared__ symbol, so that plain C++ accesses below compile to LDS / STS instead of generic LD / ST
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
    uint64_t* full = bars;
    uint64_t* empty = bars + STAGES;
    uint64_t* tfull = bars + 2 * STAGES;
    uint64_t* tempty = tfull + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmDY);
        prefetch_tmap(&tmX0);
        prefetch_tmap(&tmX1);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(smem_u32(&full[s]), 1);
            mbar_init(smem_u32(&empty[s]), 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(smem_u32(&tfull[s]), 1);
            mbar_init(smem_u32(&tempty[s]), 4);
        }
        fence_barrier_init();
    }
    if (warp == 2) {
        tmem_alloc(smem_u32(tmem_slot), Cfg::TMEM_COLS);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);
    const int Ctot = p.C0 + p.C1;

    // item -> (split, tap, co_tile, ci_tile); splits outermost so that concurrently running CTAs share pixels
    auto decode = [&](int item, int& split, int& tap, int& co_t, int& ci_t) {
        ci_t = item % p.ci_tiles;
        item /= p.ci_tiles;
        co_t = item % p.co_tiles;
        item /= p.co_tiles;
        tap = item % p.taps;
        split = item / p.taps;
    };

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
                int split, tap, co_t, ci_t;
                decode(item, split, tap, co_t, ci_t);
                const int dy = (p.ksize == 3) ? tap / 3 - 1 : 0;
                const int dx = (p.ksize == 3) ? tap % 3 - 1 : 0;
                long long b0 = (long long)split * p.blocks_per_split;
                long long b1 = b0 + p.blocks_per_split;
                if (b1 > p.px_blocks) b1 = p.px_blocks;
                for (long long b = b0; b < b1; ++b) {
                    const int bx = (int)(b % p.px_tiles_x);
                    long long t = b / p.px_tiles_x;
                    const int by = (int)(t % p.px_tiles_y);
                    const int img = (int)(t / p.px_tiles_y);
                    const int x0 = bx * WG_TW, y0 = by * WG_TH;
                    mbar_wait(smem_u32(&empty[stage]), phase ^ 1);
                    const uint32_t fb = smem_u32(&full[stage]);
                    mbar_expect_tx(fb, STAGE_BYTES);
                    uint8_t* sa = smem + stage * STAGE_BYTES;
                    uint8_t* sb = sa + Cfg::A_BYTES_;
#pragma unroll
                    for (int j = 0; j < 2; ++j)
                        tma_load_4d(smem_u32(sa + j * WG_SUB), &tmDY, fb, co_t * 128 + j * 64, x0, y0, img);
#pragma unroll
                    for (int j = 0; j < BN / 64; ++j) {
                        const int c = ci_t * BN + j * 64;
                        if (c < p.C0)
                            tma_load_4d(smem_u32(sb + j * WG_SUB), &tmX0, fb, c, x0 + dx, y0 + dy, img);
                        else
                            tma_load_4d(smem_u32(sb + j * WG_SUB), &tmX1, fb, c - p.C0, x0 + dx, y0 + dy, img);
                    }
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == 1) {
        {   // the whole warp runs the issue loop (warp-uniform operands stay in uniform registers; a lane-0 branch made
            // ptxas emit an ELECT/R2UR waterfall per MMA); mma_f16_ss / mma_commit elect one lane internally
            constexpr uint32_t IDESC = idesc_bf16_f32(128, BN, 1, 1);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
                int split, tap, co_t, ci_t;
                decode(item, split, tap, co_t, ci_t);
                long long b0 = (long long)split * p.blocks_per_split;
                long long b1 = b0 + p.blocks_per_split;
                if (b1 > p.px_blocks) b1 = p.px_blocks;
                mbar_wait(smem_u32(&tempty[acc]), acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * Cfg::ACC_STRIDE;
                uint32_t first = 1;
                for (long long b = b0; b < b1; ++b) {
                    mbar_wait(smem_u32(&full[stage]), phase);
                    tc_fence_after();
                    const uint32_t a_addr = smem_u32(smem + stage * STAGE_BYTES);
                    const uint32_t b_addr = a_addr + Cfg::A_BYTES_;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        // MN-major SW128: 16 pixels (K) = 16 rows of 128 B; LBO = next 64-channel atom, SBO = 8 rows
                        const uint64_t ad = smem_desc_sw128(a_addr + k * 2048, WG_SUB, 1024);
                        const uint64_t bd = smem_desc_sw128(b_addr + k * 2048, WG_SUB, 1024);
                        mma_f16_ss(d_tmem, ad, bd, IDESC, (first && k == 0) ? 0u : 1u);
                    }
                    first = 0;
                    mma_commit(smem_u32(&empty[stage]));
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                mma_commit(smem_u32(&tfull[acc]));
                acc ^= 1;
                if (acc == 0) acc_phase ^= 1;
            }
        }
    } else if (warp >= 4) {
        const int q = warp - 4;
        int acc = 0;
        uint32_t acc_phase = 0;
        const long long row_stride = (long long)p.taps * Ctot;
        for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
            int split, tap, co_t, ci_t;
            decode(item, split, tap, co_t, ci_t);
            long long b0 = (long long)split * p.blocks_per_split;
            const bool has_work = b0 < p.px_blocks;
            const int co = co_t * 128 + q * 32 + lane;
            mbar_wait(smem_u32(&tfull[acc]), acc_phase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * Cfg::ACC_STRIDE;
#pragma unroll 1
            for (int c = 0; c < BN; c += 32) {
                uint32_t v[32];
                tmem_ld32(taddr + c, v);
                tc_wait_ld();
                if (has_work && co < p.Cout) {
                    float* dst = p.dw + co * row_stride + (long long)tap * Ctot + ci_t * BN + c;
                    // BN divides C0+C1 (host picks BN that way), so the whole 32-column chunk is in range
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + j),
                                     "f"(__uint_as_float(v[j])), "f"(__uint_as_float(v[j + 1])),
                                     "f"(__uint_as_float(v[j + 2])), "f"(__uint_as_float(v[j + 3]))
                                     : "memory");
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&tempty[acc]));
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
    }
}

// ------------------------------------------------------------------------------------------------
// weight-gradient kernel with two accumulators per CTA.  The single-accumulator kernel above is bound by the
// L2 -> shared-memory operand traffic (48 KB per 128x256x64 MMA block); sharing one operand between two
// accumulators cuts it by 25-33 %:
//   MODE 1: two Cout tiles (256 output channels) share the activation tile  (Cout % 256 == 0)
//   MODE 2: two filter taps share the dy tile                                (3x3 layers with Cout == 128)
//   MODE 3: the three taps of one filter row share the dy tile AND one activation halo tile (18 x 4 pixels instead
//           of three shifted 16 x 4 tiles): each tap's operand is a window of the halo tile, addressed by starting
//           the MN-major descriptor at pixel (row * 18 + dx + 1) -- 34 KB per 3 MMA blocks instead of 48 KB per 2.
//           Needs 3 accumulators of BN <= 128 columns.
// TMEM: accumulator a at column a * 256 (MODE 1, 2) or a * 128 (MODE 3); no double buffering: the K loops are hundreds
// of blocks long, the epilogue is a small tail.
// ------------------------------------------------------------------------------------------------
constexpr int WG_HALO_W = WG_TW + 2;
constexpr int WG_HALO = WG_HALO_W * WG_TH * 128;     // one [18 x 4 px][64 ch] halo sub-tile = 9216 B (9 swizzle atoms)

template <int BN, int MODE, bool TRED = false>
struct Wgrad2Cfg {
    static_assert(MODE != 3 || BN <= 128, "MODE 3 keeps three accumulators of BN columns in 512 TMEM columns");
    static constexpr int NACC = (MODE == 3) ? 3 : 2;
    static constexpr int ACC_STRIDE = (MODE == 3) ? 128 : 256;
    static constexpr int A_SUBS = (MODE == 1) ? 4 : 2;
    static constexpr int B_SUBS = ((MODE == 1) ? 1 : 2) * (BN / 64);
    static constexpr int B_BYTES = (MODE == 3) ? (BN / 64) * WG_HALO : B_SUBS * WG_SUB;
    static constexpr int STAGE_BYTES = A_SUBS * WG_SUB + B_BYTES;
    // TRED: the epilogue stages 32-column chunks of the accumulators in two 16 KB shared-memory tiles and hands them to
    // the TMA engine (cp.reduce.async.bulk.tensor add.f32) instead of issuing per-lane red.global.add.v4
    static constexpr int STG_BYTES = TRED ? 2 * 128 * 128 : 0;
    static constexpr int MAX_STAGES = (227 * 1024 - 1280 - STG_BYTES - (TRED ? 1024 : 0)) / STAGE_BYTES;
    static constexpr int STAGES = MAX_STAGES > 6 ? 6 : MAX_STAGES;
    static constexpr int SMEM = STAGES * STAGE_BYTES + 1024 + 256 + STG_BYTES + (TRED ? 1024 : 0);
};

template <int BN, int MODE, bool TRED = false>
__global__ void __launch_bounds__(256, 1)
wgrad_tc2_kernel(const __grid_constant__ CUtensorMap tmDY, const __grid_constant__ CUtensorMap tmX0,
                 const __grid_constant__ CUtensorMap tmX1, const __grid_constant__ CUtensorMap tmDW,
                 const WgradTcParams p) {
    using Cfg = Wgrad2Cfg<BN, MODE, TRED>;
    constexpr int STAGES = Cfg::STAGES;
    constexpr int STAGE_BYTES = Cfg::STAGE_BYTES;
    constexpr int A_BYTES2 = Cfg::A_SUBS * WG_SUB;
    constexpr int B_ONE = (BN / 64) * WG_SUB;      // one activation (B) operand
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // 1024-aligned; an OFFSET from the
    // __shared__ symbol, so that plain C++ accesses below compile to LDS / STS instead of generic LD / ST
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
    uint64_t* full = bars;
    uint64_t* empty = bars + STAGES;
    uint64_t* tfull = bars + 2 * STAGES;
    uint64_t* tempty = tfull + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 1);
    // TRED: two 1024-aligned [128 rows][128 B] staging tiles behind the barrier block
    uint8_t* stg = smem + ((STAGES * STAGE_BYTES + 256 + 1023) & ~1023);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmDY);
        prefetch_tmap(&tmX0);
        prefetch_tmap(&tmX1);
        if (TRED) prefetch_tmap(&tmDW);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(smem_u32(&full[s]), 1);
            mbar_init(smem_u32(&empty[s]), 1);
        }
        mbar_init(smem_u32(tfull), 1);
        mbar_init(smem_u32(tempty), 4);
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
    const int Ctot = p.C0 + p.C1;
    const int tap_groups = (MODE == 2) ? (p.taps + 1) / 2 : (MODE == 3) ? 3 : p.taps;
    const int co_groups = (MODE == 1) ? (p.co_tiles + 1) / 2 : p.co_tiles;

    // item -> (split, tap group, co group, ci tile)
    auto decode = [&](int item, int& split, int& tg, int& cg, int& ci_t) {
        ci_t = item % p.ci_tiles;
        item /= p.ci_tiles;
        cg = item % co_groups;
        item /= co_groups;
        tg = item % tap_groups;
        split = item / tap_groups;
    };

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
                int split, tg, cg, ci_t;
                decode(item, split, tg, cg, ci_t);
                const int tap0 = (MODE == 2) ? 2 * tg : tg;
                const int ntap = (MODE == 2 && tap0 + 1 < p.taps) ? 2 : 1;
                const int co0 = (MODE == 1) ? cg * 256 : cg * 128;
                long long b0 = (long long)split * p.blocks_per_split;
                long long b1 = b0 + p.blocks_per_split;
                if (b1 > p.px_blocks) b1 = p.px_blocks;
                for (long long b = b0; b < b1; ++b) {
                    const int bx = (int)(b % p.px_tiles_x);
                    long long t = b / p.px_tiles_x;
                    const int by = (int)(t % p.px_tiles_y);
                    const int img = (int)(t / p.px_tiles_y);
                    const int x0 = bx * WG_TW, y0 = by * WG_TH;
                    mbar_wait(smem_u32(&empty[stage]), phase ^ 1);
                    const uint32_t fb = smem_u32(&full[stage]);
                    mbar_expect_tx(fb, A_BYTES2 + ((MODE == 3) ? Cfg::B_BYTES : ntap * B_ONE));
                    uint8_t* sa = smem + stage * STAGE_BYTES;
                    uint8_t* sb = sa + A_BYTES2;
#pragma unroll
                    for (int j = 0; j < Cfg::A_SUBS; ++j)
                        tma_load_4d(smem_u32(sa + j * WG_SUB), &tmDY, fb, co0 + j * 64, x0, y0, img);
                    if (MODE == 3) {
                        // filter row tg (dy = tg - 1): one 18 x 4 halo tile per 64 channels (tmX0 / tmX1 carry that box)
#pragma unroll
                        for (int j = 0; j < BN / 64; ++j) {
                            const int c = ci_t * BN + j * 64;
                            uint8_t* dst = sb + j * WG_HALO;
                            if (c < p.C0)
                                tma_load_4d(smem_u32(dst), &tmX0, fb, c, x0 - 1, y0 + tg - 1, img);
                            else
                                tma_load_4d(smem_u32(dst), &tmX1, fb, c - p.C0, x0 - 1, y0 + tg - 1, img);
                        }
                    }
                    for (int tp = 0; MODE != 3 && tp < ntap; ++tp) {
                        const int tap = tap0 + tp;
                        const int dy = (p.ksize == 3) ? tap / 3 - 1 : 0;
                        const int dx = (p.ksize == 3) ? tap % 3 - 1 : 0;
#pragma unroll
                        for (int j = 0; j < BN / 64; ++j) {
                            const int c = ci_t * BN + j * 64;
                            uint8_t* dst = sb + tp * B_ONE + j * WG_SUB;
                            if (c < p.C0)
                                tma_load_4d(smem_u32(dst), &tmX0, fb, c, x0 + dx, y0 + dy, img);
                            else
                                tma_load_4d(smem_u32(dst), &tmX1, fb, c - p.C0, x0 + dx, y0 + dy, img);
                        }
                    }
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == 1) {
        {   // the whole warp runs the issue loop (warp-uniform operands stay in uniform registers; a lane-0 branch made
            // ptxas emit an ELECT/R2UR waterfall per MMA); mma_f16_ss / mma_commit elect one lane internally
            constexpr uint32_t IDESC = idesc_bf16_f32(128, BN, 1, 1);
            int stage = 0;
            uint32_t phase = 0;
            uint32_t it = 0;
            for (int item = blockIdx.x; item < p.total_items; item += gridDim.x, ++it) {
                int split, tg, cg, ci_t;
                decode(item, split, tg, cg, ci_t);
                const int ntap = (MODE == 2 && 2 * tg + 1 < p.taps) ? 2 : 1;
                long long b0 = (long long)split * p.blocks_per_split;
                long long b1 = b0 + p.blocks_per_split;
                if (b1 > p.px_blocks) b1 = p.px_blocks;
                mbar_wait(smem_u32(tempty), (it & 1) ^ 1);
                tc_fence_after();
                uint32_t first = 1;
                for (long long b = b0; b < b1; ++b) {
                    mbar_wait(smem_u32(&full[stage]), phase);
                    tc_fence_after();
                    const uint32_t a_addr = smem_u32(smem + stage * STAGE_BYTES);
                    const uint32_t b_addr = a_addr + A_BYTES2;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint32_t acc_flag = (first && k == 0) ? 0u : 1u;
                        if (MODE == 3) {
                            const uint64_t ad = smem_desc_sw128(a_addr + k * 2048, WG_SUB, 1024);
#pragma unroll
                            for (int t = 0; t < 3; ++t)     // tap dx = t - 1: window starts at halo pixel (k, t)
                                mma_f16_ss(tmem_base + t * Cfg::ACC_STRIDE, ad,
                                           smem_desc_sw128(b_addr + (k * WG_HALO_W + t) * 128, WG_HALO, 1024), IDESC,
                                           acc_flag);
                        } else if (MODE == 1) {
                            const uint64_t bd = smem_desc_sw128(b_addr + k * 2048, WG_SUB, 1024);
                            mma_f16_ss(tmem_base, smem_desc_sw128(a_addr + k * 2048, WG_SUB, 1024), bd, IDESC, acc_flag);
                            mma_f16_ss(tmem_base + 256, smem_desc_sw128(a_addr + 2 * WG_SUB + k * 2048, WG_SUB, 1024), bd,
                                       IDESC, acc_flag);
                        } else {
                            const uint64_t ad = smem_desc_sw128(a_addr + k * 2048, WG_SUB, 1024);
                            mma_f16_ss(tmem_base, ad, smem_desc_sw128(b_addr + k * 2048, WG_SUB, 1024), IDESC, acc_flag);
                            if (ntap == 2)
                                mma_f16_ss(tmem_base + 256, ad, smem_desc_sw128(b_addr + B_ONE + k * 2048, WG_SUB, 1024),
                                           IDESC, acc_flag);
                        }
                    }
                    first = 0;
                    mma_commit(smem_u32(&empty[stage]));
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                mma_commit(smem_u32(tfull));
            }
        }
    } else if (warp >= 4) {
        const int q = warp - 4;
        const long long row_stride = (long long)p.taps * Ctot;
        uint32_t it = 0;
        uint32_t chunk_no = 0;          // TRED: chunks staged so far by this CTA (staging tile = chunk_no & 1)
        for (int item = blockIdx.x; item < p.total_items; item += gridDim.x, ++it) {
            int split, tg, cg, ci_t;
            decode(item, split, tg, cg, ci_t);
            const int tap0 = (MODE == 2) ? 2 * tg : (MODE == 3) ? 3 * tg : tg;
            const int ntap = (MODE == 3) ? 3 : (MODE == 2 && tap0 + 1 < p.taps) ? 2 : 1;
            mbar_wait(smem_u32(tfull), it & 1);
            tc_fence_after();
            if constexpr (TRED) {
                // (MODE 3 only) every 128 x 32 fp32 chunk of the three accumulators: TMEM -> registers -> swizzled staging
                // tile -> ONE cp.reduce.async.bulk.tensor (add.f32) into the packed gradient.  The tensor memory is
                // released as soon as the last chunk has been read, the reductions drain behind the next item's MMAs.
                const bool issuer = (warp == 4 && lane == 0);
                const int row = q * 32 + lane;
                constexpr int NCH = 3 * (BN / 32);
#pragma unroll 1
                for (int ch = 0; ch < NCH; ++ch, ++chunk_no) {
                    const int half = ch / (BN / 32), c = (ch % (BN / 32)) * 32;
                    uint32_t v[32];
                    tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + half * Cfg::ACC_STRIDE + c, v);
                    // the reduction issued two chunks ago has finished READING its staging tile
                    if (issuer && chunk_no >= 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                    asm volatile("bar.sync 1, 128;" ::: "memory");
                    tc_wait_ld();
                    if (ch == NCH - 1) {
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(smem_u32(tempty));
                    }
                    uint8_t* tile = stg + (chunk_no & 1u) * (128 * 128);
#pragma unroll
                    for (int j = 0; j < 8; ++j)     // 128B swizzle: 16-byte chunk j of row r lives at chunk j ^ (r & 7)
                        *reinterpret_cast<uint4*>(tile + row * 128 + ((j ^ (row & 7)) << 4)) =
                            make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                    fence_proxy_async();
                    asm volatile("bar.sync 1, 128;" ::: "memory");
                    if (issuer) {
                        const int col = (tap0 + half) * Ctot + ci_t * BN + c;
                        asm volatile(
                            "cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%1, %2}], [%3];" ::"l"(
                                reinterpret_cast<uint64_t>(&tmDW)),
                            "r"(col), "r"(cg * 128), "r"(smem_u32(tile))
                            : "memory");
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    }
                }
                continue;
            }
#pragma unroll 1
            for (int half = 0; half < Cfg::NACC; ++half) {
                int co, tap;
                if (MODE == 1) {
                    co = cg * 256 + half * 128 + q * 32 + lane;
                    tap = tap0;
                } else {
                    if (half >= ntap) break;
                    co = cg * 128 + q * 32 + lane;
                    tap = tap0 + half;
                }
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + half * Cfg::ACC_STRIDE;
#pragma unroll 1
                for (int c = 0; c < BN; c += 32) {
                    uint32_t v[32];
                    tmem_ld32(taddr + c, v);
                    tc_wait_ld();
                    if (co < p.Cout) {
                        float* dst = p.dw + co * row_stride + (long long)tap * Ctot + ci_t * BN + c;
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + j),
                                         "f"(__uint_as_float(v[j])), "f"(__uint_as_float(v[j + 1])),
                                         "f"(__uint_as_float(v[j + 2])), "f"(__uint_as_float(v[j + 3]))
                                         : "memory");
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(tempty));
        }
        if (TRED && warp == 4 && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static bool tc_device_ok() {
    static int ok = -1;
    if (ok < 0) ok = pu_device_supports_tc();
    return ok == 1;
}

bool conv_tc_applicable(const PuConvArgs* a) {
    if (a->dtype != PU_BF16 || !tc_device_ok()) return false;
    if (a->ksize != 1 && a->ksize != 3) return false;
    if (a->C0 <= 0 || a->C0 % 64 || a->C1 % 64 || a->Cout % 64) return false;
    if (a->W < TILE_W || a->H < TILE_H) return false;   // tiny test images: CUDA-core kernel
    return true;
}

template <int BN, int MT, bool GNB = false, bool HALO = false>
static int conv_tc_launch_bn(const PuConvArgs* a, cudaStream_t st) {
    if constexpr (!GNB) {
        if (a->gn_bwd) return conv_tc_launch_bn<BN, MT, true, HALO>(a, st);
    }
    if constexpr (!HALO) {
        static const bool halo_on = !(getenv("PU_CONV_HALO") && getenv("PU_CONV_HALO")[0] == '0');   // A/B switch
        if (a->ksize == 3 && halo_on) return conv_tc_launch_bn<BN, MT, GNB, true>(a, st);
    }
    using Cfg = ConvTcCfg<BN, MT, GNB, HALO>;
    ConvTcParams p;
    p.N = a->N; p.H = a->H; p.W = a->W; p.Cout = a->Cout; p.ksize = a->ksize;
    p.cblk0 = a->C0 / 64;
    p.cblks = (a->C0 + a->C1) / 64;
    p.nkb = a->ksize * a->ksize * p.cblks;
    p.tiles_x = cdiv(a->W, TILE_W);
    p.tiles_y = cdiv(a->H, TILE_H * MT);
    p.n_tiles = a->Cout / BN;
    long long total = (long long)a->N * p.tiles_x * p.tiles_y * p.n_tiles;
    PU_REQUIRE(total < (1LL << 31), "conv_tc: too many tiles");
    p.total_tiles = (int)total;
    p.relu = (a->flags & PU_CONV_RELU) ? 1 : 0;
    p.bias_per_sample = a->bias_per_sample;
    p.bias = a->bias;
    p.residual = (const __nv_bfloat16*)a->residual;
    p.out = (__nv_bfloat16*)a->out;
    p.qstats = a->qstats;
    p.gx0 = nullptr;
    if (a->gn_bwd) {
        const PuConvGnBwd* g = a->gn_bwd;
        p.gx0 = (const __nv_bfloat16*)g->x0;
        p.gx1 = (const __nv_bfloat16*)g->x1;
        p.gC0 = g->C0;
        p.gC1 = g->C1;
        p.gconsts = (const float4*)g->consts;
        p.gsums = g->sums;
        p.gsilu = g->silu;
        p.gdrop = g->dropout_p;
        p.gseed = g->seed;
        p.gmask = (const uint32_t*)g->keep_mask;
    }

    CUtensorMap tA0, tA1, tB;
    const int box_h = HALO ? TILE_H * MT + 2 : TILE_H;          // halo: one box holds the windows of three taps (and MT tiles)
    int rc = make_act_tmap(&tA0, a->src0, a->N, a->H, a->W, a->C0, TILE_W, box_h);
    if (rc) return rc;
    if (a->C1 > 0)
        rc = make_act_tmap(&tA1, a->src1, a->N, a->H, a->W, a->C1, TILE_W, box_h);
    else
        tA1 = tA0;
    if (rc) return rc;
    rc = make_mat_tmap(&tB, a->weight, a->Cout, (long long)a->ksize * a->ksize * (a->C0 + a->C1), BN);
    if (rc) return rc;

    PU_SMEM_ATTR((conv_tc_kernel<BN, MT, GNB, HALO>), Cfg::SMEM);
    int grid = p.total_tiles < num_sms() ? p.total_tiles : num_sms();
    conv_tc_kernel<BN, MT, GNB, HALO><<<grid, 384, Cfg::SMEM, st>>>(tA0, tA1, tB, p);
    return check_launch("conv_tc");
}

int conv_tc_launch(const PuConvArgs* a, cudaStream_t st) {
    static const int mt_env = getenv("PU_CONV_MT") ? atoi(getenv("PU_CONV_MT")) : 0;   // experiments: 1 forces MT = 1
    if (a->Cout % 256 == 0) return conv_tc_launch_bn<256, 1>(a, st);
    if (a->Cout % 192 == 0) return conv_tc_launch_bn<192, 1>(a, st);
    const bool two = a->H >= 2 * TILE_H && mt_env != 1;      // 16x16-pixel tiles need at least 16 rows
    if (a->Cout % 128 == 0) return two ? conv_tc_launch_bn<128, 2>(a, st) : conv_tc_launch_bn<128, 1>(a, st);
    return two ? conv_tc_launch_bn<64, 2>(a, st) : conv_tc_launch_bn<64, 1>(a, st);
}

bool wgrad_tc_applicable(const PuWgradArgs* a) {
    if (a->dtype != PU_BF16 || !tc_device_ok()) return false;
    if (a->ksize != 1 && a->ksize != 3) return false;
    if (a->C0 <= 0 || a->C0 % 64 || a->C1 % 64 || a->Cout % 64) return false;
    if (a->W < WG_TW || a->H < WG_TH) return false;
    return true;
}

// MODE 0: single-accumulator kernel; MODE 1 / 2: wgrad_tc2_kernel (two Cout tiles / two taps per CTA)
template <int BN, int MODE>
static int wgrad_tc_launch_bn(const PuWgradArgs* a, cudaStream_t st) {
    WgradTcParams p;
    p.N = a->N; p.H = a->H; p.W = a->W; p.C0 = a->C0; p.C1 = a->C1; p.Cout = a->Cout; p.ksize = a->ksize;
    p.px_tiles_x = cdiv(a->W, WG_TW);
    p.px_tiles_y = cdiv(a->H, WG_TH);
    p.px_blocks = (long long)a->N * p.px_tiles_x * p.px_tiles_y;
    p.co_tiles = cdiv(a->Cout, 128);
    p.ci_tiles = cdiv(a->C0 + a->C1, BN);
    p.taps = a->ksize * a->ksize;
    const int co_groups = (MODE == 1) ? (p.co_tiles + 1) / 2 : p.co_tiles;
    const int tap_groups = (MODE == 2) ? (p.taps + 1) / 2 : (MODE == 3) ? 3 : p.taps;
    int base_items = co_groups * p.ci_tiles * tap_groups;
    // split-K factor: items = base_items * splits are dealt round-robin to one CTA per SM.  Cost model per item:
    // K blocks at max(MMA, operand feed) cycles each, plus the epilogue -- the fp32 reductions of the whole accumulator
    // tile, which are not overlapped (one accumulator stage) and drain at the L2 atomic rate.  More splits fill the
    // last wave better but multiply the epilogues; pick the split count with the smallest modelled time (ties: fewer).
    constexpr int NACC = (MODE == 0) ? 1 : (MODE == 3) ? 3 : 2;
    const double op_bytes = (MODE == 0)   ? (2 + BN / 64) * 8192.0
                            : (MODE == 1) ? (4 + BN / 64) * 8192.0
                            : (MODE == 2) ? (2 + 2 * (BN / 64)) * 8192.0
                                          : 16384.0 + (BN / 64) * 9216.0;
    const double t_mma = NACC * 2.0 * BN;                        // 4 x (128 x BN x 16) MMAs per accumulator and block
    const double t_feed = op_bytes / 48.0;                       // measured L2 -> shared-memory rate, B/clk/SM
    const double t_blk = t_mma > t_feed ? t_mma : t_feed;
    static const bool tred_on = !(getenv("PU_WGRAD_TRED") && getenv("PU_WGRAD_TRED")[0] == '0');   // A/B switch
    const bool tred = MODE == 3 && tred_on;
    // per-lane reductions drain at the L2 reduction rate (measured 21 B/clk/SM) with the tensor pipe idle; the TMA-reduce
    // epilogue only has to stage the accumulators in shared memory
    const double t_epi = tred ? NACC * 128.0 * BN * 4.0 / 64.0 + 2000.0 : NACC * 128.0 * BN * 4.0 / 21.0 + 2000.0;
    long long max_split = cdivll(p.px_blocks, 8);   // at least 8 K-blocks (512 pixels) per item
    if (max_split < 1) max_split = 1;
    const int sms = num_sms();
    long long hi = cdivll(8LL * sms, base_items);
    if (hi > max_split) hi = max_split;
    if (hi < 1) hi = 1;
    long long splits = 1;
    double best = 1e300;
    for (long long s = 1; s <= hi; ++s) {
        const long long items = (long long)base_items * s;
        const long long waves = cdivll(items, sms);
        const double t = (double)waves * ((double)cdivll(p.px_blocks, s) * t_blk + t_epi);
        if (t < best * (1.0 - 1e-9)) {
            best = t;
            splits = s;
        }
    }
    p.blocks_per_split = cdivll(p.px_blocks, splits);
    p.splits = (int)cdivll(p.px_blocks, p.blocks_per_split);
    p.total_items = base_items * p.splits;
    p.dw = a->dw;

    CUtensorMap tDY, tX0, tX1;
    int rc = make_act_tmap(&tDY, a->dy, a->N, a->H, a->W, a->Cout, WG_TW, WG_TH);
    if (rc) return rc;
    const int xbw = (MODE == 3) ? WG_HALO_W : WG_TW;      // MODE 3 loads 18-pixel-wide halo tiles
    rc = make_act_tmap(&tX0, a->src0, a->N, a->H, a->W, a->C0, xbw, WG_TH);
    if (rc) return rc;
    if (a->C1 > 0)
        rc = make_act_tmap(&tX1, a->src1, a->N, a->H, a->W, a->C1, xbw, WG_TH);
    else
        tX1 = tX0;
    if (rc) return rc;

    int grid = p.total_items < num_sms() ? p.total_items : num_sms();
    if constexpr (MODE == 0) {
        using Cfg = WgradTcCfg<BN>;
        PU_SMEM_ATTR(wgrad_tc_kernel<BN>, Cfg::SMEM);
        wgrad_tc_kernel<BN><<<grid, 256, Cfg::SMEM, st>>>(tDY, tX0, tX1, p);
    } else if constexpr (MODE == 3) {
        if (tred) {
            CUtensorMap tDW;
            rc = make_dw_tmap(&tDW, a->dw, a->Cout, (long long)p.taps * (a->C0 + a->C1));
            if (rc) return rc;
            using Cfg = Wgrad2Cfg<BN, 3, true>;
            PU_SMEM_ATTR((wgrad_tc2_kernel<BN, 3, true>), Cfg::SMEM);
            wgrad_tc2_kernel<BN, 3, true><<<grid, 256, Cfg::SMEM, st>>>(tDY, tX0, tX1, tDW, p);
        } else {
            using Cfg = Wgrad2Cfg<BN, 3, false>;
            PU_SMEM_ATTR((wgrad_tc2_kernel<BN, 3, false>), Cfg::SMEM);
            wgrad_tc2_kernel<BN, 3, false><<<grid, 256, Cfg::SMEM, st>>>(tDY, tX0, tX1, tX0, p);
        }
    } else {
        using Cfg = Wgrad2Cfg<BN, MODE>;
        PU_SMEM_ATTR((wgrad_tc2_kernel<BN, MODE>), Cfg::SMEM);
        wgrad_tc2_kernel<BN, MODE><<<grid, 256, Cfg::SMEM, st>>>(tDY, tX0, tX1, tX0, p);
    }
    return check_launch("wgrad_tc");
}

template <int BN>
static int wgrad_tc_pick_mode(const PuWgradArgs* a, cudaStream_t st, int mode) {
    if (mode == 1) return wgrad_tc_launch_bn<BN, 1>(a, st);
    if (mode == 2 && a->ksize == 3) return wgrad_tc_launch_bn<BN, 2>(a, st);
    if constexpr (BN <= 128) {
        if (mode == 3 && a->ksize == 3) return wgrad_tc_launch_bn<BN, 3>(a, st);
    }
    return wgrad_tc_launch_bn<BN, 0>(a, st);
}

int wgrad_tc_launch(const PuWgradArgs* a, cudaStream_t st) {
    static const int force = getenv("PU_WGRAD_MODE") ? atoi(getenv("PU_WGRAD_MODE")) : -1;   // experiments only
    const int Ctot = a->C0 + a->C1;
    // 3x3 layers take the halo kernel (MODE 3) with 128- or 64-channel activation tiles; 1x1 layers the single
    // accumulator kernel (measured: MODE 1 / 2 lose to these on every layer of the U-Net, they stay for A/B runs)
    if (a->ksize == 3 && a->W >= WG_TW && (force < 0 || force == 3)) {
        if (Ctot % 128 == 0) return wgrad_tc_pick_mode<128>(a, st, 3);
        return wgrad_tc_pick_mode<64>(a, st, 3);
    }
    int mode = 0;
    if (force >= 0) mode = force;
    if (Ctot % 256 == 0) return wgrad_tc_pick_mode<256>(a, st, mode);
    if (Ctot % 192 == 0) return wgrad_tc_pick_mode<192>(a, st, mode);
    if (Ctot % 128 == 0) return wgrad_tc_pick_mode<128>(a, st, mode);
    return wgrad_tc_pick_mode<64>(a, st, mode);
}

}  // namespace pu

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
extern "C" int pu_conv2d(const PuConvArgs* a, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    PU_REQUIRE(a && a->src0 && a->weight && a->out, "pu_conv2d: null pointer");
    PU_REQUIRE(a->N > 0 && a->H > 0 && a->W > 0 && a->C0 > 0 && a->C1 >= 0 && a->Cout > 0, "pu_conv2d: bad shape");
    PU_REQUIRE(a->ksize == 1 || a->ksize == 3, "pu_conv2d: ksize must be 1 or 3 (got %d)", a->ksize);
    PU_REQUIRE(a->dtype == PU_F32 || a->dtype == PU_BF16, "pu_conv2d: bad dtype %d", a->dtype);
    PU_REQUIRE(a->C1 == 0 || a->src1, "pu_conv2d: C1 > 0 needs src1");
    PU_REQUIRE(a->reserved == 0, "pu_conv2d: reserved must be 0");
    PU_REQUIRE(!a->qstats || a->Cout % 4 == 0, "pu_conv2d: qstats needs Cout %% 4 == 0 (got %d)", a->Cout);
    bool tc = pu::conv_tc_applicable(a) && !(a->flags & PU_CONV_FORCE_SIMPLE);
    if (a->flags & PU_CONV_FORCE_TC)
        PU_REQUIRE(tc, "pu_conv2d: PU_CONV_FORCE_TC but the tcgen05 kernel does not apply (dtype=%d C0=%d C1=%d Cout=%d)",
                   a->dtype, a->C0, a->C1, a->Cout);
    if (a->qstats) PU_CUDA(cudaMemsetAsync(a->qstats, 0, sizeof(double) * 2 * (size_t)a->N * (a->Cout / 4), st));
    if (a->gn_bwd) {
        const PuConvGnBwd* g = a->gn_bwd;
        PU_REQUIRE(tc, "pu_conv2d: the GroupNorm-backward epilogue needs the tcgen05 kernel (bf16, channels %% 64 == 0, "
                       "image >= 8x16); got dtype=%d C0=%d C1=%d Cout=%d %dx%d", a->dtype, a->C0, a->C1, a->Cout, a->H, a->W);
        PU_REQUIRE(g->x0 && g->consts && g->sums && g->C0 > 0 && g->C1 >= 0 && g->C0 + g->C1 == a->Cout &&
                   g->C0 % 32 == 0 && g->C1 % 32 == 0 && (g->C1 == 0 || g->x1),
                   "pu_conv2d: bad gn_bwd epilogue (C0=%d C1=%d Cout=%d)", g->C0, g->C1, a->Cout);
        PU_REQUIRE(!a->bias && !a->residual && !(a->flags & PU_CONV_RELU) && !a->qstats,
                   "pu_conv2d: the gn_bwd epilogue excludes bias / residual / ReLU / qstats");
        PU_REQUIRE(g->dropout_p >= 0.f && g->dropout_p < 1.f, "pu_conv2d: bad gn_bwd dropout p");
        PU_CUDA(cudaMemsetAsync(g->sums, 0, sizeof(double) * 2 * (size_t)a->N * a->Cout, st));
    }
    if (tc) return pu::conv_tc_launch(a, st);
    pu::note_fallback("pu_conv2d", a->dtype, a->C0, a->C1, a->Cout, a->H, a->W);
    int rc = pu::conv_simple_launch(a, st);
    if (rc || !a->qstats) return rc;
    // CUDA-core path (fp32 mode, images smaller than a tensor-core tile): the statistics come from a separate pass
    return pu::gn_quad_stats_launch(a->out, a->dtype, a->N, a->H * a->W, a->Cout, a->qstats, st);
}

extern "C" int pu_conv2d_wgrad(const PuWgradArgs* a, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    PU_REQUIRE(a && a->src0 && a->dy && a->dw, "pu_conv2d_wgrad: null pointer");
    PU_REQUIRE(a->N > 0 && a->H > 0 && a->W > 0 && a->C0 > 0 && a->C1 >= 0 && a->Cout > 0, "pu_conv2d_wgrad: bad shape");
    PU_REQUIRE(a->ksize == 1 || a->ksize == 3, "pu_conv2d_wgrad: ksize must be 1 or 3");
    PU_REQUIRE(a->C1 == 0 || a->src1, "pu_conv2d_wgrad: C1 > 0 needs src1");
    size_t n = (size_t)a->Cout * a->ksize * a->ksize * (a->C0 + a->C1);
    if (!a->accumulate) PU_CUDA(cudaMemsetAsync(a->dw, 0, n * sizeof(float), st));
    bool tc = pu::wgrad_tc_applicable(a) && !(a->flags & PU_CONV_FORCE_SIMPLE);
    if (a->flags & PU_CONV_FORCE_TC) PU_REQUIRE(tc, "pu_conv2d_wgrad: PU_CONV_FORCE_TC but tcgen05 kernel does not apply");
    if (tc) return pu::wgrad_tc_launch(a, st);
    pu::note_fallback("pu_conv2d_wgrad", a->dtype, a->C0, a->C1, a->Cout, a->H, a->W);
    return pu::wgrad_simple_launch(a, st);
}
