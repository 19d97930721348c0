// Shared device/host helpers for the probunet_b200 kernels (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define PU_OK 0
#define PU_ERR_INVALID (-1)
#define PU_ERR_CUDA (-2)
#define PU_ERR_UNSUPPORTED (-3)

#define PU_F32 0
#define PU_BF16 1

// resample modes shared by gn_apply / gn_bwd (forward direction named)
#define PU_RS_NONE 0
#define PU_RS_UP 1    // forward: nearest x2 (out is 2H x 2W)
#define PU_RS_DOWN 2  // forward: 2x2 average (out is H/2 x W/2)

namespace pu {

void set_error(const char* fmt, ...);
int check_launch(const char* what);

#define PU_REQUIRE(cond, ...)                  \
    do {                                       \
        if (!(cond)) {                         \
            pu::set_error(__VA_ARGS__);        \
            return PU_ERR_INVALID;             \
        }                                      \
    } while (0)

#define PU_CUDA(call)                                                              \
    do {                                                                           \
        cudaError_t e__ = (call);                                                  \
        if (e__ != cudaSuccess) {                                                  \
            pu::set_error("%s failed: %s", #call, cudaGetErrorString(e__));        \
            return PU_ERR_CUDA;                                                    \
        }                                                                          \
    } while (0)

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device attribute: set it once per (call site, device), so a process
// that drives several GPUs (or several host threads) never launches a >48 KB kernel without it.
#define PU_SMEM_ATTR(kernel, bytes)                                                                              \
    do {                                                                                                         \
        static unsigned long long done__ = 0ull; /* bit d = set on device d */                                   \
        int dev__ = 0;                                                                                           \
        PU_CUDA(cudaGetDevice(&dev__));                                                                          \
        const unsigned long long bit__ = 1ull << (dev__ & 63);                                                   \
        if (!(__atomic_load_n(&done__, __ATOMIC_ACQUIRE) & bit__)) {                                             \
            PU_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes)));    \
            __atomic_fetch_or(&done__, bit__, __ATOMIC_RELEASE);                                                 \
        }                                                                                                        \
    } while (0)

__host__ __device__ inline int cdiv(int a, int b) { return (a + b - 1) / b; }
__host__ __device__ inline long long cdivll(long long a, long long b) { return (a + b - 1) / b; }

// ---- scalar load/store with conversion to float ----
__device__ __forceinline__ float ldf(const float* p) { return *p; }
__device__ __forceinline__ float ldf(const __nv_bfloat16* p) { return __bfloat162float(*p); }
__device__ __forceinline__ void stf(float* p, float v) { *p = v; }
__device__ __forceinline__ void stf(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

// ---- 8-wide vector load/store (8 channels) ----
__device__ __forceinline__ void ld8(const float* p, float (&v)[8]) {
    float4 a = *reinterpret_cast<const float4*>(p);
    float4 b = *reinterpret_cast<const float4*>(p + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void ld8(const __nv_bfloat16* p, float (&v)[8]) {
    uint4 r = *reinterpret_cast<const uint4*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float2 f = __bfloat1622float2(h[i]);
        v[2 * i] = f.x;
        v[2 * i + 1] = f.y;
    }
}
__device__ __forceinline__ void st8(float* p, const float (&v)[8]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ void st8(__nv_bfloat16* p, const float (&v)[8]) {
    uint4 r;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&r);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    *reinterpret_cast<uint4*>(p) = r;
}

// ---- 16-wide bf16 load/store with one 256-bit access (sm_100: LDG/STG.E.256): a lane touches a full 32-byte sector,
// so the row-per-lane access patterns of the tensor-core epilogues stop producing half-sector writes ----
__device__ __forceinline__ void ld16(const __nv_bfloat16* p, float (&v)[16]) {
    uint32_t r[8];
    asm volatile("ld.global.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "l"(p));
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&r[i]));
        v[2 * i] = f.x;
        v[2 * i + 1] = f.y;
    }
}
__device__ __forceinline__ void st16(__nv_bfloat16* p, const float (&v)[16]) {
    uint32_t r[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
        r[i] = *reinterpret_cast<const uint32_t*>(&h);
    }
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]),
                 "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}

// ReLU that propagates NaN like torch.relu (fmaxf would swallow it and hide invalid inputs from the validate flag)
__device__ __forceinline__ float relu_f(float v) { return v < 0.f ? 0.f : v; }
__device__ __forceinline__ float silu_f(float u) { return u / (1.0f + expf(-u)); }
// d silu(u) / du
__device__ __forceinline__ float dsilu_f(float u) {
    float s = 1.0f / (1.0f + expf(-u));
    return s * (1.0f + u * (1.0f - s));
}

// ---- packed fp32 arithmetic (Blackwell FFMA2 / FADD2 / FMUL2: two fp32 operations per issued instruction) ----
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
    float2 d;
    asm("{\n\t.reg .b64 ra, rb, rc, rd;\n\t"
        "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%6, %7};\n\t"
        "fma.rn.f32x2 rd, ra, rb, rc;\n\t"
        "mov.b64 {%0, %1}, rd;\n\t}"
        : "=f"(d.x), "=f"(d.y)
        : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
    return d;
}
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) {
    float2 d;
    asm("{\n\t.reg .b64 ra, rb, rd;\n\t"
        "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\t"
        "add.rn.f32x2 rd, ra, rb;\n\t"
        "mov.b64 {%0, %1}, rd;\n\t}"
        : "=f"(d.x), "=f"(d.y)
        : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return d;
}
__device__ __forceinline__ float2 fmul2(float2 a, float2 b) {
    float2 d;
    asm("{\n\t.reg .b64 ra, rb, rd;\n\t"
        "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\t"
        "mul.rn.f32x2 rd, ra, rb;\n\t"
        "mov.b64 {%0, %1}, rd;\n\t}"
        : "=f"(d.x), "=f"(d.y)
        : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return d;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Sums each of 32 per-lane values val(0..31) over the 32 lanes of the warp with 31 shuffles (recursive halving: at every step
// a lane keeps one half of its values and hands the other half to its partner).
// Returns the warp total of val(lane).  The values are produced on demand by `val`, so only 16 of them are live at a time.
template <typename F>
__device__ __forceinline__ float warp_reduce_scatter32(F&& val, int lane) {
    float v[16];
    {
        const bool up = (lane & 16) != 0;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const float a = val(i), b = val(i + 16);
            v[i] = (up ? b : a) + __shfl_xor_sync(0xffffffffu, up ? a : b, 16);
        }
    }
#pragma unroll
    for (int h = 8; h >= 1; h >>= 1) {
        const bool up = (lane & h) != 0;
#pragma unroll
        for (int i = 0; i < h; ++i) {
            const float send = up ? v[i] : v[i + h];
            const float keep = up ? v[i + h] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, h);
        }
    }
    return v[0];
}
// ---- Philox4x32 counter RNG (dropout masks are regenerated in backward from (seed, index)) ----
// 7 rounds: the smallest round count that passes BigCrush (Salmon et al., SC'11); the usual 10 is a safety margin that
// a dropout mask does not need, and the mask generation is ~1/4 of the instructions of the GroupNorm kernels.
constexpr int PHILOX_ROUNDS = 7;
__device__ __forceinline__ uint4 philox4x32(uint2 key, uint4 ctr) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < PHILOX_ROUNDS; ++r) {
        uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
        uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
        ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
        key.x += W0;
        key.y += W1;
    }
    return ctr;
}
// keep-mask bits for the 8 elements starting at element index `e8*8`
__device__ __forceinline__ uint32_t dropout_keep8(unsigned long long seed, unsigned long long e8, float p) {
    // one Philox call = 128 random bits = eight 16-bit uniforms (p is quantised to 1/65536)
    const uint4 a = philox4x32(make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)),
                               make_uint4((uint32_t)e8, (uint32_t)(e8 >> 32), 0x5eedu, 0u));
    const uint32_t thr = (uint32_t)(p * 65536.0f);  // drop when r < thr
    const uint32_t r[4] = {a.x, a.y, a.z, a.w};
    uint32_t m = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        m |= ((r[i] & 0xffffu) >= thr ? 1u : 0u) << (2 * i);
        m |= ((r[i] >> 16) >= thr ? 1u : 0u) << (2 * i + 1);
    }
    return m;
}

}  // namespace pu
