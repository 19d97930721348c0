#include "../prob_unet_mds_b200/csrc/tc_ptx.cuh"
using namespace pu::ptx;
// variant A: lane-0 branch (current); variant B: warp-uniform with elect inside the asm
__device__ __forceinline__ void mma_elect(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p, e;\n\t.reg .b32 rx;\n\t"
        "elect.sync rx|e, 0xffffffff;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void commit_elect(uint32_t bar) {
    asm volatile(
        "{\n\t.reg .pred e;\n\t.reg .b32 rx;\n\t"
        "elect.sync rx|e, 0xffffffff;\n\t"
        "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(bar) : "memory");
}
__global__ void kB(uint32_t* tslot, int nkb, uint64_t* bars) {
    extern __shared__ uint8_t smem[];
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tslot);
    int warp = threadIdx.x >> 5;
    if (warp == 1) {
        int stage = 0; uint32_t phase = 0;
        for (int kb = 0; kb < nkb; ++kb) {
            mbar_wait(smem_u32(&bars[stage]), phase);
            tc_fence_after();
            const uint32_t a_addr = smem_u32(smem + stage * 16384);
            const uint32_t b_addr = smem_u32(smem + 65536 + stage * 16384);
            const uint64_t ad0 = smem_desc_sw128(a_addr, 16, 1024), bd0 = smem_desc_sw128(b_addr, 16, 1024);
#pragma unroll
            for (int k = 0; k < 4; ++k) mma_elect(tmem_base, ad0 + 2 * k, bd0 + 2 * k, idesc_bf16_f32(128, 128, 0, 0), (kb | k) ? 1u : 0u);
            commit_elect(smem_u32(&bars[8 + stage]));
            if (++stage == 4) { stage = 0; phase ^= 1; }
        }
    }
}
