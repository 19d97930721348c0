// microbenchmark: tcgen05.ld throughput per SM (4 or 8 warps reading their lane quarter)
#include <cstdio>
#include <cuda_runtime.h>
#include "../prob_unet_mds_b200/csrc/tc_ptx.cuh"
using namespace pu::ptx;
__global__ void k(int iters, long long* out, float* sink, int nwarps) {
    __shared__ uint32_t slot;
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) { tmem_alloc(smem_u32(&slot), 512); tmem_relinquish(); }
    tc_fence_before(); __syncthreads(); tc_fence_after();
    uint32_t base = slot;
    float acc = 0.f;
    __syncthreads();
    long long t0 = clock64();
    if (warp < nwarps) {
        uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
        for (int i = 0; i < iters; ++i) {
            #pragma unroll
            for (int c = 0; c < 512; c += 32) {
                uint32_t v[32];
                tmem_ld32(base + lane_off + c, v);
                tc_wait_ld();
                #pragma unroll
                for (int j = 0; j < 32; ++j) acc += __uint_as_float(v[j]);
            }
        }
    }
    __syncthreads();
    long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
    if (acc == 12345.f) sink[0] = acc;
    tc_fence_before(); __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(base, 512); }
}
int main() {
    long long* d; float* s; cudaMalloc(&d, 8 * 148); cudaMalloc(&s, 4);
    for (int nw : {4, 8, 16}) {
        int iters = 200;
        k<<<148, nw * 32 < 128 ? 128 : nw * 32, 0>>>(iters, d, s, nw);
        cudaDeviceSynchronize();
        long long h; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
        double bytes = (double)iters * 512 * 128 * 4;   // per SM: all 128 lanes x 512 cols read by warps 0-3 (x2 if 8 warps)
        bytes *= nw / 4;
        printf("warps=%d cycles=%lld bytes/clk/SM=%.1f err=%s\n", nw, h, bytes / h, cudaGetErrorString(cudaGetLastError()));
    }
}
