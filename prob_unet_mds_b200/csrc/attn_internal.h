// Internal dispatch between the CUDA-core and the tcgen05 attention kernels.
#pragma once
#include <cuda_runtime.h>

namespace pu {
int attention_fwd_simple(const void* qkv, void* out, float* lse, int N, int T, int heads, int dtype, cudaStream_t st);
int attention_delta(const void* out, const void* dout, float* delta, int N, int T, int heads, int dtype, cudaStream_t st);
int attention_bwd_simple(const void* qkv, const void* dout, const float* lse, const float* delta, void* dqkv, int N,
                         int T, int heads, int dtype, cudaStream_t st);
bool attention_tc_applicable(int N, int T, int heads, int dtype);
bool attention_bwd_tc_applicable(int N, int T, int heads, int dtype);
int attention_fwd_tc(const void* qkv, void* out, float* lse, int N, int T, int heads, cudaStream_t st);
// dbias (optional, [3C] fp32): column sums of dqkv; *dbias_done says whether the kernels produced it
int attention_bwd_tc(const void* qkv, const void* dout, const float* lse, const float* delta, void* dqkv,
                     float* dq_acc, float* dbias, bool* dbias_done, int N, int T, int heads, cudaStream_t st);
}  // namespace pu
