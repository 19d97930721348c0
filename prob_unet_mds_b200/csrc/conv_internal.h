// Internal dispatch between the CUDA-core and the tcgen05 convolution kernels.
#pragma once
#include <cuda_runtime.h>

#include "../../include/probunet_b200.h"

namespace pu {
int conv_simple_launch(const PuConvArgs* a, cudaStream_t st);
int wgrad_simple_launch(const PuWgradArgs* a, cudaStream_t st);
// tcgen05 paths; *_applicable says whether shape/dtype/device fit
bool conv_tc_applicable(const PuConvArgs* a);
int conv_tc_launch(const PuConvArgs* a, cudaStream_t st);
bool wgrad_tc_applicable(const PuWgradArgs* a);
int wgrad_tc_launch(const PuWgradArgs* a, cudaStream_t st);
// (sum, sumsq) per quad of channels of an NHWC tensor, accumulated into q[N][C/4][2] (gn.cu); the CUDA-core fallback of
// the statistics that conv_tc_kernel's epilogue produces
int gn_quad_stats_launch(const void* x, int dtype, int N, int HW, int C, double* q, cudaStream_t st);
// One line on stderr per distinct shape when a bf16 layer misses the tensor-core tiles (channels not multiples of 64,
// images narrower than one tile) and runs on the CUDA-core kernel instead -- a ~100x slower path that should not be
// silent.  fp32 mode always runs there by design and is not reported.  PU_QUIET_FALLBACK=1 silences it.
void note_fallback(const char* what, int dtype, int C0, int C1, int Cout, int H, int W);
}  // namespace pu
