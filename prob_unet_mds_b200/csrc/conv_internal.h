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
}  // namespace pu
