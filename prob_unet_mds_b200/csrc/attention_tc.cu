// tcgen05 flash attention (placeholder until the kernel lands: reports "not applicable" so the CUDA-core
// kernel in attention_simple.cu is used).
#include "attn_internal.h"
#include "common.cuh"

namespace pu {
bool attention_tc_applicable(int, int, int, int) { return false; }
int attention_fwd_tc(const void*, void*, float*, int, int, int, cudaStream_t) {
    set_error("attention_fwd_tc: not built");
    return PU_ERR_UNSUPPORTED;
}
int attention_bwd_tc(const void*, const void*, const float*, const float*, void*, int, int, int, cudaStream_t) {
    set_error("attention_bwd_tc: not built");
    return PU_ERR_UNSUPPORTED;
}
}  // namespace pu
