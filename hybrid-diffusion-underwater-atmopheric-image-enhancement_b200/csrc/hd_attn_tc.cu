// placeholder until the tcgen05 attention / weight-gradient kernels land: report "not covered" so
// that the operator layer routes to the CUDA-core kernels.
#include "hd_common.cuh"
extern "C" int hd_attn_tc_supported(int S, int C) { return 0; }
extern "C" int hd_attn_fwd_tc(const void*, void*, float*, int, int, int, cudaStream_t) { hd_set_error("hd_attn_fwd_tc: not built"); return HD_ERR_UNSUPPORTED; }
extern "C" int hd_attn_bwd_tc(const void*, const void*, const void*, const float*, float*, void*, int, int, int, cudaStream_t) { hd_set_error("hd_attn_bwd_tc: not built"); return HD_ERR_UNSUPPORTED; }
