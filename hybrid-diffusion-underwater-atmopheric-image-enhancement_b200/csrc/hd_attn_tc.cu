// placeholder until the tcgen05 attention / weight-gradient kernels land: report "not covered" so
// that the operator layer routes to the CUDA-core kernels.
#include "hd_common.cuh"
extern "C" int hd_attn_tc_supported(int S, int C) { return 0; }
extern "C" int hd_attn_fwd_tc(const void*, void*, float*, int, int, int, cudaStream_t) { hd_set_error("hd_attn_fwd_tc: not built"); return HD_ERR_UNSUPPORTED; }
extern "C" int hd_attn_bwd_tc(const void*, const void*, const void*, const float*, float*, void*, int, int, int, cudaStream_t) { hd_set_error("hd_attn_bwd_tc: not built"); return HD_ERR_UNSUPPORTED; }
extern "C" int hd_wgrad_tc_supported(int, int, int, int, int, int, int, int) { return 0; }
extern "C" long long hd_wgrad_tc_workspace(int, int, int, int, int, int, int, int, int) { return 0; }
extern "C" int hd_wgrad_tc(const void*, int, const void*, int, int, const void*, int, int, float*, void*, long long, int, int, int, int, cudaStream_t) { hd_set_error("hd_wgrad_tc: not built"); return HD_ERR_UNSUPPORTED; }
