#include "common.cuh"
namespace dlux {
size_t gemm_tc_workspace_bytes() { return 0; }
int launch_gemm_tc(const GemmParams& p, cudaStream_t st) { return DLUX_ERR_UNSUPPORTED; }
}
