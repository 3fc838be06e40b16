// gemm_tc.cu — tcgen05 tensor-core dense product (placeholder until the sm_100a kernel lands;
// returns GPS_ERR_UNSUPPORTED so the dispatcher uses the exact CUDA-core kernel).
#include "gemm.cuh"

namespace gps {
int gemm_tc(const GemmParams&, cudaStream_t) { return GPS_ERR_UNSUPPORTED; }
}  // namespace gps
