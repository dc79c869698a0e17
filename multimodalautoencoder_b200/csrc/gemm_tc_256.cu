#include "gemm_tc_kernel.cuh"

namespace mmae {
cudaError_t tc_launch_256(bool a_mn, bool b_mn, const TcParams& p, int grid, cudaStream_t st) {
  return tc_launch_impl<256>(a_mn, b_mn, p, grid, st);
}
}  // namespace mmae
