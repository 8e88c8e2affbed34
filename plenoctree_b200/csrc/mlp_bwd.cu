// mlp_bwd.cu — classic (unfused) launch of the dgrad chain; the body lives in bwd_body.cuh.
#include "bwd_body.cuh"

namespace pob {

__global__ void __launch_bounds__(BWD_THREADS, 1) mlp_bwd_kernel(const __grid_constant__ BwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  bwd_body(p, smem, int(blockIdx.x), int(gridDim.x));
}

cudaError_t launch_mlp_bwd(const BwdParams& p, int num_sms, cudaStream_t stream) {
  if (p.M <= 0) return cudaSuccess;
  const long long iters = (p.M + 2 * TILE_M - 1) / (2 * TILE_M);
  const int grid = int(iters < num_sms ? iters : num_sms);
  cudaError_t e = cudaFuncSetAttribute(mlp_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)SB_TOTAL);
  if (e != cudaSuccess) return e;
  mlp_bwd_kernel<<<grid, BWD_THREADS, SB_TOTAL, stream>>>(p);
  return cudaGetLastError();
}

}  // namespace pob
