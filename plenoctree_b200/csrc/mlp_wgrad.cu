// mlp_wgrad.cu — classic (unfused) launch of the weight-gradient contraction; body in wgrad_body.cuh.
#include "wgrad_body.cuh"

namespace pob {

__global__ void __launch_bounds__(WG_THREADS, 1) mlp_wgrad_kernel(const __grid_constant__ WgradParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  wgrad_body(p, smem, int(blockIdx.x));
}

cudaError_t launch_mlp_wgrad(const WgradParams& p, int num_ctas, cudaStream_t stream) {
  cudaError_t e = cudaFuncSetAttribute(mlp_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)WG_SMEM);
  if (e != cudaSuccess) return e;
  mlp_wgrad_kernel<<<num_ctas, WG_THREADS, WG_SMEM, stream>>>(p);
  return cudaGetLastError();
}

}  // namespace pob
