// mlp_bwdw.cu — fused backward: dgrad producers and wgrad consumers in ONE persistent launch.
//
// CTAs [0, NP) run the dgrad chain (bwd_body.cuh) and, instead of spilling every dZ_l / dO tile to an
// HBM array, copy it into a per-(producer, layer) slot pair that is rewritten every iteration; CTAs
// [NP, NP + NC) own one layer each (wgrad_body.cuh), poll the slot's `produced` counter, pull the tile
// with bulk copies, publish `consumed` as soon as it has landed in shared memory, and contract it with
// the forward-saved h_{l-1} tile.  The dZ tiles therefore live in L2 for a few microseconds only:
// 4.5 GB of HBM writes and 4.5 GB of HBM reads per training step disappear.
// All CTAs are co-resident (grid <= SM count, one CTA per SM), which is what makes the spin-waits safe.
#include "bwd_body.cuh"
#include "wgrad_body.cuh"

namespace pob {

struct BwdwParams {
  BwdParams b;
  WgradParams w;
};

__global__ void __launch_bounds__(BWD_THREADS, 1) mlp_bwdw_kernel(const __grid_constant__ BwdwParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int NP = p.b.q.NP;
  if (int(blockIdx.x) < NP) bwd_body(p.b, smem, int(blockIdx.x), NP);
  else wgrad_body(p.w, smem, int(blockIdx.x) - NP);
}

cudaError_t launch_mlp_bwdw(const BwdParams& b, const WgradParams& w, int num_consumers, cudaStream_t stream) {
  if (b.M <= 0) return cudaSuccess;
  BwdwParams p;
  p.b = b;
  p.w = w;
  const int smem = SB_TOTAL > WG_SMEM ? int(SB_TOTAL) : int(WG_SMEM);
  cudaError_t e = cudaFuncSetAttribute(mlp_bwdw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return e;
  mlp_bwdw_kernel<<<b.q.NP + num_consumers, BWD_THREADS, smem, stream>>>(p);
  return cudaGetLastError();
}

}  // namespace pob
