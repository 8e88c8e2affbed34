// probe.cu — UMMA "lab bench": runs a host-described list of tcgen05.mma instructions on
// host-provided shared-memory images and returns the TMEM accumulator.  tests/ uses it to pin
// every descriptor convention the production kernels rely on (K-major SW128 activations,
// K-major SW64 weight slots, MN-major SW128 operands of the weight-gradient contraction)
// against a numpy matmul, independently of the big fused kernels.
#include "common.cuh"
#include "kernels.h"

namespace pob {

__global__ void __launch_bounds__(128, 1)
umma_probe_kernel(const uint8_t* __restrict__ a_img, uint32_t a_bytes,
                  const uint8_t* __restrict__ b_img, uint32_t b_bytes, uint32_t b_off,
                  const uint64_t* __restrict__ adesc, const uint64_t* __restrict__ bdesc,
                  const uint32_t* __restrict__ dcol, const uint32_t* __restrict__ accum,
                  int nops, uint32_t idesc, int out_cols, float* __restrict__ out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar_load, bar_mma;
  __shared__ uint32_t tmem_base_s;

  const uint32_t base = smem_u32(smem);
  if (warp_id() == 0) tmem_alloc(smem_u32(&tmem_base_s), 512);
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar_load), 1);
    mbar_init(smem_u32(&bar_mma), 1);
    fence_mbar_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;

  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(smem_u32(&bar_load), a_bytes + b_bytes);
    bulk_g2s(base, a_img, a_bytes, smem_u32(&bar_load));
    bulk_g2s(base + b_off, b_img, b_bytes, smem_u32(&bar_load));
    mbar_wait(smem_u32(&bar_load), 0);
    tc_fence_after();
    const uint64_t base_enc = uint64_t((base >> 4) & 0x3FFF);
    for (int i = 0; i < nops; ++i) {
      umma_f16(tmem + dcol[i], adesc[i] + base_enc, bdesc[i] + base_enc, idesc, accum[i]);
    }
    umma_commit(smem_u32(&bar_mma));
  }
  __syncwarp();
  mbar_wait(smem_u32(&bar_mma), 0);
  tc_fence_after();

  const int row = threadIdx.x;  // == TMEM lane
  const uint32_t lane_addr = tmem + (uint32_t(warp_id() * 32) << 16);
  for (int c0 = 0; c0 < out_cols; c0 += 16) {
    uint32_t v[16];
    tmem_ld16(lane_addr + c0, v);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 16; ++j)
      if (c0 + j < out_cols) out[size_t(row) * out_cols + c0 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp_id() == 0) tmem_dealloc(tmem, 512);
}

// Two-CTA variant (cta_group::2): CTA r of the pair stages a_img + r*a_bytes and b_img + r*b_bytes at the
// same shared-memory offsets; the leader issues the MMAs (M = 256 in idesc) once the peer reports its
// operands staged, commits to both CTAs, and each CTA dumps its 128 accumulator lanes: out[r*128 + lane].
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
umma_probe_pair_kernel(const uint8_t* __restrict__ a_img, uint32_t a_bytes,
                       const uint8_t* __restrict__ b_img, uint32_t b_bytes, uint32_t b_off,
                       const uint64_t* __restrict__ adesc, const uint64_t* __restrict__ bdesc,
                       const uint32_t* __restrict__ dcol, const uint32_t* __restrict__ accum,
                       int nops, uint32_t idesc, int out_cols, float* __restrict__ out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar_load, bar_peer, bar_mma;
  __shared__ uint32_t tmem_base_s;

  const uint32_t base = smem_u32(smem);
  const uint32_t rank = cluster_ctarank();
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar_load), 1);
    mbar_init(smem_u32(&bar_peer), 1);
    mbar_init(smem_u32(&bar_mma), 1);
    fence_mbar_init();
  }
  cluster_sync_all();     // barriers of both CTAs initialised before any remote arrive / multicast commit
  if (warp_id() == 0) tmem_alloc_pair(smem_u32(&tmem_base_s), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;

  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(smem_u32(&bar_load), a_bytes + b_bytes);
    bulk_g2s(base, a_img + size_t(rank) * a_bytes, a_bytes, smem_u32(&bar_load));
    bulk_g2s(base + b_off, b_img + size_t(rank) * b_bytes, b_bytes, smem_u32(&bar_load));
    mbar_wait(smem_u32(&bar_load), 0);
    if (rank != 0) {
      mbar_arrive_cluster(mapa_cluster(smem_u32(&bar_peer), 0));
    } else {
      mbar_wait_cluster(smem_u32(&bar_peer), 0);
      tc_fence_after();
      const uint64_t base_enc = uint64_t((base >> 4) & 0x3FFF);
      for (int i = 0; i < nops; ++i)
        umma_f16_pair(tmem + dcol[i], adesc[i] + base_enc, bdesc[i] + base_enc, idesc, accum[i]);
      umma_commit_pair(smem_u32(&bar_mma), 0x3);
    }
  }
  __syncwarp();
  mbar_wait(smem_u32(&bar_mma), 0);
  tc_fence_after();

  const int row = threadIdx.x;  // == TMEM lane
  const uint32_t lane_addr = tmem + (uint32_t(warp_id() * 32) << 16);
  float* o = out + size_t(rank) * 128 * out_cols;
  for (int c0 = 0; c0 < out_cols; c0 += 16) {
    uint32_t v[16];
    tmem_ld16(lane_addr + c0, v);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 16; ++j)
      if (c0 + j < out_cols) o[size_t(row) * out_cols + c0 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp_id() == 0) tmem_dealloc_pair(tmem, 512);
}

cudaError_t launch_umma_probe_pair(const void* a_img, uint32_t a_bytes, const void* b_img,
                                   uint32_t b_bytes, uint32_t b_off, const uint64_t* adesc,
                                   const uint64_t* bdesc, const uint32_t* dcol, const uint32_t* accum,
                                   int nops, uint32_t idesc, int out_cols, float* out,
                                   cudaStream_t stream) {
  const int smem = 200 * 1024;
  cudaError_t e = cudaFuncSetAttribute(umma_probe_pair_kernel,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return e;
  umma_probe_pair_kernel<<<2, 128, smem, stream>>>((const uint8_t*)a_img, a_bytes,
                                                   (const uint8_t*)b_img, b_bytes, b_off, adesc,
                                                   bdesc, dcol, accum, nops, idesc, out_cols, out);
  return cudaGetLastError();
}

cudaError_t launch_umma_probe(const void* a_img, uint32_t a_bytes, const void* b_img,
                              uint32_t b_bytes, uint32_t b_off, const uint64_t* adesc,
                              const uint64_t* bdesc, const uint32_t* dcol, const uint32_t* accum,
                              int nops, uint32_t idesc, int out_cols, float* out,
                              cudaStream_t stream) {
  const int smem = 200 * 1024;
  cudaError_t e = cudaFuncSetAttribute(umma_probe_kernel,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return e;
  umma_probe_kernel<<<1, 128, smem, stream>>>((const uint8_t*)a_img, a_bytes,
                                              (const uint8_t*)b_img, b_bytes, b_off, adesc, bdesc,
                                              dcol, accum, nops, idesc, out_cols, out);
  return cudaGetLastError();
}

}  // namespace pob
