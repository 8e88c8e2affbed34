// l2_probe.cu — what does the SM <-> L2 fabric deliver on B200 when a persistent kernel streams L2-resident weight
// slots into shared memory (bulk TMA) while its epilogue streams activation tiles out to HBM?
// The training forward / backward kernels of round 1 all ran at 6.2-7.2 TB/s of L2-level traffic (weights in from
// L2 + activation tiles out), whichever mix: this probe measures that ceiling directly, and whether cluster
// multicast of the weight slots lowers the L2-side cost.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/l2_probe scripts/l2_probe.cu && scripts/l2_probe
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(b), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint32_t b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t b, uint32_t ph) {
  uint32_t ok;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(b), "r"(ph) : "memory");
  } while (!ok);
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_g2s_mc(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint16_t mask) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "h"(mask) : "memory");
}

constexpr int RING = 8;

// Warp 0 streams `slot_bytes` slots of an L2-resident image (wbytes, wrapped) through an 8-deep ring; the other
// 256 threads store `store_per_slot` bytes of tile data per loaded slot to a large HBM buffer (16 B per thread,
// 512 B contiguous per warp instruction, like the activation saves).  rounds = slots per CTA.
__global__ void __launch_bounds__(288, 1)
stream_kernel(const uint8_t* __restrict__ w, uint32_t wbytes, uint32_t slot_bytes, int rounds,
              uint4* __restrict__ out, size_t out_stride16, int store16_per_slot) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t full[RING];
  if (threadIdx.x == 0) {
    for (int i = 0; i < RING; ++i) mbar_init(s32(&full[i]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    if (slot_bytes == 0) return;
    if (threadIdx.x == 0) {
      const uint32_t nslots = wbytes / slot_bytes;
      uint32_t src = (blockIdx.x * 7) % nslots;
      for (int i = 0; i < rounds + RING; ++i) {
        const int s = i % RING;
        if (i >= RING) mbar_wait(s32(&full[s]), ((i / RING) - 1) & 1);
        if (i < rounds) {
          mbar_expect(s32(&full[s]), slot_bytes);
          bulk_g2s(s32(smem) + s * slot_bytes, w + size_t(src) * slot_bytes, slot_bytes, s32(&full[s]));
          if (++src == nslots) src = 0;
        }
      }
    }
  } else if (store16_per_slot > 0) {
    const int t = threadIdx.x - 32;
    uint4* p = out + size_t(blockIdx.x) * out_stride16;
    const uint4 v = make_uint4(t, blockIdx.x, 3, 4);
    const size_t total = size_t(rounds) * store16_per_slot;
    for (size_t i = t; i < total; i += 256) p[i] = v;
  }
}

// Cluster multicast: every CTA of a cluster of CS receives every slot; CTA r issues slots r, r+CS, ... of each
// 8-slot round with the full cluster mask.  One cluster barrier per round protects the ring.
template <int CS>
__global__ void __launch_bounds__(32, 1)
mc_kernel(const uint8_t* __restrict__ w, uint32_t wbytes, uint32_t slot_bytes, int rounds8, int use_mc) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t full[RING];
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  if (threadIdx.x == 0) {
    for (int i = 0; i < RING; ++i) mbar_init(s32(&full[i]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  const uint32_t nslots = wbytes / slot_bytes;
  uint32_t src = ((blockIdx.x / CS) * 8) % nslots;
  for (int r = 0; r < rounds8; ++r) {
    if (threadIdx.x == 0) {
      for (int s = 0; s < RING; ++s) mbar_expect(s32(&full[s]), slot_bytes);
    }
    // every CTA has armed its barriers before any multicast copy may complete on them
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    if (threadIdx.x == 0) {
      for (int s = 0; s < RING; ++s) {
        const uint32_t sl = (src + s) % nslots;
        if (use_mc) {
          if (s % CS == (int)rank)
            bulk_g2s_mc(s32(smem) + s * slot_bytes, w + size_t(sl) * slot_bytes, slot_bytes, s32(&full[s]), uint16_t((1u << CS) - 1));
        } else {
          bulk_g2s(s32(smem) + s * slot_bytes, w + size_t(sl) * slot_bytes, slot_bytes, s32(&full[s]));
        }
      }
      for (int s = 0; s < RING; ++s) mbar_wait(s32(&full[s]), r & 1);
    }
    src = (src + RING) % nslots;
    __syncwarp();
  }
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

template <int CS>
int run_mc(const uint8_t* w, uint32_t wbytes, uint32_t slot, int sms, int use_mc) {
  const int rounds8 = 2000;
  const int grid = (sms / CS) * CS;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(32);
  cfg.dynamicSmemBytes = RING * slot;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CS;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  CK(cudaFuncSetAttribute(mc_kernel<CS>, cudaFuncAttributeMaxDynamicSharedMemorySize, RING * slot));
  if (CS > 8) CK(cudaFuncSetAttribute(mc_kernel<CS>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    CK(cudaEventRecord(e0));
    CK(cudaLaunchKernelEx(&cfg, mc_kernel<CS>, w, wbytes, slot, rounds8, use_mc));
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  const double delivered = double(grid) * rounds8 * RING * slot;
  printf("{\"probe\": \"cluster_slots\", \"cluster\": %d, \"multicast\": %d, \"slot_bytes\": %u, \"ctas\": %d, \"ms\": %.4f, \"delivered_TBps\": %.3f}\n",
         CS, use_mc, slot, grid, best, delivered / best * 1e-9);
  return 0;
}

int main() {
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  int clk = 0;
  cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf("{\"device\": \"%s\", \"sms\": %d, \"l2_bytes\": %d, \"clock_khz\": %d}\n", prop.name, sms, prop.l2CacheSize, clk);
  const uint32_t wbytes = 1 << 20;   // one MLP's weight image: L2 resident
  uint8_t* w;
  CK(cudaMalloc(&w, wbytes));
  CK(cudaMemset(w, 1, wbytes));
  const size_t out_bytes = size_t(12) << 30;
  uint4* out;
  CK(cudaMalloc(&out, out_bytes));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  CK(cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, RING * 16384));

  // (slot_bytes, store bytes per slot): reads only, stores only, and the training mixes
  const int cfgs[][2] = {{16384, 0}, {8192, 0}, {4096, 0}, {0, 16384}, {16384, 16384}, {16384, 8192}, {8192, 16384}, {16384, 32768}};
  for (auto& c : cfgs) {
    const uint32_t slot = c[0];
    const int st16 = c[1] / 16;
    int rounds = 20000;
    if (st16 > 0) {
      const size_t cap = out_bytes / sms / (size_t(st16) * 16);
      if (size_t(rounds) > cap) rounds = int(cap);
    }
    const size_t per_cta16 = size_t(rounds) * (st16 > 0 ? st16 : 1);
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
      CK(cudaEventRecord(e0));
      stream_kernel<<<sms, 288, RING * 16384>>>(w, wbytes, slot, rounds, out, per_cta16, st16);
      CK(cudaEventRecord(e1));
      CK(cudaEventSynchronize(e1));
      CK(cudaGetLastError());
      float ms;
      CK(cudaEventElapsedTime(&ms, e0, e1));
      if (ms < best) best = ms;
    }
    const double rd = double(sms) * rounds * slot, wr = double(sms) * rounds * c[1];
    printf("{\"probe\": \"stream\", \"slot_bytes\": %u, \"store_bytes_per_slot\": %d, \"ms\": %.4f, \"l2_read_TBps\": %.3f, \"hbm_write_TBps\": %.3f, \"sum_TBps\": %.3f}\n",
           slot, c[1], best, rd / best * 1e-9, wr / best * 1e-9, (rd + wr) / best * 1e-9);
  }
  for (int mc = 0; mc < 2; ++mc) {
    if (run_mc<2>(w, wbytes, 16384, sms, mc)) return 1;
    if (run_mc<4>(w, wbytes, 16384, sms, mc)) return 1;
    if (run_mc<8>(w, wbytes, 16384, sms, mc)) return 1;
  }
  if (run_mc<16>(w, wbytes, 8192, sms, 0)) printf("{\"probe\": \"cluster16 unicast failed\"}\n");
  if (run_mc<16>(w, wbytes, 8192, sms, 1)) printf("{\"probe\": \"cluster16 multicast failed\"}\n");
  return 0;
}
