// overlap_probe.cu — how much of an SM's activation-store stream can run UNDER a saturated tcgen05.mma stream?
// The training forward / backward kernels must push 64 KB per 128-sample tile and layer out of the SM while the
// tensor core works; round-1/2 traces show the stores and the MMA phase adding up instead of overlapping.  This
// probe runs, on every SM at once: one warp issuing back-to-back SS-mode MMAs (M=128 N=256 K=16, or the
// cta_group::2 form), optionally a weight-slot TMA stream, and 8 warps storing at full speed by one of several
// mechanisms.  It reports cycles per MMA and store bytes per clock per SM for every combination.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I plenoctree_b200/csrc -o scripts/overlap_probe scripts/overlap_probe.cu
#include <cstdio>
#include <cstdlib>

#include "common.cuh"

using namespace pob;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

enum StoreMode { ST_NONE = 0, ST_REG = 1, ST_REG_CS = 2, ST_BULK = 3, ST_LDS_STG = 4, ST_REG_V8 = 5 };

struct Args {
  uint8_t* out;            // per-CTA destination region
  size_t out_stride;       // bytes per CTA
  size_t wrap;             // wrap the destination after this many bytes (small = L2 resident)
  const uint8_t* w;        // 1 MB weight image
  int n_mma;               // MMAs to issue (in batches of 32 with a commit + wait each)
  int store_mode;
  int with_weights;        // stream one 16 KB (8 KB in pair mode) slot per 4 MMAs
  int n_cols;              // MMA N (256 or 128)
  unsigned long long* res; // [cta][4]: mma cycles, stored bytes, store cycles
};

constexpr uint32_t P_A = 0;                 // 64 KB A tile
constexpr uint32_t P_W = 65536;             // 4 x 16 KB weight ring
constexpr uint32_t P_S = P_W + 65536;       // 64 KB staging tile for the bulk / LDS store modes
constexpr uint32_t P_TOTAL = P_S + 65536;

template <bool PAIR>
__device__ __forceinline__ void body(const Args& a, uint8_t* smem) {
  __shared__ __align__(8) uint64_t full[4], empty[4], done;
  __shared__ uint32_t tmem_base_s;
  __shared__ volatile int stop_flag;
  const uint32_t warp = warp_id(), lane = lane_id();
  const uint32_t sbase = smem_u32(smem);
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; ++i) { mbar_init(smem_u32(&full[i]), 1); mbar_init(smem_u32(&empty[i]), 1); }
    mbar_init(smem_u32(&done), 1);
    fence_mbar_init();
    stop_flag = 0;
  }
  for (uint32_t i = threadIdx.x; i < P_TOTAL / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  fence_proxy_async_smem();
  if (PAIR) cluster_sync_all();
  if (warp == 8) { if (PAIR) tmem_alloc_pair(smem_u32(&tmem_base_s), 512); else tmem_alloc(smem_u32(&tmem_base_s), 512); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const uint32_t slot_bytes = PAIR ? 8192 : 16384;

  if (warp == 8) {            // weight producer
    if (a.with_weights) {
      uint32_t slot = 0, phase = 0;
      const int nslots = a.n_mma / 4;
      for (int j = 0; j < nslots; ++j) {
        mbar_wait(smem_u32(&empty[slot]), phase ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(smem_u32(&full[slot]), slot_bytes);
          bulk_g2s(sbase + P_W + slot * 16384, a.w + size_t((j * 7 + blockIdx.x) & 63) * 16384, slot_bytes, smem_u32(&full[slot]));
        }
        __syncwarp();
        if (++slot == 4) { slot = 0; phase ^= 1; }
      }
    }
  } else if (warp == 9) {     // MMA issuer
    if (!PAIR || rank == 0) {
      const uint32_t idesc = make_idesc_f16(PAIR ? 256 : 128, a.n_cols);
      constexpr uint64_t A_HI = make_sdesc_hi(1024, LAYOUT_SW128) | (uint64_t(1) << 16);
      constexpr uint64_t W_HI = make_sdesc_hi(512, LAYOUT_SW64) | (uint64_t(1) << 16);
      uint32_t slot = 0, phase = 0, dphase = 0;
      const long long t0 = clock64();
      for (int i = 0; i < a.n_mma; i += 4) {
        if (a.with_weights) mbar_wait(smem_u32(&full[slot]), phase);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t ad = A_HI | uint64_t(((sbase + P_A) >> 4) & 0x3FFF);
          const uint64_t bd = W_HI | uint64_t(((sbase + P_W + slot * 16384) >> 4) & 0x3FFF);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint32_t d = tmem + uint32_t(k & 1) * 256u;
            if (PAIR) umma_f16_pair(d, ad + uint64_t(k & 1) * 2u + uint64_t(k >> 1) * 1024u, bd + uint64_t(k & 1) * 2u, idesc, 1u);
            else umma_f16(d, ad + uint64_t(k & 1) * 2u + uint64_t(k >> 1) * 1024u, bd + uint64_t(k & 1) * 2u, idesc, 1u);
          }
          if (a.with_weights) { if (PAIR) umma_commit_pair(smem_u32(&empty[slot]), 0x3); else umma_commit(smem_u32(&empty[slot])); }
          if ((i & 31) == 28) { if (PAIR) umma_commit_pair(smem_u32(&done), 0x3); else umma_commit(smem_u32(&done)); }
        }
        __syncwarp();
        if ((i & 31) == 28) { mbar_wait(smem_u32(&done), dphase); dphase ^= 1; }   // bounded run-ahead, like a layer
        if (++slot == 4) { slot = 0; phase ^= 1; }
      }
      const long long t1 = clock64();
      if (lane == 0) { a.res[blockIdx.x * 4 + 0] = (unsigned long long)(t1 - t0); }
    } else {
      // peer: keeps its store warps running until the leader's last batch of MMAs has completed
      uint32_t dphase = 0;
      for (int i = 0; i < a.n_mma; i += 32) { mbar_wait(smem_u32(&done), dphase); dphase ^= 1; }
    }
    stop_flag = 1;
  } else {                    // 8 store warps
    uint8_t* const base = a.out + size_t(blockIdx.x) * a.out_stride;
    const uint4 v = make_uint4(threadIdx.x, 2, 3, 4);
    size_t off = 0, total = 0;
    const long long t0 = clock64();
    if (a.store_mode != ST_NONE) {
      while (!stop_flag) {
        // one 64 KB tile per round: 256 threads x 16 x 16 B
        if (a.store_mode == ST_REG || a.store_mode == ST_REG_CS) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            uint4* p = reinterpret_cast<uint4*>(base + off + size_t(i) * 4096) + threadIdx.x;
            if (a.store_mode == ST_REG) *p = v; else __stcs(p, v);
          }
        } else if (a.store_mode == ST_REG_V8) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            uint8_t* p = base + off + size_t(i) * 8192 + threadIdx.x * 32;
            asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
          }
        } else if (a.store_mode == ST_LDS_STG) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const uint4 q = reinterpret_cast<const uint4*>(smem + P_S + i * 4096)[threadIdx.x];
            reinterpret_cast<uint4*>(base + off + size_t(i) * 4096)[threadIdx.x] = q;
          }
        } else if (a.store_mode == ST_BULK) {
          if (lane == 0) {
            bulk_s2g(base + off + warp * 8192, sbase + P_S + warp * 8192, 8192);
            bulk_commit();
            bulk_wait_read_all();
          }
          __syncwarp();
        }
        off += 65536;
        total += 65536;
        if (off >= a.wrap) off = 0;
      }
      if (a.store_mode == ST_BULK && lane == 0) bulk_wait_all();
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) { a.res[blockIdx.x * 4 + 1] = total; a.res[blockIdx.x * 4 + 2] = (unsigned long long)(t1 - t0); }
  }
  tc_fence_before();
  if (PAIR) { cluster_sync_all(); if (warp == 8) tmem_dealloc_pair(tmem, 512); }
  else { __syncthreads(); if (warp == 8) tmem_dealloc(tmem, 512); }
}

__global__ void __launch_bounds__(320, 1) probe_single(const __grid_constant__ Args a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  body<false>(a, smem);
}
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(320, 1) probe_pair(const __grid_constant__ Args a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  body<true>(a, smem);
}

int main() {
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  Args a;
  const size_t stride = size_t(64) << 20;   // 64 MB per CTA
  CK(cudaMalloc(&a.out, stride * sms));
  uint8_t* w;
  CK(cudaMalloc(&w, 1 << 20));
  CK(cudaMemset(w, 0, 1 << 20));
  a.w = w;
  a.out_stride = stride;
  CK(cudaMalloc(&a.res, sizeof(unsigned long long) * 4 * sms));
  CK(cudaFuncSetAttribute(probe_single, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)P_TOTAL));
  CK(cudaFuncSetAttribute(probe_pair, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)P_TOTAL));
  const char* names[] = {"none", "st.global from registers", "st.global.cs from registers", "bulk s2g from a staging tile",
                         "LDS + st.global", "st.global.v8 from registers"};
  unsigned long long* h = (unsigned long long*)malloc(sizeof(unsigned long long) * 4 * sms);
  for (int pair = 0; pair < 2; ++pair)
    for (int weights = 0; weights < 2; ++weights)
      for (int l2 = 0; l2 < 2; ++l2)
        for (int mode = 0; mode < 6; ++mode) {
          if (mode == 0 && l2) continue;
          if (getenv("PROBE_QUICK") && !(mode == 0 && weights == 0)) continue;
          a.n_mma = 16384;
          a.n_cols = 256;
          a.store_mode = mode;
          a.with_weights = weights;
          a.wrap = l2 ? 131072 : stride;   // L2-resident destination vs streaming to HBM
          CK(cudaMemset(a.res, 0, sizeof(unsigned long long) * 4 * sms));
          for (int rep = 0; rep < 2; ++rep) {
            if (pair) probe_pair<<<sms, 320, P_TOTAL>>>(a); else probe_single<<<sms, 320, P_TOTAL>>>(a);
            CK(cudaDeviceSynchronize());
          }
          CK(cudaMemcpy(h, a.res, sizeof(unsigned long long) * 4 * sms, cudaMemcpyDeviceToHost));
          double mc = 0, sb = 0, sc = 0;
          int nm = 0;
          for (int i = 0; i < sms; ++i) {
            if (h[i * 4]) { mc += double(h[i * 4]); ++nm; }
            sb += double(h[i * 4 + 1]);
            sc += double(h[i * 4 + 2]);
          }
          printf("{\"pair\": %d, \"weight_stream\": %d, \"dest\": \"%s\", \"store\": \"%s\", \"cycles_per_mma\": %.1f, \"store_B_per_clk_per_sm\": %.2f}\n",
                 pair, weights, l2 ? "L2" : "HBM", names[mode], mc / nm / a.n_mma, sc > 0 ? sb / sc * 1.0 : 0.0);
          fflush(stdout);
        }
  for (int pair = 0; pair < 2; ++pair) {
    a.n_mma = 16384; a.n_cols = 128; a.store_mode = 0; a.with_weights = 0; a.wrap = stride;
    CK(cudaMemset(a.res, 0, sizeof(unsigned long long) * 4 * sms));
    for (int rep = 0; rep < 2; ++rep) {
      if (pair) probe_pair<<<sms, 320, P_TOTAL>>>(a); else probe_single<<<sms, 320, P_TOTAL>>>(a);
      CK(cudaDeviceSynchronize());
    }
    CK(cudaMemcpy(h, a.res, sizeof(unsigned long long) * 4 * sms, cudaMemcpyDeviceToHost));
    double mc = 0; int nm = 0;
    for (int i = 0; i < sms; ++i) if (h[i * 4]) { mc += double(h[i * 4]); ++nm; }
    printf("{\"pair\": %d, \"mma_n\": 128, \"cycles_per_mma\": %.1f}\n", pair, mc / nm / a.n_mma);
  }
  return 0;
}
