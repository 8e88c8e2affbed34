// write_bw_probe.cu — which store flavour reaches the highest write-only HBM bandwidth on B200?
// (the training forward / backward kernels are bound by their activation stores; torch.fill_ measures 3.9 TB/s)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o write_bw_probe scripts/write_bw_probe.cu && ./write_bw_probe
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

template <int MODE>
__global__ void store_kernel(uint4* __restrict__ dst, size_t n16) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const uint4 v = make_uint4(threadIdx.x, blockIdx.x, 3, 4);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) {
    if (MODE == 0) dst[i] = v;
    else if (MODE == 1) __stcs(dst + i, v);
    else if (MODE == 2) __stwt(dst + i, v);
    else if (MODE == 3) __stcg(dst + i, v);
  }
}

// 32-byte stores (sm_100: st.global.v8.b32), optionally with an L2 evict_first hint
template <int MODE>
__global__ void store32_kernel(uint4* __restrict__ dst, size_t n32) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const uint4 v = make_uint4(threadIdx.x, blockIdx.x, 3, 4);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n32; i += stride) {
    if (MODE == 0)
      asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%1,%2,%3,%4};" ::"l"(dst + 2 * i), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
    else
      asm volatile("st.global.L1::no_allocate.L2::evict_first.v8.b32 [%0], {%1,%2,%3,%4,%1,%2,%3,%4};" ::"l"(dst + 2 * i), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
  }
}

// each CTA streams contiguous 64 KB blocks (like an activation tile), warp-contiguous 512 B per instruction
template <int MODE>
__global__ void tile_store_kernel(uint4* __restrict__ dst, size_t ntiles) {
  const uint4 v = make_uint4(threadIdx.x, blockIdx.x, 3, 4);
  for (size_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
    uint4* p = dst + t * 4096;   // 64 KB
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) {
      if (MODE == 0) p[i] = v;
      else __stcs(p + i, v);
    }
  }
}

// bulk (TMA) stores from shared memory: 16 KB per instruction
__global__ void bulk_store_kernel(uint8_t* __restrict__ dst, size_t nchunks, int chunk_bytes) {
  extern __shared__ __align__(128) uint8_t smem[];
  for (int i = threadIdx.x; i < chunk_bytes / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = i;
  __syncthreads();
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (threadIdx.x == 0) {
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
    int inflight = 0;
    for (size_t c = blockIdx.x; c < nchunks; c += gridDim.x) {
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + c * (size_t)chunk_bytes), "r"(s), "r"(chunk_bytes) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      if (++inflight >= 8) {
        asm volatile("cp.async.bulk.wait_group.read 4;" ::: "memory");
        inflight = 4;
      }
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
}

template <typename F>
float time_ms(F f, int reps) {
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  f();
  cudaDeviceSynchronize();
  cudaEventRecord(a);
  for (int i = 0; i < reps; ++i) f();
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms = 0;
  cudaEventElapsedTime(&ms, a, b);
  return ms / reps;
}

int main() {
  const size_t bytes = (size_t)4 << 30;
  uint8_t* buf;
  CK(cudaMalloc(&buf, bytes));
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const size_t n16 = bytes / 16;
  printf("{\"sms\": %d, \"bytes\": %zu", sms, bytes);
  float ms = time_ms([&] { cudaMemsetAsync(buf, 1, bytes); }, 5);
  printf(", \"memset_gbps\": %.1f", bytes / ms / 1e6);
  const int grids[3] = {sms * 2, sms * 8, sms * 32};
  const char* names[5] = {"st_default", "st_cs", "st_wt", "st_cg", "st_v8"};
  for (int g = 0; g < 3; ++g) {
    const int grid = grids[g];
    float r[5];
    r[0] = time_ms([&] { store_kernel<0><<<grid, 256>>>((uint4*)buf, n16); }, 5);
    r[1] = time_ms([&] { store_kernel<1><<<grid, 256>>>((uint4*)buf, n16); }, 5);
    r[2] = time_ms([&] { store_kernel<2><<<grid, 256>>>((uint4*)buf, n16); }, 5);
    r[3] = time_ms([&] { store_kernel<3><<<grid, 256>>>((uint4*)buf, n16); }, 5);
    r[4] = time_ms([&] { store32_kernel<0><<<grid, 256>>>((uint4*)buf, n16 / 2); }, 5);
    float r5 = time_ms([&] { store32_kernel<1><<<grid, 256>>>((uint4*)buf, n16 / 2); }, 5);
    printf(", \"st_v8_evict_first_grid%d_gbps\": %.1f", grid, bytes / r5 / 1e6);
    for (int m = 0; m < 5; ++m) printf(", \"%s_grid%d_gbps\": %.1f", names[m], grid, bytes / r[m] / 1e6);
  }
  ms = time_ms([&] { tile_store_kernel<0><<<sms, 256>>>((uint4*)buf, bytes / 65536); }, 5);
  printf(", \"tile64k_default_1cta_per_sm_gbps\": %.1f", bytes / ms / 1e6);
  ms = time_ms([&] { tile_store_kernel<1><<<sms, 256>>>((uint4*)buf, bytes / 65536); }, 5);
  printf(", \"tile64k_cs_1cta_per_sm_gbps\": %.1f", bytes / ms / 1e6);
  ms = time_ms([&] { tile_store_kernel<0><<<sms * 4, 256>>>((uint4*)buf, bytes / 65536); }, 5);
  printf(", \"tile64k_default_4cta_per_sm_gbps\": %.1f", bytes / ms / 1e6);
  for (int cb : {16384, 65536}) {
    cudaFuncSetAttribute(bulk_store_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, cb);
    ms = time_ms([&] { bulk_store_kernel<<<sms, 128, cb>>>(buf, bytes / cb, cb); }, 5);
    printf(", \"bulk_tma_%dk_1cta_per_sm_gbps\": %.1f", cb / 1024, bytes / ms / 1e6);
    ms = time_ms([&] { bulk_store_kernel<<<sms * 2, 128, cb>>>(buf, bytes / cb, cb); }, 5);
    printf(", \"bulk_tma_%dk_2cta_per_sm_gbps\": %.1f", cb / 1024, bytes / ms / 1e6);
  }
  cudaError_t e = cudaDeviceSynchronize();
  printf(", \"status\": \"%s\"}\n", cudaGetErrorString(e));
  return 0;
}
