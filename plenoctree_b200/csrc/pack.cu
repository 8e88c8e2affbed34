// pack.cu — flat fp32 parameters (reference order) -> tensor-core operand images.
//
// Flat layout of one MLP (matches the flax pytree MLP_i/Dense_0..Dense_9, kernels [in,out];
// octree/nerf/models.py:75-102 documents the Dense index <-> layer mapping):
//   Dense_0 63x256, Dense_1..4 256x256, Dense_5 319x256 ([h4 | posenc] rows), Dense_6..7,
//   Dense_8 256x1 (sigma), Dense_9 256x3K (rgb / SH coefficients, channel-major c*K+k).
//
// Forward images  (w_hi / w_lo): sequence of K-major SW64 slots [rows = out feature][32 k],
//   in the order mlp_fwd consumes them.  Heads rows are re-ordered to [sigma, (k, c) ...] so the
//   epilogue can index the SH basis with compile-time constants.
// Backward images (wt_hi): the transposed weights, slots [rows = in feature][32 out features],
//   in the order mlp_bwd consumes them (heads, then Dense_7 .. Dense_1).
#include "common.cuh"
#include "kernels.h"

namespace pob {

namespace {

struct PackArgs {
  const float* flat;
  FlatLayout L;
  int K, NH;
  uint8_t *w_hi, *w_lo, *wt_hi;
  float* bias;
};

__device__ __forceinline__ void put_hilo(uint8_t* hi, uint8_t* lo, size_t off, float v) {
  __half h = __float2half_rn(v);
  *reinterpret_cast<__half*>(hi + off) = h;
  if (lo) *reinterpret_cast<__half*>(lo + off) = __float2half_rn(v - __half2float(h));
}

// heads: packed output column n' -> value of W[in=i -> n'] and bias
__device__ __forceinline__ float heads_weight(const PackArgs& a, int i, int n) {
  if (n == 0) return a.flat[a.L.w_off[8] + i];  // Dense_8 kernel [256,1]
  const int k = (n - 1) / 3, c = (n - 1) % 3;
  if (k >= a.K) return 0.f;
  return a.flat[a.L.w_off[9] + i * (3 * a.K) + c * a.K + k];
}

__global__ void pack_weights_kernel(const __grid_constant__ PackArgs a) {
  const int NH = a.NH;
  const long long n_fwd_trunk = (long long)FWD_TRUNK_SLOTS * 256 * 32;
  const long long n_fwd_heads = (long long)FWD_HEAD_SLOTS * NH * 32;
  const int hs = (NH + 31) / 32;
  const long long n_bwd = (long long)(hs + 56) * 256 * 32;
  const long long n_bias = 8 * 256 + MAX_NH;
  const long long total = n_fwd_trunk + n_fwd_heads + n_bwd + n_bias;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    if (t < n_fwd_trunk) {
      const int slot = int(t / (256 * 32));
      const int n = int(t / 32) % 256, kk = int(t % 32);
      // slot -> (layer, j)
      int l = 0, j = slot;
      while (j >= fwd_slots_of_layer(l)) {
        j -= fwd_slots_of_layer(l);
        ++l;
      }
      const int kin = 32 * j + kk;  // input feature index in the layer's [in] axis
      float v = 0.f;
      if (fwd_has_bias_slot(l) && j == 8) {
        if (kk == 31) v = a.flat[a.L.b_off[l] + n];            // bias slot: k = 31 <-> posenc column 63 (= 1)
      } else if (kin < a.L.in_dim[l]) {
        v = a.flat[a.L.w_off[l] + kin * 256 + n];
      } else if (kin == a.L.in_dim[l] && !fwd_has_bias_slot(l)) {
        v = a.flat[a.L.b_off[l] + n];                          // layers 0 / 5: the padding row k = 63 of posenc
      }
      put_hilo(a.w_hi, a.w_lo, size_t(slot) * WSLOT_BYTES + w_slot_offset(n, kk), v);
    } else if (t < n_fwd_trunk + n_fwd_heads) {
      const long long u = t - n_fwd_trunk;
      const int j = int(u / (NH * 32));
      const int n = int(u / 32) % NH, kk = int(u % 32);
      float v;
      if (j < 8) v = heads_weight(a, 32 * j + kk, n);
      else {
        v = 0.f;
        if (kk == 31) {
          if (n == 0) v = a.flat[a.L.b_off[8]];
          else {
            const int k = (n - 1) / 3, c = (n - 1) % 3;
            if (k < a.K) v = a.flat[a.L.b_off[9] + c * a.K + k];
          }
        }
      }
      put_hilo(a.w_hi, a.w_lo, size_t(FWD_TRUNK_SLOTS) * WSLOT_BYTES + size_t(j) * NH * 64 + w_slot_offset(n, kk), v);
    } else if (t < n_fwd_trunk + n_fwd_heads + n_bwd) {
      const long long u = t - n_fwd_trunk - n_fwd_heads;
      const int slot = int(u / (256 * 32));
      const int i = int(u / 32) % 256, kk = int(u % 32);  // i = in feature (row), kk = out feature
      float v = 0.f;
      if (slot < hs) {
        const int n = 32 * slot + kk;
        if (n < NH) v = heads_weight(a, i, n);
      } else {
        const int l = 7 - (slot - hs) / 8;  // Dense_7 .. Dense_1
        const int ko = 32 * ((slot - hs) % 8) + kk;
        v = a.flat[a.L.w_off[l] + i * 256 + ko];
      }
      put_hilo(a.wt_hi, nullptr, size_t(slot) * WSLOT_BYTES + w_slot_offset(i, kk), v);
    } else {
      const int b = int(t - n_fwd_trunk - n_fwd_heads - n_bwd);
      float v = 0.f;
      if (b < 8 * 256) {
        v = a.flat[a.L.b_off[b / 256] + (b % 256)];
      } else {
        const int n = b - 8 * 256;
        if (n == 0) v = a.flat[a.L.b_off[8]];
        else {
          const int k = (n - 1) / 3, c = (n - 1) % 3;
          if (k < a.K) v = a.flat[a.L.b_off[9] + c * a.K + k];
        }
      }
      a.bias[b] = v;
    }
  }
}

}  // namespace

cudaError_t launch_pack_weights(const float* flat, int K, uint8_t* w_hi, uint8_t* w_lo,
                                uint8_t* wt_hi, float* bias, cudaStream_t stream) {
  PackArgs a;
  a.flat = flat;
  a.L = flat_layout(K);
  a.K = K;
  a.NH = heads_width(K);
  a.w_hi = w_hi;
  a.w_lo = w_lo;
  a.wt_hi = wt_hi;
  a.bias = bias;
  pack_weights_kernel<<<592, 256, 0, stream>>>(a);
  return cudaGetLastError();
}

}  // namespace pob
