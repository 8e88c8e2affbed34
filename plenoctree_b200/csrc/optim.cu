// optim.cu — gradient finalisation and the optimiser step.
//   reduce_grads : per-CTA wgrad partials -> flat fp32 gradient in reference parameter order
//   adam         : flax.optim.Adam.apply_gradient (nerf_sh/train.py:119, models.py:44)
#include "common.cuh"
#include "kernels.h"

namespace pob {

// Work split of the wgrad launch.  The kernel is bound by the tile bytes each role streams
// (128 KB per tile for the 256x256 layers, ~80-96 KB for Dense_0 / the skip rows / the heads).
int wgrad_assign_roles(WgradParams& p, int n_in, int role_start[WG_NUM_ROLES],
                       int role_count[WG_NUM_ROLES]) {
  int n = n_in < WG_MAX_CTAS ? n_in : WG_MAX_CTAS;
  if (n < WG_NUM_ROLES) n = WG_NUM_ROLES;  // one CTA per role at the very least (they time-share SMs)
  int small = (n * 8) / 100;               // per small role
  if (small < 1) small = 1;
  int big = (n - 3 * small) / 7;
  if (big < 1) big = 1;
  small = (n - 7 * big) / 3;
  if (small < 1) small = 1;
  int cta = 0;
  for (int r = 0; r < WG_NUM_ROLES; ++r) {
    const int c = r < 7 ? big : small;
    role_start[r] = cta;
    role_count[r] = c;
    for (int i = 0; i < c; ++i, ++cta) {
      p.cta_role[cta] = short(r);
      p.cta_index[cta] = short(i);
      p.cta_count[cta] = short(c);
    }
  }
  for (int i = cta; i < WG_MAX_CTAS; ++i) {   // spare CTAs idle (role -1)
    p.cta_role[i] = -1;
    p.cta_index[i] = 0;
    p.cta_count[i] = 1;
  }
  return cta;
}

namespace {

struct ReduceArgs {
  const float* partials;
  const float* partials2;   // optional second launch with the same role tables (sparsity level)
  int role_start[WG_NUM_ROLES], role_count[WG_NUM_ROLES];
  FlatLayout L;
  int K, NH;
  float inv_scale;
  float* grad;
};

__global__ void reduce_grads_kernel(const __grid_constant__ ReduceArgs a) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= a.L.total) return;
  // locate (layer, kernel|bias, local index)
  int layer = 0;
  while (layer < 9 && e >= a.L.w_off[layer + 1]) ++layer;
  const bool is_bias = e >= a.L.b_off[layer];
  int role, off;
  if (!is_bias) {
    const int local = e - a.L.w_off[layer];
    const int out_dim = a.L.out_dim[layer];
    const int i = local / out_dim, o = local % out_dim;  // kernel [in, out]
    if (layer == 0) {
      role = 7;
      off = o * 64 + i;
    } else if (layer < 8) {
      if (layer == 5 && i >= 256) {
        role = 8;
        off = o * 64 + (i - 256);
      } else {
        role = layer <= 4 ? layer - 1 : (layer == 5 ? 4 : layer - 1);
        off = o * 256 + i;
      }
    } else {
      role = 9;  // heads: D[in feature][packed column]
      int n;
      if (layer == 8) n = 0;
      else {
        const int c = o / a.K, k = o % a.K;
        n = 1 + 3 * k + c;
      }
      off = i * a.NH + n;
    }
  } else {
    const int o = e - a.L.b_off[layer];
    if (layer < 8) {
      role = layer == 0 ? 7 : (layer <= 4 ? layer - 1 : (layer == 5 ? 4 : layer - 1));
      off = 65536 + o;
    } else {
      role = 9;
      int n;
      if (layer == 8) n = 0;
      else {
        const int c = o / a.K, k = o % a.K;
        n = 1 + 3 * k + c;
      }
      off = 65536 + n;
    }
  }
  float s = 0.f;
  const float* p = a.partials + size_t(a.role_start[role]) * WG_PARTIAL_FLOATS + off;
  for (int c = 0; c < a.role_count[role]; ++c) s += p[size_t(c) * WG_PARTIAL_FLOATS];
  if (a.partials2) {
    const float* p2 = a.partials2 + size_t(a.role_start[role]) * WG_PARTIAL_FLOATS + off;
    for (int c = 0; c < a.role_count[role]; ++c) s += p2[size_t(c) * WG_PARTIAL_FLOATS];
  }
  a.grad[e] = s * a.inv_scale;
}

__global__ void adam_kernel(float* __restrict__ param, const float* __restrict__ grad, float* __restrict__ m,
                            float* __restrict__ v, long long n, float lr, float bc1, float bc2,
                            const float* __restrict__ lr_step, float beta1, float beta2, float eps, float grad_mult,
                            float wd) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (lr_step) {   // learning rate and step count live on the device (replayable CUDA graphs)
    lr = __ldg(lr_step);
    const float t = __ldg(lr_step + 1) + 1.0f;
    bc1 = 1.0f - powf(beta1, t);
    bc2 = 1.0f - powf(beta2, t);
  }
  const float p = param[i];
  const float g = grad[i] * grad_mult + wd * p;
  const float mi = (1.0f - beta1) * g + beta1 * m[i];
  const float vi = (1.0f - beta2) * g * g + beta2 * v[i];
  m[i] = mi;
  v[i] = vi;
  const float mh = mi / bc1, vh = vi / bc2;
  param[i] = p - lr * mh / (sqrtf(vh) + eps);
}

}  // namespace

cudaError_t launch_reduce_grads(const float* partials, const int role_start[WG_NUM_ROLES],
                                const int role_count[WG_NUM_ROLES], int K, float inv_scale,
                                float* grad_flat, cudaStream_t stream, const float* partials2) {
  ReduceArgs a;
  a.partials = partials;
  a.partials2 = partials2;
  for (int r = 0; r < WG_NUM_ROLES; ++r) {
    a.role_start[r] = role_start[r];
    a.role_count[r] = role_count[r];
  }
  a.L = flat_layout(K);
  a.K = K;
  a.NH = heads_width(K);
  a.inv_scale = inv_scale;
  a.grad = grad_flat;
  reduce_grads_kernel<<<(a.L.total + 255) / 256, 256, 0, stream>>>(a);
  return cudaGetLastError();
}

cudaError_t launch_adam(float* param, const float* grad, float* m, float* v, long long n, float lr,
                        float step, const float* lr_step_dev, float beta1, float beta2, float eps, float grad_mult,
                        float weight_decay_coef, cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  const double t = double(step) + 1.0;
  const float bc1 = float(1.0 - pow(double(beta1), t));
  const float bc2 = float(1.0 - pow(double(beta2), t));
  adam_kernel<<<unsigned((n + 255) / 256), 256, 0, stream>>>(param, grad, m, v, n, lr, bc1, bc2, lr_step_dev,
                                                              beta1, beta2, eps, grad_mult, weight_decay_coef);
  return cudaGetLastError();
}

}  // namespace pob
