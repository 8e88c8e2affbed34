// render.cu — per-ray stages of NerfModel.__call__ that are not GEMMs: stratified sampling,
// volumetric alpha-compositing (forward and backward), inverse-CDF hierarchical resampling with the
// union sort, MSE loss gradient.  One warp per ray, all fp32, warp-shuffle scans / reductions.
//
//   sample_along_rays       nerf_sh/nerf/model_utils.py:104-142
//   volumetric_rendering    nerf_sh/nerf/model_utils.py:176-222
//   piecewise_constant_pdf  nerf_sh/nerf/model_utils.py:225-286
//   sample_pdf              nerf_sh/nerf/model_utils.py:289-314
//   loss_fn (MSE part)      nerf_sh/train.py:86-96
#include "common.cuh"
#include "kernels.h"

namespace pob {

namespace {

constexpr int RAYS_PER_BLOCK = 4;
constexpr int MAX_SEG = 8;  // samples per lane  (N <= 256)
constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}
// exclusive product scan across lanes
__device__ __forceinline__ float warp_excl_prod(float v, int lane) {
  float inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    float t = __shfl_up_sync(FULL, inc, o);
    if (lane >= o) inc *= t;
  }
  float ex = __shfl_up_sync(FULL, inc, 1);
  return lane == 0 ? 1.f : ex;
}
// exclusive sum scan across lanes
__device__ __forceinline__ float warp_excl_sum(float v, int lane) {
  float inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    float t = __shfl_up_sync(FULL, inc, o);
    if (lane >= o) inc += t;
  }
  float ex = __shfl_up_sync(FULL, inc, 1);
  return lane == 0 ? 0.f : ex;
}
// exclusive suffix sum across lanes (sum over lanes > me)
__device__ __forceinline__ float warp_excl_suffix_sum(float v, int lane) {
  float inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    float t = __shfl_down_sync(FULL, inc, o);
    if (lane + o < 32) inc += t;
  }
  float ex = __shfl_down_sync(FULL, inc, 1);
  return lane == 31 ? 0.f : ex;
}

// ---------------------------------------------------------------------------------------------
// Stratified sampling.  z_base[N] = near*(1-t)+far*t (or the lindisp form) is tabulated by the host
// with the reference's own expression so that no linspace rounding ambiguity enters.
// ---------------------------------------------------------------------------------------------
__global__ void sample_coarse_kernel(const float* __restrict__ z_base, const float* __restrict__ t_rand,
                                     int R, int N, float* __restrict__ z_out) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= (long long)R * N) return;
  const int j = int(i % N);
  float z = z_base[j];
  if (t_rand) {
    const float zl = j > 0 ? z_base[j - 1] : z, zu = j < N - 1 ? z_base[j + 1] : z;
    const float lower = j > 0 ? __fmul_rn(0.5f, __fadd_rn(z, zl)) : z;
    const float upper = j < N - 1 ? __fmul_rn(0.5f, __fadd_rn(zu, z)) : z;
    z = __fadd_rn(lower, __fmul_rn(__fsub_rn(upper, lower), t_rand[i]));
  }
  z_out[i] = z;
}

// ---------------------------------------------------------------------------------------------
// Per-ray recomputation shared by the forward and backward compositing kernels
// ---------------------------------------------------------------------------------------------
struct RaySeg {
  float4 c[MAX_SEG];   // (r,g,b,sigma) of my samples
  float z[MAX_SEG];
  float alpha[MAX_SEG], om[MAX_SEG], dist[MAX_SEG];
  float T0;            // transmittance in front of my first sample
};

template <int S>
__device__ __forceinline__ void load_ray(const float4* __restrict__ rgbs, const float* __restrict__ z,
                                         const float* __restrict__ dirs, long long ray, int N, int lane,
                                         RaySeg& r) {
  const float dx = dirs[3 * ray], dy = dirs[3 * ray + 1], dz = dirs[3 * ray + 2];
  const float dnorm = sqrtf(dx * dx + dy * dy + dz * dz);
  float prod = 1.f;
#pragma unroll
  for (int i = 0; i < S; ++i) {
    const int idx = lane * S + i;
    if (idx < N) {
      r.c[i] = rgbs[ray * N + idx];
      r.z[i] = z[ray * N + idx];
      const float zn = (idx + 1 < N) ? z[ray * N + idx + 1] : 0.f;
      float d = (idx + 1 < N) ? __fsub_rn(zn, r.z[i]) : 1e10f;
      d = __fmul_rn(d, dnorm);
      r.dist[i] = d;
      r.alpha[i] = 1.0f - expf(-r.c[i].w * d);
      r.om[i] = (1.0f - r.alpha[i]) + 1e-10f;
      prod *= r.om[i];
    } else {
      r.c[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      r.z[i] = 0.f;
      r.dist[i] = 0.f;
      r.alpha[i] = 0.f;
      r.om[i] = 1.f;
    }
  }
  r.T0 = warp_excl_prod(prod, lane);
}

struct CompositeArgs {
  const float4* rgbs;
  const float* z;
  const float* dirs;
  int R, N, white_bkgd;
  float *out_rgb, *out_disp, *out_acc, *out_weights;
};

template <int S>
__global__ void __launch_bounds__(RAYS_PER_BLOCK * 32)
composite_fwd_kernel(const CompositeArgs a) {
  const int lane = threadIdx.x & 31;
  const long long ray = blockIdx.x * (long long)RAYS_PER_BLOCK + (threadIdx.x >> 5);
  if (ray >= a.R) return;
  RaySeg r;
  load_ray<S>(a.rgbs, a.z, a.dirs, ray, a.N, lane, r);
  float T = r.T0, cr = 0.f, cg = 0.f, cb = 0.f, depth = 0.f, acc = 0.f;
#pragma unroll
  for (int i = 0; i < S; ++i) {
    const int idx = lane * S + i;
    const float w = r.alpha[i] * T;
    if (idx < a.N) {
      cr += w * r.c[i].x;
      cg += w * r.c[i].y;
      cb += w * r.c[i].z;
      depth += w * r.z[i];
      acc += w;
      if (a.out_weights) a.out_weights[ray * a.N + idx] = w;
    }
    T *= r.om[i];
  }
  cr = warp_sum(cr);
  cg = warp_sum(cg);
  cb = warp_sum(cb);
  depth = warp_sum(depth);
  acc = warp_sum(acc);
  if (lane == 0) {
    const float inv_eps = 1e10f;
    float disp = acc / depth;
    disp = (disp > 0.f && disp < inv_eps && acc > 1e-10f) ? disp : inv_eps;
    if (a.white_bkgd) {
      cr += 1.0f - acc;
      cg += 1.0f - acc;
      cb += 1.0f - acc;
    }
    a.out_rgb[3 * ray] = cr;
    a.out_rgb[3 * ray + 1] = cg;
    a.out_rgb[3 * ray + 2] = cb;
    if (a.out_disp) a.out_disp[ray] = disp;
    if (a.out_acc) a.out_acc[ray] = acc;
  }
}

// ---------------------------------------------------------------------------------------------
// Backward of compositing + MSE.  Emits, per sample, the gradient w.r.t. the PRE-activation head
// outputs after SH evaluation: (d pre_r, d pre_g, d pre_b, d sigma_raw), already multiplied by
// sigmoid' and relu' — mlp_bwd only has to expand it with the SH basis.
//   dL/dC = gscale * (C - px);   gscale = loss_scale * 2 / (3 * R_global_per_rank)
// ---------------------------------------------------------------------------------------------
struct CompositeBwdArgs {
  const float4* rgbs;
  const float* z;
  const float* dirs;
  const float* comp_rgb;   // [R,3] forward result
  const float* pixels;     // [R,3]
  int R, N, white_bkgd;
  float gscale;
  float4* G;               // [R,N]
  float* sq_err_sum;       // += sum_{rays,ch} (C - px)^2   (loss numerator)
};

template <int S>
__global__ void __launch_bounds__(RAYS_PER_BLOCK * 32)
composite_bwd_kernel(const CompositeBwdArgs a) {
  const int lane = threadIdx.x & 31;
  const long long ray = blockIdx.x * (long long)RAYS_PER_BLOCK + (threadIdx.x >> 5);
  if (ray >= a.R) return;
  RaySeg r;
  load_ray<S>(a.rgbs, a.z, a.dirs, ray, a.N, lane, r);
  const float ex = a.comp_rgb[3 * ray] - a.pixels[3 * ray];
  const float ey = a.comp_rgb[3 * ray + 1] - a.pixels[3 * ray + 1];
  const float ez = a.comp_rgb[3 * ray + 2] - a.pixels[3 * ray + 2];
  const float dcx = a.gscale * ex, dcy = a.gscale * ey, dcz = a.gscale * ez;
  const float bg = a.white_bkgd ? 1.f : 0.f;
  // pass 1: weights, g_i * w_i and its suffix sums
  float T = r.T0;
  float w[S], Tpre[S], gi[S];
  float local = 0.f;
#pragma unroll
  for (int i = 0; i < S; ++i) {
    Tpre[i] = T;
    w[i] = r.alpha[i] * T;
    gi[i] = dcx * (r.c[i].x - bg) + dcy * (r.c[i].y - bg) + dcz * (r.c[i].z - bg);
    local += gi[i] * w[i];
    T *= r.om[i];
  }
  float suffix = warp_excl_suffix_sum(local, lane);  // sum over lanes > me
#pragma unroll
  for (int i = S - 1; i >= 0; --i) {
    const int idx = lane * S + i;
    if (idx < a.N) {
      // dL/dalpha_i = g_i T_i - (sum_{k>i} g_k w_k) / (1 - alpha_i + eps)
      const float dalpha = gi[i] * Tpre[i] - suffix / r.om[i];
      const float dsigma = dalpha * r.dist[i] * (1.0f - r.alpha[i]);
      float4 g;
      g.x = w[i] * dcx * r.c[i].x * (1.0f - r.c[i].x);
      g.y = w[i] * dcy * r.c[i].y * (1.0f - r.c[i].y);
      g.z = w[i] * dcz * r.c[i].z * (1.0f - r.c[i].z);
      g.w = r.c[i].w > 0.f ? dsigma : 0.f;
      a.G[ray * a.N + idx] = g;
    }
    suffix += gi[i] * w[i];
  }
  if (lane == 0 && a.sq_err_sum) atomicAdd(a.sq_err_sum, ex * ex + ey * ey + ez * ez);
}

// ---------------------------------------------------------------------------------------------
// Hierarchical resampling: inverse-CDF sampling of Nf new depths from the interior coarse weights,
// then the sorted union with the Nc coarse depths.
// ---------------------------------------------------------------------------------------------
struct PdfArgs {
  const float* z_c;       // [R,Nc]
  const float* weights;   // [R,Nc]
  const float* u;         // [Nf] table (u_per_ray = 0) or [R,Nf]
  int u_per_ray;
  int R, Nc, Nf;
  float* z_out;           // [R, Nc+Nf]
};

__global__ void __launch_bounds__(RAYS_PER_BLOCK * 32) sample_pdf_kernel(const PdfArgs a) {
  __shared__ float s_bins[RAYS_PER_BLOCK][256];
  __shared__ float s_cdf[RAYS_PER_BLOCK][256];
  __shared__ float s_sort[RAYS_PER_BLOCK][256];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const long long ray = blockIdx.x * (long long)RAYS_PER_BLOCK + wid;
  if (ray >= a.R) return;
  const int Nc = a.Nc, Nf = a.Nf;
  const int nb = Nc - 1;   // bins (mid points): 63
  const int nw = Nc - 2;   // interior weights:   62
  float* bins = s_bins[wid];
  float* cdf = s_cdf[wid];
  float* sb = s_sort[wid];
  const float* zc = a.z_c + ray * Nc;
  const float* wt = a.weights + ray * Nc;
  for (int i = lane; i < nb; i += 32) bins[i] = __fmul_rn(0.5f, __fadd_rn(zc[i + 1], zc[i]));
  // weights[..., 1:-1], padded so that the sum is at least eps
  const int S = (nw + 31) / 32;
  float wl[MAX_SEG];
  float lsum = 0.f;
#pragma unroll
  for (int i = 0; i < MAX_SEG; ++i) {
    const int idx = lane * S + i;
    wl[i] = (i < S && idx < nw) ? wt[idx + 1] : 0.f;
    lsum += wl[i];
  }
  float wsum = warp_sum(lsum);
  const float padding = fmaxf(0.f, 1e-5f - wsum);
  const float padw = padding / float(nw);
  wsum += padding;
  // pdf and inclusive cumsum over pdf[:-1]
  float run = 0.f;
#pragma unroll
  for (int i = 0; i < MAX_SEG; ++i) {
    const int idx = lane * S + i;
    if (i < S && idx < nw) {
      wl[i] = (wl[i] + padw) / wsum;
      run += wl[i];
    } else {
      wl[i] = 0.f;
    }
  }
  float pre = warp_excl_sum(run, lane);
#pragma unroll
  for (int i = 0; i < MAX_SEG; ++i) {
    const int idx = lane * S + i;
    if (i < S && idx < nw) {
      pre += wl[i];
      if (idx < nw - 1) cdf[idx + 1] = fminf(1.f, pre);
    }
  }
  if (lane == 0) {
    cdf[0] = 0.f;
    cdf[nb - 1] = 1.f;
  }
  __syncwarp();
  // union buffer: coarse depths first
  for (int i = lane; i < Nc; i += 32) sb[i] = zc[i];
  for (int j = lane; j < Nf; j += 32) {
    const float u = a.u_per_ray ? a.u[ray * Nf + j] : a.u[j];
    // count of cdf entries <= u (cdf is non-decreasing): upper bound
    int lo = 0, hi = nb;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (cdf[mid] <= u) lo = mid + 1; else hi = mid;
    }
    const int i0 = lo > 0 ? lo - 1 : 0;
    const int i1 = lo < nb ? lo : nb - 1;
    const float c0 = cdf[i0], c1 = cdf[i1];
    float t = __fdiv_rn(__fsub_rn(u, c0), __fsub_rn(c1, c0));
    if (t != t) t = 0.f;                      // nan_to_num
    t = fminf(fmaxf(t, 0.f), 1.f);            // +-inf clip like the reference
    const float b0 = bins[i0], b1 = bins[i1];
    sb[Nc + j] = __fadd_rn(b0, __fmul_rn(t, __fsub_rn(b1, b0)));
  }
  for (int i = Nc + Nf + lane; i < 256; i += 32) sb[i] = __int_as_float(0x7f800000);
  __syncwarp();
  // bitonic sort of 256 keys, 4 compare-exchanges per lane per pass
  for (int k = 2; k <= 256; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int t = lane + 32 * q;                // 0..127: index of the compare-exchange
        const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
        const int p = i | j;
        const bool up = (i & k) == 0;
        const float x = sb[i], y = sb[p];
        if ((x > y) == up) {
          sb[i] = y;
          sb[p] = x;
        }
      }
      __syncwarp();
    }
  }
  float* zo = a.z_out + ray * (long long)(Nc + Nf);
  for (int i = lane; i < Nc + Nf; i += 32) zo[i] = sb[i];
}

// sparsity-loss gradient (nerf_sh/train.py:77-83): G.w = coef * exp(-len * relu(s)) * [s > 0]; the forward
// epilogue has already applied the relu (rgbs.w)
__global__ void sparsity_grad_kernel(const float4* __restrict__ rgbs, int n, float length, float coef,
                                     float4* __restrict__ G, float* __restrict__ exp_sum) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  float e = 0.f;
  if (i < n) {
    const float s = rgbs[i].w;
    e = expf(-length * s);
    G[i] = make_float4(0.f, 0.f, 0.f, s > 0.f ? coef * e : 0.f);
  }
  e = warp_sum(e);
  if ((threadIdx.x & 31) == 0 && exp_sum) atomicAdd(exp_sum, e);
}

// ---- random draws of one training step (stratified jitter, inverse-CDF uniforms, sparsity points) ----------------
// Philox4x32-10 (Salmon et al., SC'11), counter = (element / 4, stream id, step), key = seed: one launch replaces
// the seven ATen launches (3 x rand + scale / shift) of a step.  The reference draws from jax.random's threefry
// streams, which cannot be reproduced without JAX; parity tests inject their draws instead (SURVEY.md 7.2 RNG).
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, ctr.x), lo0 = 0xD2511F53u * ctr.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, ctr.z), lo1 = 0xCD9E8D57u * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += 0x9E3779B9u;
    key.y += 0xBB67AE85u;
  }
  return ctr;
}
// 24 random bits -> [0, 1): never 1, like random.uniform
__device__ __forceinline__ float u01(uint32_t x) { return float(x >> 8) * (1.0f / 16777216.0f); }

__global__ void draw_uniforms_kernel(unsigned long long seed, float step_host, const float* __restrict__ step_dev,
                                     float* __restrict__ t_rand, long long n_t, float* __restrict__ u, long long n_u,
                                     float* __restrict__ sp, long long n_sp, float sp_radius) {
  const uint32_t step = uint32_t(step_dev ? __ldg(step_dev) : step_host);
  const long long q4 = (n_t + 3) / 4, r4 = (n_u + 3) / 4, s4 = (n_sp + 3) / 4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < q4 + r4 + s4;
       i += (long long)gridDim.x * blockDim.x) {
    int stream = 0;
    long long j = i;
    float* dst = t_rand;
    long long n = n_t;
    if (j >= q4) { j -= q4; stream = 1; dst = u; n = n_u; }
    if (stream == 1 && j >= r4) { j -= r4; stream = 2; dst = sp; n = n_sp; }
    const uint4 r = philox4x32_10(make_uint4(uint32_t(j), uint32_t(j >> 32), uint32_t(stream), step),
                                  make_uint2(uint32_t(seed), uint32_t(seed >> 32)));
    float v[4] = {u01(r.x), u01(r.y), u01(r.z), u01(r.w)};
    if (stream == 2) {   // random.uniform(key, (npoints, 3), minval=-radius, maxval=radius)  (train.py:79)
#pragma unroll
      for (int k = 0; k < 4; ++k) v[k] = fmaf(v[k], 2.0f * sp_radius, -sp_radius);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (4 * j + k < n) dst[4 * j + k] = v[k];
  }
}

template <typename F>
cudaError_t dispatch_seg(int N, F&& f) {
  const int S = (N + 31) / 32;
  switch (S) {
    case 1: return f(std::integral_constant<int, 1>());
    case 2: return f(std::integral_constant<int, 2>());
    case 3: return f(std::integral_constant<int, 3>());
    case 4: return f(std::integral_constant<int, 4>());
    case 5: return f(std::integral_constant<int, 5>());
    case 6: return f(std::integral_constant<int, 6>());
    case 7: return f(std::integral_constant<int, 7>());
    case 8: return f(std::integral_constant<int, 8>());
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace

cudaError_t launch_sample_coarse(const float* z_base, const float* t_rand, int R, int N, float* z_out,
                                 cudaStream_t st) {
  const long long n = (long long)R * N;
  if (n == 0) return cudaSuccess;
  sample_coarse_kernel<<<unsigned((n + 255) / 256), 256, 0, st>>>(z_base, t_rand, R, N, z_out);
  return cudaGetLastError();
}

cudaError_t launch_composite_fwd(const float4* rgbs, const float* z, const float* dirs, int R, int N,
                                 int white_bkgd, float* out_rgb, float* out_disp, float* out_acc,
                                 float* out_weights, cudaStream_t st) {
  if (R == 0) return cudaSuccess;
  CompositeArgs a{rgbs, z, dirs, R, N, white_bkgd, out_rgb, out_disp, out_acc, out_weights};
  const unsigned grid = (R + RAYS_PER_BLOCK - 1) / RAYS_PER_BLOCK;
  return dispatch_seg(N, [&](auto s) {
    composite_fwd_kernel<decltype(s)::value><<<grid, RAYS_PER_BLOCK * 32, 0, st>>>(a);
    return cudaGetLastError();
  });
}

cudaError_t launch_composite_bwd(const float4* rgbs, const float* z, const float* dirs,
                                 const float* comp_rgb, const float* pixels, int R, int N, int white_bkgd,
                                 float gscale, float4* G, float* sq_err_sum, cudaStream_t st) {
  if (R == 0) return cudaSuccess;
  CompositeBwdArgs a{rgbs, z, dirs, comp_rgb, pixels, R, N, white_bkgd, gscale, G, sq_err_sum};
  const unsigned grid = (R + RAYS_PER_BLOCK - 1) / RAYS_PER_BLOCK;
  return dispatch_seg(N, [&](auto s) {
    composite_bwd_kernel<decltype(s)::value><<<grid, RAYS_PER_BLOCK * 32, 0, st>>>(a);
    return cudaGetLastError();
  });
}

cudaError_t launch_sample_pdf(const float* z_c, const float* weights, const float* u, int u_per_ray, int R,
                              int Nc, int Nf, float* z_out, cudaStream_t st) {
  if (R == 0) return cudaSuccess;
  if (Nc < 3 || Nc + Nf > 256 || Nc - 2 > 32 * MAX_SEG) return cudaErrorInvalidValue;
  PdfArgs a{z_c, weights, u, u_per_ray, R, Nc, Nf, z_out};
  const unsigned grid = (R + RAYS_PER_BLOCK - 1) / RAYS_PER_BLOCK;
  sample_pdf_kernel<<<grid, RAYS_PER_BLOCK * 32, 0, st>>>(a);
  return cudaGetLastError();
}

cudaError_t launch_draw_uniforms(unsigned long long seed, float step, const float* step_dev, float* t_rand,
                                 long long n_t, float* u, long long n_u, float* sp, long long n_sp, float sp_radius,
                                 cudaStream_t st) {
  const long long work = (n_t + 3) / 4 + (n_u + 3) / 4 + (n_sp + 3) / 4;
  if (work == 0) return cudaSuccess;
  const int grid = int((work + 255) / 256 < 2048 ? (work + 255) / 256 : 2048);
  draw_uniforms_kernel<<<grid, 256, 0, st>>>(seed, step, step_dev, t_rand, n_t, u, n_u, sp, n_sp, sp_radius);
  return cudaGetLastError();
}

cudaError_t launch_sparsity_grad(const float4* rgbs, int n, float length, float coef, float4* G,
                                 float* exp_sum, cudaStream_t st) {
  if (n == 0) return cudaSuccess;
  sparsity_grad_kernel<<<(n + 255) / 256, 256, 0, st>>>(rgbs, n, length, coef, G, exp_sum);
  return cudaGetLastError();
}

}  // namespace pob
