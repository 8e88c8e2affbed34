#pragma once
// wgrad_body.cuh — device body of mlp_wgrad.cu: weight-gradient contraction over samples (the wgrad half of jax.value_and_grad,
// nerf_sh/train.py:116):   dW_l[out, in] = sum_s dZ_l[s, out] * h_{l-1}[s, in],  db_l = sum_s dZ_l[s, :]
//
// A 256x256 fp32 accumulator is exactly one SM's tensor memory (2 x 128 lanes x 256 columns), so
// every persistent CTA owns ONE layer ("role") for the whole launch and streams the [sample x
// feature] tile images that mlp_fwd (h_l, posenc) and mlp_bwd (dZ_l, dO) left in global memory.
// Both MMA operands are read MN-major straight from those images (K = samples): no transposes.
// dZ / dO / posenc tiles are K-major SW128 images (read MN-major with the same swizzle), the h_l tiles
// are "T" images (no swizzle, 128 B core matrices; tests/test_umma_probe.py pins both conventions).
// CTAs of the same role split the tiles round-robin and each writes an fp32 partial; reduce_grads
// (optim.cu) sums the partials into the flat gradient (deterministic, no atomics).
//
// Roles: 0..6 = Dense_1,2,3,4,5(h4 rows),6,7   A = dZ_l (256 out)  B = h_{l-1} (256 in)
//        7    = Dense_0                         A = dZ_0            B = posenc   (64)
//        8    = Dense_5 (posenc rows)           A = dZ_5            B = posenc   (64)
//        9    = heads (Dense_8 | Dense_9)       A = h_7 (256 in)    B = dO       (NH)  [transposed result]
// Bias gradients are column sums of the A (or, for the heads, B) tile, computed by four otherwise
// idle warps from the staged shared-memory tiles.
#include "common.cuh"
#include "kernels.h"

namespace pob {

namespace {

constexpr int WG_THREADS = 192;       // warps 0-3: bias sums + final drain, warp 4: loads, warp 5: MMA
constexpr int WG_STAGES = 3;
constexpr int WG_SUB = 64;            // samples per stage
constexpr uint32_t WG_HALF = 4 * WG_SUB * 128;     // one operand sub-image: 4 chunks x 64 rows x 128 B
constexpr uint32_t WG_STAGE_BYTES = 2 * WG_HALF;   // 64 KB
constexpr uint32_t WG_SMEM = WG_STAGES * WG_STAGE_BYTES;

struct WgBarriers {
  uint64_t full[WG_STAGES];
  uint64_t empty[WG_STAGES];
  uint64_t done;
};

struct RoleInfo {
  int a_kind, a_layer;   // 0: dZ[layer], 1: H[layer]
  int b_kind, b_layer;   // 0: H[layer], 1: E, 2: dO
  int b_chunks, N;
  int bias_from_b;       // heads: bias = column sums of dO
  int has_bias;
  int a_t, b_t;          // operand is a forward-saved h tile in the T layout (layouts.py: t_tile_offset)
};

__device__ __forceinline__ RoleInfo role_info(int role, int NH) {
  RoleInfo r;
  r.bias_from_b = 0;
  r.has_bias = 1;
  r.a_t = 0;
  r.b_t = 0;
  if (role < 7) {
    r.b_t = 1;
    const int l = role < 4 ? role + 1 : (role == 4 ? 5 : role + 1);  // 1,2,3,4,5,6,7
    r.a_kind = 0; r.a_layer = l; r.b_kind = 0; r.b_layer = l - 1; r.b_chunks = 4; r.N = 256;
  } else if (role == 7) {
    r.a_kind = 0; r.a_layer = 0; r.b_kind = 1; r.b_layer = 0; r.b_chunks = 1; r.N = 64;
  } else if (role == 8) {
    r.a_kind = 0; r.a_layer = 5; r.b_kind = 1; r.b_layer = 0; r.b_chunks = 1; r.N = 64; r.has_bias = 0;
  } else {
    r.a_kind = 1; r.a_layer = 7; r.b_kind = 2; r.b_layer = 0; r.b_chunks = (NH + 63) / 64; r.N = NH;
    r.bias_from_b = 1;
    r.a_t = 1;
  }
  return r;
}

}  // namespace


// One unit of work = one 128-sample tile (two 64-sample stages).
struct WgItem {
  const uint8_t* a_ptr;
  const uint8_t* b_ptr;
};

// cta indexes cta_role/index/count and the partials
__device__ __forceinline__ void wgrad_body(const WgradParams& p, uint8_t* smem, const int cta) {
  __shared__ __align__(8) WgBarriers bars;
  __shared__ uint32_t tmem_base_s;

  const uint32_t warp = warp_id(), lane = lane_id();
  const uint32_t sbase = smem_u32(smem);
  const int role = p.cta_role[cta];
  const int ridx = p.cta_index[cta];
  const int rcnt = p.cta_count[cta];
  float* const out_w = p.partials + size_t(cta) * WG_PARTIAL_FLOATS;
  float* const out_b = out_w + 65536;
  const RoleInfo R = role_info(role < 0 ? 0 : role, p.NH);
  const uint32_t b_half_bytes = uint32_t(R.b_chunks) * WG_SUB * 128;

  // work list: tiles t = ridx + i*rcnt
  const long long total_tiles = p.seg_tiles;
  const long long n_items = role < 0 ? 0 : ((total_tiles > ridx) ? (total_tiles - ridx + rcnt - 1) / rcnt : 0);

  auto get_item = [&](long long i) -> WgItem {
    WgItem it;
    const long long lt = ridx + i * rcnt;
    const WgradSegment& sg = p.seg;
    it.a_ptr = (R.a_kind == 0 ? sg.dz : sg.h) + (size_t(lt) * NUM_TRUNK + R.a_layer) * A_TILE_BYTES;
    if (R.b_kind == 0) it.b_ptr = sg.h + (size_t(lt) * NUM_TRUNK + R.b_layer) * A_TILE_BYTES;
    else if (R.b_kind == 1) it.b_ptr = sg.e + size_t(lt) * E_TILE_BYTES;
    else it.b_ptr = sg.d_o + size_t(lt) * (2 * A_CHUNK_BYTES);
    return it;
  };

  if (threadIdx.x == 0) {
    for (int i = 0; i < WG_STAGES; ++i) {
      mbar_init(smem_u32(&bars.full[i]), 1);
      mbar_init(smem_u32(&bars.empty[i]), 5);
    }
    mbar_init(smem_u32(&bars.done), 1);
    fence_mbar_init();
  }
  if (warp == 4) tmem_alloc(smem_u32(&tmem_base_s), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  bool any_mma = false;

  if (warp == 4) {
    // ================================ loader ====================================
    // whole-warp control flow, one elected lane issues (see mlp_fwd.cu)
    uint32_t st = 0, phase = 0;
    for (long long i = 0; i < n_items; ++i) {
      const WgItem it = get_item(i);
      for (int sub = 0; sub < 2; ++sub) {
        mbar_wait(smem_u32(&bars.empty[st]), phase ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(smem_u32(&bars.full[st]), WG_HALF + b_half_bytes);
          const uint32_t dst = sbase + st * WG_STAGE_BYTES;
          // SW128 images: 64 rows of each 64-column chunk; T images: two whole 32-row groups (contiguous)
          if (R.a_t) {
            bulk_g2s(dst, it.a_ptr + size_t(sub) * WG_HALF, WG_HALF, smem_u32(&bars.full[st]));
          } else {
#pragma unroll
            for (int c = 0; c < 4; ++c)
              bulk_g2s(dst + c * (WG_SUB * 128), it.a_ptr + size_t(c) * A_CHUNK_BYTES + sub * (WG_SUB * 128),
                       WG_SUB * 128, smem_u32(&bars.full[st]));
          }
          if (R.b_t) {
            bulk_g2s(dst + WG_HALF, it.b_ptr + size_t(sub) * WG_HALF, WG_HALF, smem_u32(&bars.full[st]));
          } else {
            for (int c = 0; c < R.b_chunks; ++c)
              bulk_g2s(dst + WG_HALF + c * (WG_SUB * 128),
                       it.b_ptr + size_t(c) * A_CHUNK_BYTES + sub * (WG_SUB * 128), WG_SUB * 128,
                       smem_u32(&bars.full[st]));
          }
        }
        __syncwarp();
        if (++st == WG_STAGES) {
          st = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 5) {
    // ================================= MMA ======================================
    uint32_t st = 0, phase = 0;
    const uint32_t idesc = make_idesc_f16(128, R.N, 1, 1);
    // MN-major SW128: LBO = stride between 64-feature chunks (8 KB here), SBO = 8-sample group
    constexpr uint64_t DESC_HI = make_sdesc_hi(1024, LAYOUT_SW128) | (uint64_t((WG_SUB * 128) >> 4) << 16);
    // MN-major, no swizzle (T images): LBO = next 8 samples = 128 B, SBO = next 8 features = 512 B
    constexpr uint64_t T_HI = make_sdesc_hi(512, LAYOUT_NONE) | (uint64_t(128 >> 4) << 16);
    bool first = true;
    for (long long i = 0; i < n_items; ++i) {
      for (int sub = 0; sub < 2; ++sub) {
        mbar_wait(smem_u32(&bars.full[st]), phase);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t a0 = sbase + st * WG_STAGE_BYTES;
          const uint64_t ad0 = (R.a_t ? T_HI : DESC_HI) | uint64_t((a0 >> 4) & 0x3FFF);
          const uint64_t bd0 = (R.b_t ? T_HI : DESC_HI) | uint64_t(((a0 + WG_HALF) >> 4) & 0x3FFF);
#pragma unroll
          for (int ks = 0; ks < WG_SUB / 16; ++ks) {
            const uint32_t acc = !(first && sub == 0 && ks == 0);
            // SW128: 16 samples = 2048 bytes = +128 encoded; features 128..255 = +2 chunks = +1024 encoded
            // T    : 32-sample group = 16 KB = +1024 encoded, 16 samples inside it = 256 B = +16 encoded;
            //        features 128..255 = 16 units x 512 B = +512 encoded
            const uint32_t sw_k = uint32_t(ks) * 128u, t_k = uint32_t(ks >> 1) * 1024u + uint32_t(ks & 1) * 16u;
            const uint64_t ad = ad0 + (R.a_t ? t_k : sw_k), bd = bd0 + (R.b_t ? t_k : sw_k);
            umma_f16(tmem, ad, bd, idesc, acc);
            umma_f16(tmem + 256, ad + (R.a_t ? 512u : 1024u), bd, idesc, acc);
          }
          umma_commit(smem_u32(&bars.empty[st]));
        }
        __syncwarp();
        if (++st == WG_STAGES) {
          st = 0;
          phase ^= 1;
        }
      }
      first = false;
    }
    if (elect_one()) umma_commit(smem_u32(&bars.done));
    __syncwarp();
  } else if (warp < 4) {
    // ========================= bias column sums (warps 0-3) ======================
    const int fp = threadIdx.x;            // feature pair 0..127 -> features 2fp, 2fp+1
    float s0 = 0.f, s1 = 0.f;
    uint32_t st = 0, phase = 0;
    const int nfeat = R.bias_from_b ? R.N : 256;
    const bool active = R.has_bias && (2 * fp < nfeat);
    const uint32_t src_off = (R.bias_from_b ? WG_HALF : 0) + uint32_t(fp >> 5) * (WG_SUB * 128);
    const uint32_t unit = uint32_t(fp & 31) >> 2, wsel = uint32_t(fp & 3) * 4;
    for (long long i = 0; i < n_items; ++i) {
      any_mma = true;
      for (int sub = 0; sub < 2; ++sub) {
        mbar_wait(smem_u32(&bars.full[st]), phase);
        if (active) {
          const uint8_t* base = smem + st * WG_STAGE_BYTES + src_off;
#pragma unroll 8
          for (int r = 0; r < WG_SUB; ++r) {
            const uint32_t w =
                *reinterpret_cast<const uint32_t*>(base + r * 128 + ((unit ^ uint32_t(r & 7)) << 4) + wsel);
            const float2 f = unpack_f16x2(w);
            s0 += f.x;
            s1 += f.y;
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&bars.empty[st]));
        if (++st == WG_STAGES) {
          st = 0;
          phase ^= 1;
        }
      }
    }
    if (role >= 0) {
      out_b[2 * fp] = active ? s0 : 0.f;
      out_b[2 * fp + 1] = active ? s1 : 0.f;
    }
    // ============================ drain accumulators =============================
    mbar_wait(smem_u32(&bars.done), 0);
    tc_fence_after();
    const int m = threadIdx.x;  // TMEM lane
    if (role >= 0) {
#pragma unroll 1
      for (int half = 0; half < 2; ++half) {
        float* dst = out_w + size_t(half * 128 + m) * R.N;
        for (int c0 = 0; c0 < R.N; c0 += 16) {
          uint32_t v[16];
          if (any_mma) {
            tmem_ld16(tmem + (uint32_t(warp * 32) << 16) + half * 256 + c0, v);
            tmem_ld_wait();
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = 0u;
          }
#pragma unroll
          for (int j = 0; j < 16; j += 4)
            *reinterpret_cast<uint4*>(dst + c0 + j) = make_uint4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem, 512);
}

}  // namespace pob
