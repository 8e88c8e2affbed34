// bwd_body.cuh — device body of the dgrad chain, shared by mlp_bwd.cu (classic) and mlp_bwdw.cu (fused).
#pragma once
// mlp_bwd.cu — data-gradient chain of the NeRF-SH MLP (the dgrad half of jax.value_and_grad in
// nerf_sh/train.py:116), fused per 256-sample iteration like mlp_fwd:
//
//   G' (per-sample d pre_rgb[3], d sigma_raw from render.cu)  --SH basis-->  dO [128 x NH]
//   dH_7 = dO . W_heads ;  dZ_l = dH_l * relu'(h_l) ;  dH_{l-1} = dZ_l . W_l   (l = 7..1)
//
// ReLU masks come from the forward pass (1 bit per activation), the transposed weights from the
// packed `wt_hi` images.  Every dZ_l tile (and dO) is stored to global memory in the same
// swizzled tile-image format as the forward activations; mlp_wgrad contracts them over samples.
// No gradient w.r.t. the inputs is needed (layer 0 and the skip slice of layer 5 stop here).
#include "common.cuh"
#include "kernels.h"

namespace pob {

namespace {

constexpr int BWD_THREADS = 320;
constexpr int BWD_PRODUCER_WARP = 8;
constexpr int BWD_MMA_WARP = 9;
constexpr int BWD_WSLOTS = 6;

constexpr uint32_t SB_A0 = 0;
constexpr uint32_t SB_A1 = SB_A0 + A_TILE_BYTES;
constexpr uint32_t SB_W = SB_A1 + A_TILE_BYTES;
constexpr uint32_t SB_TOTAL = SB_W + BWD_WSLOTS * WSLOT_BYTES;  // 128K + 96K = 224K

struct BwdBarriers {
  uint64_t full[BWD_WSLOTS];
  uint64_t empty[BWD_WSLOTS];
  uint64_t a_ready[2];
  uint64_t d_ready[2];
};

}  // namespace

// cta / ncta: index of this CTA among the dgrad CTAs of the launch and their number
__device__ __forceinline__ void bwd_body(const BwdParams& p, uint8_t* smem, const int cta, const int ncta) {
  __shared__ __align__(8) BwdBarriers bars;
  __shared__ uint32_t tmem_base_s;

  const long long num_iters = (p.M + 2 * TILE_M - 1) / (2 * TILE_M);
  const long long mrows = num_iters * 2 * TILE_M;
  const uint32_t warp = warp_id(), lane = lane_id();
  const uint32_t sbase = smem_u32(smem);
  const int NH = p.NH;
  const int hs = (NH + 31) / 32;            // K slots of the heads dgrad
  const int do_chunks = (NH + 63) / 64;     // 64-wide chunks of the dO tile image

  if (threadIdx.x == 0) {
    for (int i = 0; i < BWD_WSLOTS; ++i) {
      mbar_init(smem_u32(&bars.full[i]), 1);
      mbar_init(smem_u32(&bars.empty[i]), 1);
    }
    for (int g = 0; g < 2; ++g) {
      mbar_init(smem_u32(&bars.a_ready[g]), 4);
      mbar_init(smem_u32(&bars.d_ready[g]), 1);
    }
    fence_mbar_init();
  }
  if (warp == BWD_PRODUCER_WARP) tmem_alloc(smem_u32(&tmem_base_s), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;

  if (warp == BWD_PRODUCER_WARP) {
    // whole-warp control flow, one elected lane issues (see mlp_fwd.cu)
    uint32_t slot = 0, phase = 0;
    const int nslots = hs + 7 * 8;
    for (long long it = cta; it < num_iters; it += ncta) {
      for (int j = 0; j < nslots; ++j) {
        mbar_wait(smem_u32(&bars.empty[slot]), phase ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(smem_u32(&bars.full[slot]), WSLOT_BYTES);
          bulk_g2s(sbase + SB_W + slot * WSLOT_BYTES, p.w.wt_hi + size_t(j) * WSLOT_BYTES, WSLOT_BYTES,
                   smem_u32(&bars.full[slot]));
        }
        __syncwarp();
        if (++slot == BWD_WSLOTS) {
          slot = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == BWD_MMA_WARP) {
    uint32_t slot = 0, phase = 0, aphase = 0;
    const uint32_t idesc = make_idesc_f16(TILE_M, WIDTH);
    constexpr uint64_t A_HI = make_sdesc_hi(1024, LAYOUT_SW128) | (uint64_t(1) << 16);
    constexpr uint64_t W_HI = make_sdesc_hi(512, LAYOUT_SW64) | (uint64_t(1) << 16);
    for (long long it = cta; it < num_iters; it += ncta) {
      for (int grp = 0; grp < 8; ++grp) {      // heads, then Dense_7 .. Dense_1
        const int ns = (grp == 0) ? hs : 8;
        for (int j = 0; j < ns; ++j) {
          const uint32_t a_off = uint32_t(j >> 1) * A_CHUNK_BYTES + uint32_t(j & 1) * 64u;
          mbar_wait(smem_u32(&bars.full[slot]), phase);
          const uint64_t bd0 = W_HI | uint64_t(((sbase + SB_W + slot * WSLOT_BYTES) >> 4) & 0x3FFF);
#pragma unroll
          for (int g = 0; g < 2; ++g) {
            if (j == 0) mbar_wait(smem_u32(&bars.a_ready[g]), aphase);
            tc_fence_after();
            if (elect_one()) {
              const uint32_t a_base = sbase + (g ? SB_A1 : SB_A0) + a_off;
              const uint64_t ad0 = A_HI | uint64_t((a_base >> 4) & 0x3FFF);
              const uint32_t d = tmem + uint32_t(g) * 256u;
              umma_f16(d, ad0, bd0, idesc, j != 0);
              umma_f16(d, ad0 + 2, bd0 + 2, idesc, 1u);
              if (j == ns - 1) umma_commit(smem_u32(&bars.d_ready[g]));
              if (g == 1) umma_commit(smem_u32(&bars.empty[slot]));
            }
            __syncwarp();
          }
          if (++slot == BWD_WSLOTS) {
            slot = 0;
            phase ^= 1;
          }
        }
        aphase ^= 1;
      }
    }
  } else {
    const int g = warp >> 2;
    const int row = int((warp & 3) * 32 + lane);
    uint8_t* const a_tile = smem + (g ? SB_A1 : SB_A0);
    const uint32_t d_tmem = tmem + (uint32_t((warp & 3) * 32) << 16) + uint32_t(g) * 256u;
    uint32_t dphase = 0;
    const bool fused = p.q.slots != nullptr;
    const bool flagger = (threadIdx.x & 127) == 0;
    uint32_t kiter = 0;   // iterations completed by this CTA (queue sequence number)

    // destination of queue `qi` for tile g (fused) / wait until its previous content has been consumed
    auto q_slot = [&](int qi) -> uint8_t* {
      return p.q.slots + ((size_t(cta) * BWDW_QUEUES + qi) * 2 + g) * A_TILE_BYTES;
    };
    long long stall_cycles = 0;
    const long long t_begin = clock64();
    auto q_wait_free = [&](int qi) {
      if (fused && flagger) {
        const long long t0 = clock64();
        const uint32_t* c = p.q.consumed + (size_t(cta) * BWDW_QUEUES + qi) * 2;
        while (ld_acquire_gpu(c) < kiter) {
        }
        if (qi == 1 + (7 - SKIP_LAYER))   // dZ_5 has a second reader (the skip rows of Dense_5)
          while (ld_acquire_gpu(c + 1) < kiter) {
          }
        stall_cycles += clock64() - t0;
      }
    };
    // after a tile copy: every thread of the group is done reading a_tile (the next epilogue may overwrite
    // it) and, on the fused path, the tile is published to its consumer
    auto q_publish = [&](int qi) {
      if (fused) __threadfence();
      named_bar_sync(1 + g, 128);
      if (fused && flagger) st_release_gpu(p.q.produced + (size_t(cta) * BWDW_QUEUES + qi) * 2 + g, kiter + 1);
    };

    for (long long it = cta; it < num_iters; it += ncta, ++kiter) {
      const long long tile_idx = it * 2 + g;
      const long long s = tile_idx * TILE_M + row;
      // ---- dO row from the per-sample gradient and the SH basis ----
      {
        float4 gq = make_float4(0.f, 0.f, 0.f, 0.f);
        float basis[25];
#pragma unroll
        for (int k = 0; k < 25; ++k) basis[k] = 0.f;   // padded rows: 0 * garbage must not become NaN
        basis[0] = 1.f;
        if (s < p.M) {
          gq = p.G[s];
          const long long vi = p.n_per_ray > 0 ? s / p.n_per_ray : s;
          const float* vd = p.viewdirs + 3 * vi;
          if (p.sh_deg >= 0) sh_basis(p.sh_deg, __ldg(vd), __ldg(vd + 1), __ldg(vd + 2), basis);
        }
        const float gc[3] = {gq.x, gq.y, gq.z};
        // dO / dZ tiles go to global memory straight from the registers (a bulk store out of shared
        // memory competes with the MMA operand reads of the next GEMM)
        uint8_t* const do_glob = fused ? q_slot(0) : p.save_do + size_t(tile_idx) * (2 * A_CHUNK_BYTES);
        // every warp of the group must be done copying the previous iteration's dZ_0 image out of a_tile
        named_bar_sync(1 + g, 128);
#pragma unroll
        for (int u = 0; u < 16; ++u) {            // 16-byte units of 8 columns, up to 128 columns
          if (u * 8 < do_chunks * 64) {
            uint32_t w[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              float f[2];
#pragma unroll
              for (int e = 0; e < 2; ++e) {
                const int n = u * 8 + 2 * i + e;
                float v = 0.f;
                if (n == 0) v = gq.w;
                else if (n < 1 + 3 * 25) {
                  const int k = (n - 1) / 3, c = (n - 1) % 3;
                  if (k < p.K) v = gc[c] * basis[k < 25 ? k : 24];
                }
                f[e] = v;
              }
              w[i] = pack_f16x2(f[0], f[1]);
            }
            const uint32_t off = uint32_t(u >> 3) * A_CHUNK_BYTES + uint32_t(row) * 128u +
                                 ((uint32_t(u & 7) ^ uint32_t(row & 7)) << 4);
            *reinterpret_cast<uint4*>(a_tile + off) = make_uint4(w[0], w[1], w[2], w[3]);
          }
        }
        fence_proxy_async_smem();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&bars.a_ready[g]));
        // dO tile -> global / queue (after the hand-over, overlapping the heads dgrad GEMM)
        q_wait_free(0);
        named_bar_sync(1 + g, 128);
        {
          const int t = int(threadIdx.x & 127);
          const uint4* src = reinterpret_cast<const uint4*>(a_tile) + t;
          uint4* dst = reinterpret_cast<uint4*>(do_glob) + t;
          for (int i = 0; i < do_chunks * (A_CHUNK_BYTES / 16 / 128); ++i) dst[i * 128] = src[i * 128];
        }
        q_publish(0);
      }
      // ---- dZ_7 .. dZ_0 ----
      for (int l = NUM_TRUNK - 1; l >= 0; --l) {
        // relu mask of h_l (bit 31-i of word c <-> column 32c+i), prefetched before the wait
        const uint4* mp = reinterpret_cast<const uint4*>(p.mask + (size_t(l) * mrows + s) * 8);
        const uint4 m0 = __ldg(mp), m1 = __ldg(mp + 1);
        const uint32_t mw[8] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w};
        mbar_wait(smem_u32(&bars.d_ready[g]), dphase);
        dphase ^= 1;
        tc_fence_after();
        uint8_t* const dz_glob =
            fused ? q_slot(1 + (7 - l)) : p.save_dz + (size_t(tile_idx) * NUM_TRUNK + l) * A_TILE_BYTES;
        uint32_t va[32], vb[32];
        tmem_ld32(d_tmem, va);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          uint32_t(&v)[32] = (c & 1) ? vb : va;
          tmem_ld_wait();
          if (c + 1 < 8) tmem_ld32(d_tmem + (c + 1) * 32, (c & 1) ? va : vb);   // prefetch next chunk
          const uint32_t m = mw[c];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            uint32_t w[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int e0 = 8 * u + 2 * i;
              const float f0 = (m & (0x80000000u >> e0)) ? __uint_as_float(v[e0]) : 0.f;
              const float f1 = (m & (0x80000000u >> (e0 + 1))) ? __uint_as_float(v[e0 + 1]) : 0.f;
              w[i] = pack_f16x2(f0, f1);
            }
            const uint32_t unit = uint32_t((c & 1) * 4 + u);
            const uint32_t off = uint32_t(c >> 1) * A_CHUNK_BYTES + uint32_t(row) * 128u +
                                 ((unit ^ uint32_t(row & 7)) << 4);
            *reinterpret_cast<uint4*>(a_tile + off) = make_uint4(w[0], w[1], w[2], w[3]);
          }
        }
        fence_proxy_async_smem();
        if (l > 0) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(&bars.a_ready[g]));
        }
        // dZ_l tile -> global / queue, after the hand-over: the copy overlaps the next GEMM (see mlp_fwd.cu)
        q_wait_free(1 + (7 - l));
        named_bar_sync(1 + g, 128);
        {
          const int t = int(threadIdx.x & 127);
          const uint4* src = reinterpret_cast<const uint4*>(a_tile) + t;
          uint4* dst = reinterpret_cast<uint4*>(dz_glob) + t;
#pragma unroll 8
          for (int i = 0; i < A_TILE_BYTES / 16 / 128; ++i) dst[i * 128] = src[i * 128];
        }
        q_publish(1 + (7 - l));
      }
    }
    if (fused && p.q.stall && (threadIdx.x & 127) == 0) {
      p.q.stall[size_t(cta) * 4 + g] = (unsigned long long)stall_cycles;
      p.q.stall[size_t(cta) * 4 + 2 + g] = (unsigned long long)(clock64() - t_begin);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == BWD_PRODUCER_WARP) tmem_dealloc(tmem, 512);
}


}  // namespace pob
