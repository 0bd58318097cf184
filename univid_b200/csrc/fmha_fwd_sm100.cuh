// Flash-attention forward for Wan DiT tokens on sm_100a: head_dim 128, bf16 in / bf16 out,
// non-causal, optional per-batch key length mask (reference: flash_attention() k_lens,
// models/wan/utils/modules/attention.py:72-80) and optional per-key modifiers used by the
// text-weighted cross-attention (reference: Wan22ContextWrapper hook, models/model_pipeline.py:1756-1803).
//
// Structure (one CTA per 256 query rows of one (batch, head); 3 warpgroups):
//   warps 0-3   softmax group 0  (query rows   0..127, one row per thread)   setmaxnreg.inc
//   warps 4-7   softmax group 1  (query rows 128..255)                       setmaxnreg.inc
//   warp  8     tcgen05.mma issuer (one elected thread) + TMEM allocator     setmaxnreg.dec
//   warp  9     TMA producer (one elected thread)                            setmaxnreg.dec
//   warps 10-11 idle (they only exist so the third warpgroup can give its registers away)
// TMEM (512 columns x 128 lanes):  S0 | S1 | O0 | O1, 128 fp32 columns each.  P (bf16) aliases the
// first 64 columns of its S tile and is consumed as the A operand of the PV MMA straight from TMEM.
// K/V tiles (128 keys x 128 dims) alternate through one TMA ring; Q stays resident in smem and its
// buffer is reused to stage O for the TMA store.
//
// Per KV tile j and query tile t the dependency chain is
//   S_t = Q_t K_j^T  (SS MMA)  -> s_full[t] -> softmax (row max, lazy rescale of O_t, exp2, row sum,
//   P_t -> TMEM) -> p_full[t] -> O_t += P_t V_j (TS MMA) ; S_t = Q_t K_{j+1}^T ...
// and the two query tiles ping-pong so the tensor pipe works on one while the other is in softmax.
// tcgen05 MMAs issued by one thread execute in order, so S_t(j+1) overwriting the P_t(j) columns is
// safe, and s_full[t] (a tcgen05.commit) also implies PV_t(j-1) has completed, which is what allows
// the softmax warps to rescale O_t in place without a dedicated correction group.
#pragma once
#include "ptx.cuh"

namespace uvb {

constexpr int kBlockM = 128;          // query rows per tile (UMMA M)
constexpr int kBlockN = 128;          // keys per KV tile  (UMMA N of QK^T, K of PV)
constexpr int kHeadDim = 128;
constexpr int kQTiles = 2;            // query tiles per CTA
constexpr int kTileBytes = kBlockN * kHeadDim * 2;   // 32 KiB, one [128 x 128] bf16 tile
constexpr int kHalfTile = kTileBytes / 2;            // one 64-column swizzle panel
constexpr int kFmhaThreads = 384;
constexpr int kRegsSoftmax = 224;                    // 2*128*224 + 128*56 = 64512 <= 65536
constexpr int kRegsOther = 56;
constexpr float kRescaleThreshold = 8.0f;            // lazy rescale: only when max grows by > 2^8

template <int kStages>
struct FmhaSmem {
  static constexpr int kQOff = 0;
  static constexpr int kKvOff = kQTiles * kTileBytes;
  static constexpr int kBarOff = kKvOff + kStages * kTileBytes;
  // barriers: q_full[2] kv_full[S] kv_empty[S] s_full[2] p_full[2] o_full[2] + tmem ptr
  static constexpr int kNumBars = 2 + 2 * kStages + 6;
  static constexpr int kBytes = kBarOff + kNumBars * 8 + 16;
  static constexpr int kDynBytes = kBytes + 1024;  // slack for 1024 B alignment
};

struct FmhaParams {
  CUtensorMap tm_q;   // dims (d, token, head, batch), box (64, 128, 1, 1), SWIZZLE_128B
  CUtensorMap tm_k;
  CUtensorMap tm_v;
  CUtensorMap tm_o;
  const int* k_lens;              // [B] or nullptr (= Lk)
  const float* key_logit_scale;   // [Lk] or nullptr : logits[:, j] *= key_logit_scale[j]
  const float* key_pv_weight;     // [Lk] or nullptr : P[:, j] *= w[j] after the row sum
  const float* out_bias;          // [N*128] or nullptr : added to the normalised output
  int Lq;
  int Lk;
  float scale_log2;               // softmax_scale * log2(e)
};

template <int kStages, bool kKeyMod>
__global__ void __launch_bounds__(kFmhaThreads, 1)
fmha_fwd_kernel(const __grid_constant__ FmhaParams p) {
  using SM = FmhaSmem<kStages>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* smem_q = smem + SM::kQOff;
  uint8_t* smem_kv = smem + SM::kKvOff;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SM::kBarOff);
  uint64_t* q_full = bars;                       // [2]
  uint64_t* kv_full = bars + 2;                  // [kStages]
  uint64_t* kv_empty = kv_full + kStages;        // [kStages]
  uint64_t* s_full = kv_empty + kStages;         // [2]
  uint64_t* p_full = s_full + 2;                 // [2]
  uint64_t* o_full = p_full + 2;                 // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(o_full + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q_row0 = blockIdx.x * (kQTiles * kBlockM);
  const int head = blockIdx.y;
  const int batch = blockIdx.z;

  int k_len = p.Lk;
  if (p.k_lens != nullptr) k_len = min(max(p.k_lens[batch], 0), p.Lk);
  const int n_kv = (k_len + kBlockN - 1) / kBlockN;

  if (threadIdx.x == 0) {
    mbar_init(&q_full[0], 1);
    mbar_init(&q_full[1], 1);
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    for (int t = 0; t < 2; ++t) {
      mbar_init(&s_full[t], 1);
      mbar_init(&p_full[t], kBlockM);
      mbar_init(&o_full[t], 1);
    }
    fence_mbar_init();
  }
  if (warp == 8) {
    tmem_alloc(tmem_ptr, 512);
    tmem_relinquish();
  }
  if (warp == 9 && lane == 0) {
    tma_prefetch_desc(&p.tm_q);
    tma_prefetch_desc(&p.tm_k);
    tma_prefetch_desc(&p.tm_v);
    tma_prefetch_desc(&p.tm_o);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp >= 8) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegsOther));
    if (n_kv > 0 && warp == 9) {
      // ===================================== TMA producer =====================================
      if (lane == 0) {
        for (int t = 0; t < kQTiles; ++t) {
          mbar_arrive_expect_tx(&q_full[t], kTileBytes);
          uint8_t* dst = smem_q + t * kTileBytes;
          tma_load_4d_hint(dst, &p.tm_q, &q_full[t], 0, q_row0 + t * kBlockM, head, batch,
                           kEvictFirst);
          tma_load_4d_hint(dst + kHalfTile, &p.tm_q, &q_full[t], 64, q_row0 + t * kBlockM, head,
                           batch, kEvictFirst);
        }
        // ring order matches consumption order: K0, V0, K1, V1, ...
        const int total = 2 * n_kv;
        for (int i = 0; i < total; ++i) {
          const int stage = i % kStages;
          const uint32_t phase = (i / kStages) & 1;
          mbar_wait(&kv_empty[stage], phase ^ 1);
          mbar_arrive_expect_tx(&kv_full[stage], kTileBytes);
          uint8_t* dst = smem_kv + stage * kTileBytes;
          const CUtensorMap* tm = (i & 1) ? &p.tm_v : &p.tm_k;
          const int key0 = (i >> 1) * kBlockN;
          tma_load_4d_hint(dst, tm, &kv_full[stage], 0, key0, head, batch, kEvictLast);
          tma_load_4d_hint(dst + kHalfTile, tm, &kv_full[stage], 64, key0, head, batch, kEvictLast);
        }
      }
    } else if (n_kv > 0 && warp == 8) {
      // ===================================== MMA issuer =======================================
      if (lane == 0) {
        constexpr uint32_t idesc_qk = umma_idesc_bf16(kBlockM, kBlockN, 0, 0);
        constexpr uint32_t idesc_pv = umma_idesc_bf16(kBlockM, kHeadDim, 0, 1);
        const uint32_t q_addr = smem_u32(smem_q);
        const uint32_t kv_addr = smem_u32(smem_kv);

        // S_t = Q_t K^T : A, B K-major; 8 K-steps of 16; panel = kk/4, 32 B per step in the atom
        auto issue_qk = [&](int t, uint32_t k_stage_addr) {
          const uint32_t qa = q_addr + t * kTileBytes;
  #pragma unroll
          for (int kk = 0; kk < kHeadDim / 16; ++kk) {
            const uint32_t off = (kk >> 2) * kHalfTile + (kk & 3) * 32;
            umma_ss(tmem_base + t * kBlockN, umma_desc_sw128(qa + off, 16, 1024),
                    umma_desc_sw128(k_stage_addr + off, 16, 1024), idesc_qk, kk > 0 ? 1u : 0u);
          }
        };
        // O_t (+)= P_t V : A = P from TMEM (8 columns per 16 keys), B = V, MN-major:
        // 16 keys = 2 KiB per step, the two 64-dim panels are kHalfTile apart (LBO), 8-key groups 1 KiB (SBO)
        auto issue_pv = [&](int t, uint32_t v_stage_addr, bool accumulate) {
  #pragma unroll
          for (int kk = 0; kk < kBlockN / 16; ++kk) {
            umma_ts(tmem_base + 2 * kBlockN + t * kHeadDim, tmem_base + t * kBlockN + kk * 8,
                    umma_desc_sw128(v_stage_addr + kk * 2048, kHalfTile, 1024), idesc_pv,
                    (accumulate || kk > 0) ? 1u : 0u);
          }
        };

        int it = 0;  // ring index
        mbar_wait(&q_full[0], 0);
        mbar_wait(&kv_full[0], 0);
        tc_fence_after();
        issue_qk(0, kv_addr);
        tc_commit(&s_full[0]);
        mbar_wait(&q_full[1], 0);
        tc_fence_after();
        issue_qk(1, kv_addr);
        tc_commit(&s_full[1]);
        tc_commit(&kv_empty[0]);
        it = 1;
        for (int j = 0; j < n_kv; ++j) {
          const bool has_next = (j + 1) < n_kv;
          const int sv = it % kStages;
          const uint32_t pv_phase = (it / kStages) & 1;
          const int sk = (it + 1) % kStages;
          const uint32_t pk_phase = ((it + 1) / kStages) & 1;
          const uint32_t v_addr = kv_addr + sv * kTileBytes;
          const uint32_t k_addr = kv_addr + sk * kTileBytes;
          mbar_wait(&kv_full[sv], pv_phase);
  #pragma unroll
          for (int t = 0; t < kQTiles; ++t) {
            mbar_wait(&p_full[t], j & 1);
            tc_fence_after();
            issue_pv(t, v_addr, j > 0);
            if (has_next) {
              if (t == 0) {
                mbar_wait(&kv_full[sk], pk_phase);
                tc_fence_after();
              }
              issue_qk(t, k_addr);
              tc_commit(&s_full[t]);
            } else {
              tc_commit(&o_full[t]);
            }
          }
          tc_commit(&kv_empty[sv]);
          if (has_next) tc_commit(&kv_empty[sk]);
          it += 2;
        }
      }
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRegsSoftmax));
    if (n_kv == 0) {
      // No valid key: the output rows are defined as zero (flash-attn varlen convention).
      const int t = warp >> 2;
      const int row = (warp & 3) * 32 + lane;
      uint8_t* so = smem_q + t * kTileBytes;
#pragma unroll
      for (int c = 0; c < 16; ++c) {
        *reinterpret_cast<uint4*>(so + (c >> 3) * kHalfTile + row * 128 + ((c & 7) << 4)) =
            make_uint4(0, 0, 0, 0);
      }
      fence_proxy_async_smem();
      named_bar_sync(1 + t, kBlockM);
      if ((warp & 3) == 0 && lane == 0) {
        tma_store_4d(&p.tm_o, so, 0, q_row0 + t * kBlockM, head, batch);
        tma_store_4d(&p.tm_o, so + kHalfTile, 64, q_row0 + t * kBlockM, head, batch);
        tma_store_commit();
        tma_store_wait0();
      }
    } else {
      // ===================================== softmax groups ====================================
      const int t = warp >> 2;
      const int wq = warp & 3;
      const int row = wq * 32 + lane;
      const uint32_t lane_addr = static_cast<uint32_t>(wq * 32) << 16;
      const uint32_t tS = tmem_base + lane_addr + t * kBlockN;
      const uint32_t tO = tmem_base + lane_addr + 2 * kBlockN + t * kHeadDim;
      const float scale_log2 = p.scale_log2;
      float m = -INFINITY;  // running row max in raw-logit units
      float l = 0.f;        // running row sum of exp2((s - m) * scale_log2)

      for (int j = 0; j < n_kv; ++j) {
        mbar_wait(&s_full[t], j & 1);
        tc_fence_after();
        uint32_t sr[kBlockN];
        tmem_ld_x32(tS + 0, sr + 0);
        tmem_ld_x32(tS + 32, sr + 32);
        tmem_ld_x32(tS + 64, sr + 64);
        tmem_ld_x32(tS + 96, sr + 96);
        tmem_wait_ld();

        if constexpr (kKeyMod) {
          if (p.key_logit_scale != nullptr) {
            const float4* ks = reinterpret_cast<const float4*>(p.key_logit_scale + j * kBlockN);
  #pragma unroll
            for (int c = 0; c < kBlockN / 4; ++c) {
              // tail reads past Lk are masked below; the host pads the array to a tile multiple
              const float4 w = __ldg(ks + c);
              sr[4 * c + 0] = __float_as_uint(__uint_as_float(sr[4 * c + 0]) * w.x);
              sr[4 * c + 1] = __float_as_uint(__uint_as_float(sr[4 * c + 1]) * w.y);
              sr[4 * c + 2] = __float_as_uint(__uint_as_float(sr[4 * c + 2]) * w.z);
              sr[4 * c + 3] = __float_as_uint(__uint_as_float(sr[4 * c + 3]) * w.w);
            }
          }
        }
        const int valid = k_len - j * kBlockN;
        if (valid < kBlockN) {
  #pragma unroll
          for (int c = 0; c < kBlockN; ++c) {
            if (c >= valid) sr[c] = 0xff800000u;  // -inf
          }
        }

        float mx0 = __uint_as_float(sr[0]), mx1 = __uint_as_float(sr[1]);
        float mx2 = __uint_as_float(sr[2]), mx3 = __uint_as_float(sr[3]);
  #pragma unroll
        for (int c = 4; c < kBlockN; c += 4) {
          mx0 = fmaxf(mx0, __uint_as_float(sr[c + 0]));
          mx1 = fmaxf(mx1, __uint_as_float(sr[c + 1]));
          mx2 = fmaxf(mx2, __uint_as_float(sr[c + 2]));
          mx3 = fmaxf(mx3, __uint_as_float(sr[c + 3]));
        }
        const float tile_max = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));

        if (j == 0) {
          m = tile_max;
        } else {
          const bool need = (tile_max - m) * scale_log2 > kRescaleThreshold;
          if (__any_sync(0xffffffffu, need)) {
            // s_full[t] of this iteration implies PV_t(j-1) is complete and PV_t(j) is not issued
            // until we arrive on p_full[t]: O_t is ours to rescale.
            const float m_new = fmaxf(m, tile_max);
            const float f = ex2_approx((m - m_new) * scale_log2);
            l *= f;
            m = m_new;
  #pragma unroll
            for (int c = 0; c < kHeadDim / 32; ++c) {
              uint32_t orr[32];
              tmem_ld_x32(tO + c * 32, orr);
              tmem_wait_ld();
  #pragma unroll
              for (int i = 0; i < 32; ++i) orr[i] = __float_as_uint(__uint_as_float(orr[i]) * f);
              tmem_st_x32(tO + c * 32, orr);
            }
            tmem_wait_st();
          }
        }

        const float neg_ms = -m * scale_log2;
        float sum0 = 0.f, sum1 = 0.f, sum2 = 0.f, sum3 = 0.f;
        uint32_t pk[kBlockN / 2];
        const float* pvw = nullptr;
        if constexpr (kKeyMod) {
          if (p.key_pv_weight != nullptr) pvw = p.key_pv_weight + j * kBlockN;
        }
  #pragma unroll
        for (int c = 0; c < kBlockN; c += 4) {
          float e0 = ex2_approx(fmaf(__uint_as_float(sr[c + 0]), scale_log2, neg_ms));
          float e1 = ex2_approx(fmaf(__uint_as_float(sr[c + 1]), scale_log2, neg_ms));
          float e2 = ex2_approx(fmaf(__uint_as_float(sr[c + 2]), scale_log2, neg_ms));
          float e3 = ex2_approx(fmaf(__uint_as_float(sr[c + 3]), scale_log2, neg_ms));
          sum0 += e0;
          sum1 += e1;
          sum2 += e2;
          sum3 += e3;
          if constexpr (kKeyMod) {
            if (pvw != nullptr) {
              const float4 w = __ldg(reinterpret_cast<const float4*>(pvw + c));
              e0 *= w.x;
              e1 *= w.y;
              e2 *= w.z;
              e3 *= w.w;
            }
          }
          pk[c / 2 + 0] = pack_bf16x2(e0, e1);
          pk[c / 2 + 1] = pack_bf16x2(e2, e3);
        }
        l += (sum0 + sum1) + (sum2 + sum3);

        tmem_st_x32(tS + 0, pk + 0);
        tmem_st_x32(tS + 32, pk + 32);
        tmem_wait_st();
        tc_fence_before();
        mbar_arrive(&p_full[t]);
      }

      // ------------------------------- epilogue: O_t / l -> bf16 -> smem -> TMA store -------------
      mbar_wait(&o_full[t], 0);
      tc_fence_after();
      const float inv = 1.0f / l;
      uint8_t* so = smem_q + t * kTileBytes;  // Q_t is dead once o_full[t] has fired
      const float* bias = nullptr;
      if constexpr (kKeyMod) {
        if (p.out_bias != nullptr) bias = p.out_bias + head * kHeadDim;
      }
  #pragma unroll
      for (int c = 0; c < kHeadDim / 32; ++c) {
        uint32_t orr[32];
        tmem_ld_x32(tO + c * 32, orr);
        tmem_wait_ld();
        uint32_t ob[16];
  #pragma unroll
        for (int i = 0; i < 16; ++i) {
          float a = __uint_as_float(orr[2 * i]) * inv;
          float b = __uint_as_float(orr[2 * i + 1]) * inv;
          if constexpr (kKeyMod) {
            if (bias != nullptr) {
              a += __ldg(bias + c * 32 + 2 * i);
              b += __ldg(bias + c * 32 + 2 * i + 1);
            }
          }
          ob[i] = pack_bf16x2(a, b);
        }
        // 32 columns = 4 x 16-byte chunks of panel (c/2); SWIZZLE_128B: chunk ^= row % 8
        uint8_t* prow = so + (c >> 1) * kHalfTile + row * 128;
  #pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          const int chunk = (c & 1) * 4 + q4;
          *reinterpret_cast<uint4*>(prow + ((chunk ^ (row & 7)) << 4)) =
              make_uint4(ob[4 * q4], ob[4 * q4 + 1], ob[4 * q4 + 2], ob[4 * q4 + 3]);
        }
      }
      fence_proxy_async_smem();
      named_bar_sync(1 + t, kBlockM);
      if (wq == 0 && lane == 0) {
        tma_store_4d(&p.tm_o, so, 0, q_row0 + t * kBlockM, head, batch);
        tma_store_4d(&p.tm_o, so + kHalfTile, 64, q_row0 + t * kBlockM, head, batch);
        tma_store_commit();
        tma_store_wait0();
      }
    }
  }

  // ------------------------------------------ teardown ------------------------------------------
  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    __syncwarp();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace uvb
