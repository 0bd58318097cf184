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
// TMEM (512 columns x 128 lanes):  S0a S0b | S1a S1b | O0 | O1.  The key axis advances in 64-key
// sub-steps and every query tile owns TWO 64-column S buffers, so the MMA thread issues
// S_t(s+2) = Q_t K(s+2)^T right after O_t += P_t(s) V(s): while the softmax group works on step s the
// scores of step s+1 are already (being) computed, and the softmax<->MMA hand-off is a producer/consumer
// pipeline instead of a serial dependency chain (round-1 ncu: with one S buffer per tile the softmax warps
// spent half their time waiting for S and the tensor pipe was 53 % busy).  P (bf16) aliases the first 32
// columns of its S buffer and is consumed as the A operand of the PV MMA straight from TMEM.
// K/V tiles (128 keys x 128 dims) alternate through one TMA ring; Q stays resident in smem and its
// buffer is reused to stage O for the TMA store.
//
// Ordering facts the protocol relies on: tcgen05 MMAs issued by one thread execute in order, so
// S_t(s+2) overwriting the buffer that held P_t(s) is safe once PV_t(s) has been issued before it; the
// softmax warps rescale O_t in place (lazily, only when a row max grows by more than 2^8) after waiting
// for pv_done[t] of the previous step, and PV_t(s) is not issued before they arrive on p_full[t].
#pragma once
#include "ptx.cuh"

namespace uvb {

constexpr int kBlockM = 128;          // query rows per tile (UMMA M)
constexpr int kBlockN = 128;          // keys per K/V smem tile (TMA granularity)
// kStepN (template): keys per softmax/MMA sub-step (UMMA N of QK^T, K of PV): 128 -> one S buffer per
// query tile, 64 -> two S buffers per tile (the MMA thread runs two steps ahead of the softmax)
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
  // barriers: q_full[2] kv_full[S] kv_empty[S] s_full[2][2] p_full[2][2] pv_done[2] o_full[2] + tmem ptr
  static constexpr int kNumBars = 2 + 2 * kStages + 12;
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

// exp2 of 64 scores of one row -> 32 packed bf16x2 probabilities + partial row sums.  One pair in every
// kPolyEvery goes through the FMA-pipe polynomial instead of MUFU.EX2 (0 = never): the XU pipe does 16
// exp2/clk/SM, exactly the rate at which the tensor pipe consumes a 128x128 tile, so offloading a share
// of them is what lets the softmax keep ahead of the MMAs.
template <int kPolyEvery, bool kKeyMod>
__device__ __forceinline__ void softmax_exp64(const uint32_t* sr, float scale_log2, float neg_ms,
                                              float2& sum_a, float2& sum_b, uint32_t* pk,
                                              const float* pvw) {
  const float2 sc = make_float2(scale_log2, scale_log2);
  const float2 nm = make_float2(neg_ms, neg_ms);
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    const float2 x = ffma2(make_float2(__uint_as_float(sr[2 * i]), __uint_as_float(sr[2 * i + 1])), sc, nm);
    float2 e;
    if (kPolyEvery > 0 && (i % (kPolyEvery > 0 ? kPolyEvery : 1)) == (kPolyEvery - 1)) {
      e = ex2_poly2(x);
    } else {
      e = make_float2(ex2_approx(x.x), ex2_approx(x.y));
    }
    if (i & 1) {
      sum_b = fadd2(sum_b, e);
    } else {
      sum_a = fadd2(sum_a, e);
    }
    if constexpr (kKeyMod) {
      if (pvw != nullptr) {
        const float2 w = __ldg(reinterpret_cast<const float2*>(pvw) + i);
        e.x *= w.x;
        e.y *= w.y;
      }
    }
    pk[i] = pack_bf16x2(e.x, e.y);
  }
}

template <int kStages, int kStepN, int kPolyEvery, bool kKeyMod>
__global__ void __launch_bounds__(kFmhaThreads, 1)
fmha_fwd_kernel(const __grid_constant__ FmhaParams p) {
  static_assert(kStepN == 64 || kStepN == 128, "sub-step must be 64 or 128 keys");
  constexpr int kNB = kBlockN / kStepN;   // S buffers per query tile == sub-steps per K/V tile
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
  uint64_t* s_full = kv_empty + kStages;         // [tile][buffer]
  uint64_t* p_full = s_full + 4;                 // [tile][buffer]
  uint64_t* pv_done = p_full + 4;                // [tile]
  uint64_t* o_full = pv_done + 2;                // [tile]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(o_full + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q_row0 = blockIdx.x * (kQTiles * kBlockM);
  const int head = blockIdx.y;
  const int batch = blockIdx.z;

  int k_len = p.Lk;
  if (p.k_lens != nullptr) k_len = min(max(p.k_lens[batch], 0), p.Lk);
  const int n_kv = (k_len + kBlockN - 1) / kBlockN;     // K/V tiles to load
  const int n_steps = (k_len + kStepN - 1) / kStepN;    // sub-steps

  if (threadIdx.x == 0) {
    mbar_init(&q_full[0], 1);
    mbar_init(&q_full[1], 1);
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    for (int t = 0; t < 2; ++t) {
      for (int b = 0; b < 2; ++b) {
        mbar_init(&s_full[2 * t + b], 1);
        mbar_init(&p_full[2 * t + b], 4);   // one arrive per softmax warp
      }
      mbar_init(&pv_done[t], 1);
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
  // warp-uniform copy (a plain smem load is not provably uniform and would force ptxas to wrap every
  // tcgen05 instruction in an R2UR.BROADCAST waterfall loop)
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr, 0);

  if (warp >= 8) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegsOther));
    if (n_kv > 0 && warp == 9) {
      // ===================================== TMA producer =====================================
      // the whole warp runs the loop (converged, uniform operands); one elected lane issues
      for (int t = 0; t < kQTiles; ++t) {
        uint8_t* dst = smem_q + t * kTileBytes;
        if (elect_one()) {
          mbar_arrive_expect_tx(&q_full[t], kTileBytes);
          tma_load_4d_hint(dst, &p.tm_q, &q_full[t], 0, q_row0 + t * kBlockM, head, batch, kEvictFirst);
          tma_load_4d_hint(dst + kHalfTile, &p.tm_q, &q_full[t], 64, q_row0 + t * kBlockM, head, batch,
                           kEvictFirst);
        }
        __syncwarp();
      }
      // ring order matches consumption order: K0, V0, K1, V1, ...
      const int total = 2 * n_kv;
      for (int i = 0; i < total; ++i) {
        const int stage = i % kStages;
        const uint32_t phase = (i / kStages) & 1;
        mbar_wait(&kv_empty[stage], phase ^ 1);
        uint8_t* dst = smem_kv + stage * kTileBytes;
        const CUtensorMap* tm = (i & 1) ? &p.tm_v : &p.tm_k;
        const int key0 = (i >> 1) * kBlockN;
        if (elect_one()) {
          mbar_arrive_expect_tx(&kv_full[stage], kTileBytes);
          tma_load_4d_hint(dst, tm, &kv_full[stage], 0, key0, head, batch, kEvictLast);
          tma_load_4d_hint(dst + kHalfTile, tm, &kv_full[stage], 64, key0, head, batch, kEvictLast);
        }
        __syncwarp();
      }
    } else if (n_kv > 0 && warp == 8) {
      // ===================================== MMA issuer =======================================
      // The whole warp runs the control flow converged so every operand is warp-uniform; one elected
      // lane (always the same one) issues the tcgen05.mma / tcgen05.commit instructions.  Descriptors are
      // built once; per instruction only a 64-bit add of a compile-time offset remains.  (Round-1 ncu:
      // with the loop inside `if (lane == 0)` ptxas emitted an ELECT/R2UR.BROADCAST waterfall around each
      // UTCHMMA and descriptor math on the uniform datapath -- ~110 cycles per MMA, the real bottleneck.)
      constexpr uint32_t idesc_qk = umma_idesc_bf16(kBlockM, kStepN, 0, 0);
      constexpr uint32_t idesc_pv = umma_idesc_bf16(kBlockM, kHeadDim, 0, 1);
      const uint64_t q_desc = umma_desc_sw128(smem_u32(smem_q), 16, 1024);            // K-major A
      const uint64_t k_desc = umma_desc_sw128(smem_u32(smem_kv), 16, 1024);           // K-major B
      const uint64_t v_desc = umma_desc_sw128(smem_u32(smem_kv), kHalfTile, 1024);    // MN-major B
      constexpr uint64_t kTile16 = kTileBytes >> 4;           // descriptor address units are 16 B
      constexpr uint64_t kStep16 = (kStepN * 128) >> 4;       // kStepN key rows of one 128 B panel

      auto wait_full = [&](int ring) {
        mbar_wait(&kv_full[ring % kStages], (ring / kStages) & 1);
        tc_fence_after();
      };
      auto commit = [&](uint64_t* bar) {
        if (elect_one()) tc_commit(bar);
        __syncwarp();
      };
      // S_t[buf] = Q_t K(step)^T : A, B K-major, N = kStepN keys; 8 K-steps of 16 dims: panel = kk/4,
      // 32 B per K-step inside the 128 B swizzle atom
      auto issue_qk = [&](int t, int step, int buf) {
        const uint64_t qa = q_desc + t * kTile16;
        const uint64_t ka = k_desc + ((2 * (step / kNB)) % kStages) * kTile16 + (step % kNB) * kStep16;
        const uint32_t d = tmem_base + t * 128 + buf * kStepN;
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < kHeadDim / 16; ++kk) {
            const uint64_t off = ((kk >> 2) * kHalfTile + (kk & 3) * 32) >> 4;
            umma_ss(d, qa + off, ka + off, idesc_qk, kk > 0 ? 1u : 0u);
          }
        }
        __syncwarp();
      };
      // O_t (+)= P_t[buf] V(step) : A = P from TMEM (8 columns per 16 keys), B = V rows of the step,
      // MN-major: 16 keys = 2 KiB per K-step, the two 64-dim panels are kHalfTile apart (LBO), 8-key
      // groups 1 KiB apart (SBO)
      auto issue_pv = [&](int t, int step, int buf, int kk0) {   // 4 K-steps = 64 keys from kk0
        const uint64_t va = v_desc + ((2 * (step / kNB) + 1) % kStages) * kTile16 + (step % kNB) * kStep16;
        const uint32_t d = tmem_base + 256 + t * kHeadDim;
        const uint32_t a = tmem_base + t * 128 + buf * kStepN;
        const uint32_t acc0 = (step > 0 || kk0 > 0) ? 1u : 0u;
        if (elect_one()) {
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4) {
            const int kk = kk0 + k4;
            umma_ts(d, a + kk * 8, va + kk * (2048 >> 4), idesc_pv, k4 > 0 ? 1u : acc0);
          }
        }
        __syncwarp();
      };
      // a K/V tile is last read by its last sub-step (or by the very last step)
      auto last_use = [&](int step) { return (step % kNB) == kNB - 1 || step == n_steps - 1; };

      // prologue: scores of the first kNB steps into the S buffers of each tile
      mbar_wait(&q_full[0], 0);
      wait_full(0);
      const int pre = n_steps < kNB ? n_steps : kNB;
      for (int t = 0; t < kQTiles; ++t) {
        if (t == 1) {
          mbar_wait(&q_full[1], 0);
          tc_fence_after();
        }
        for (int s0 = 0; s0 < pre; ++s0) {
          issue_qk(t, s0, s0);
          commit(&s_full[2 * t + s0]);
        }
      }
      commit(&kv_empty[0]);   // K tile 0 is fully consumed by the prologue

      for (int step = 0; step < n_steps; ++step) {
        const int buf = step % kNB;
        const uint32_t par = (step / kNB) & 1;
        const int nxt = step + kNB;
        const int v_ring = 2 * (step / kNB) + 1;
        const int k_ring = 2 * (nxt / kNB);
        if (buf == 0) wait_full(v_ring);                            // V tile of this step group
#pragma unroll
        for (int t = 0; t < kQTiles; ++t) {
          if constexpr (kNB == 1) {
            // split-P: the first 64 keys of P_t are signalled while the softmax still works on the rest
            mbar_wait(&p_full[2 * t + 0], par);
            tc_fence_after();
            issue_pv(t, step, 0, 0);
            mbar_wait(&p_full[2 * t + 1], par);
            tc_fence_after();
            issue_pv(t, step, 0, 4);
          } else {
            mbar_wait(&p_full[2 * t + buf], par);
            tc_fence_after();
            issue_pv(t, step, buf, 0);
          }
          commit(&pv_done[t]);
          if (nxt < n_steps) {
            if (t == 0 && (nxt % kNB) == 0) wait_full(k_ring);      // next K tile
            issue_qk(t, nxt, buf);
            commit(&s_full[2 * t + buf]);
          } else if (step == n_steps - 1) {
            commit(&o_full[t]);
          }
        }
        if (last_use(step)) commit(&kv_empty[v_ring % kStages]);
        if (nxt < n_steps && last_use(nxt)) commit(&kv_empty[k_ring % kStages]);
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
      const uint32_t tS = tmem_base + lane_addr + t * 128;
      const uint32_t tO = tmem_base + lane_addr + 256 + t * kHeadDim;
      const float scale_log2 = p.scale_log2;
      float m = -INFINITY;  // running row max in raw-logit units
      float l = 0.f;        // running row sum of exp2((s - m) * scale_log2)

      for (int step = 0; step < n_steps; ++step) {
        const int buf = step % kNB;
        mbar_wait(&s_full[2 * t + buf], (step / kNB) & 1);
        tc_fence_after();
        uint32_t sr[kStepN];
        tmem_ld_x32(tS + buf * kStepN, sr);
        tmem_ld_x32(tS + buf * kStepN + 32, sr + 32);
        if constexpr (kStepN == 128) {
          tmem_ld_x32(tS + buf * kStepN + 64, sr + 64);
          tmem_ld_x32(tS + buf * kStepN + 96, sr + 96);
        }
        tmem_wait_ld();

        if constexpr (kKeyMod) {
          if (p.key_logit_scale != nullptr) {
            const float4* ks = reinterpret_cast<const float4*>(p.key_logit_scale + step * kStepN);
#pragma unroll
            for (int c = 0; c < kStepN / 4; ++c) {
              // entries past Lk are masked below; the host pads the array to a multiple of 128
              const float4 w = __ldg(ks + c);
              sr[4 * c + 0] = __float_as_uint(__uint_as_float(sr[4 * c + 0]) * w.x);
              sr[4 * c + 1] = __float_as_uint(__uint_as_float(sr[4 * c + 1]) * w.y);
              sr[4 * c + 2] = __float_as_uint(__uint_as_float(sr[4 * c + 2]) * w.z);
              sr[4 * c + 3] = __float_as_uint(__uint_as_float(sr[4 * c + 3]) * w.w);
            }
          }
        }
        const int valid = k_len - step * kStepN;
        if (valid < kStepN) {
#pragma unroll
          for (int c = 0; c < kStepN; ++c) {
            if (c >= valid) sr[c] = 0xff800000u;  // -inf
          }
        }

        float mx0 = __uint_as_float(sr[0]), mx1 = __uint_as_float(sr[1]);
        float mx2 = __uint_as_float(sr[2]), mx3 = __uint_as_float(sr[3]);
#pragma unroll
        for (int c = 4; c < kStepN; c += 4) {
          mx0 = fmaxf(mx0, __uint_as_float(sr[c + 0]));
          mx1 = fmaxf(mx1, __uint_as_float(sr[c + 1]));
          mx2 = fmaxf(mx2, __uint_as_float(sr[c + 2]));
          mx3 = fmaxf(mx3, __uint_as_float(sr[c + 3]));
        }
        const float tile_max = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));

        if (step == 0) {
          m = tile_max;
        } else {
          const bool need = (tile_max - m) * scale_log2 > kRescaleThreshold;
          if (__any_sync(0xffffffffu, need)) {
            // O_t may still be accumulating PV_t(step-1): wait for it.  PV_t(step) cannot be issued
            // before we arrive on p_full below, so pv_done[t] is at most one phase ahead of us.
            mbar_wait(&pv_done[t], (step - 1) & 1);
            tc_fence_after();
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
        float2 sum_a = make_float2(0.f, 0.f), sum_b = make_float2(0.f, 0.f);
        const float* pvw = nullptr;
        if constexpr (kKeyMod) {
          if (p.key_pv_weight != nullptr) pvw = p.key_pv_weight + step * kStepN;
        }
#pragma unroll
        for (int h = 0; h < kStepN / 64; ++h) {
          uint32_t pk[32];
          softmax_exp64<kPolyEvery, kKeyMod>(sr + 64 * h, scale_log2, neg_ms, sum_a, sum_b, pk,
                                             pvw == nullptr ? nullptr : pvw + 64 * h);
          tmem_st_x32(tS + buf * kStepN + 32 * h, pk);
          tmem_wait_st();
          tc_fence_before();
          __syncwarp();
          // kStepN == 128: one barrier per 64-key half (split-P); kStepN == 64: one per S buffer
          if (lane == 0) mbar_arrive(&p_full[2 * t + (kNB == 1 ? h : buf)]);
        }
        l += (sum_a.x + sum_a.y) + (sum_b.x + sum_b.y);
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
