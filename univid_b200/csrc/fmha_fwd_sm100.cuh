// Flash-attention forward for Wan DiT tokens on sm_100a: head_dim 128, bf16 in / bf16 out,
// non-causal, optional per-batch key length mask (reference: flash_attention() k_lens,
// models/wan/utils/modules/attention.py:72-80) and optional per-key modifiers used by the
// text-weighted cross-attention (reference: Wan22ContextWrapper hook, models/model_pipeline.py:1756-1803).
//
// Work decomposition (persistent, "stream-K" over the key axis).  A *unit* is 256 query rows of one
// (batch, head) against all keys.  The grid is one CTA per SM (G CTAs).  The U units are dealt out in
// W = U / G full rounds (CTA g takes unit w*G + g, so concurrently running CTAs share ~1.15 heads' K/V in
// L2 exactly like a plain grid would); the remaining R = U - W*G units are NOT a partial wave: their
// R * n_kv key tiles are cut into G equal contiguous ranges, so every CTA does the same amount of work.
// A unit cut across CTAs is finished by the CTA that holds its last key tile (the owner): the other
// CTAs write their un-normalised partial (O fp32, row max, row sum) to a workspace slot and raise a
// flag; the owner merges the partials in its epilogue.  Non-owner ranges are processed first and the
// owner always waits on lower-numbered CTAs, so there are no wait chains.  With 1536 units on 148 SMs
// (cfg #2) this turns 11 waves into 10.38; with 384 units (4-way Ulysses) 3 waves into 2.59.
//
// Inside a CTA (3 warpgroups):
//   warps 0-3   softmax group 0  (query rows   0..127, one row per thread)   setmaxnreg.inc
//   warps 4-7   softmax group 1  (query rows 128..255)                       setmaxnreg.inc
//   warp  8     tcgen05.mma issuer (one elected thread) + TMEM allocator     setmaxnreg.dec
//   warp  9     TMA producer (one elected thread)                            setmaxnreg.dec
//   warps 10-11 idle (they only exist so the third warpgroup can give its registers away)
// TMEM (512 columns x 128 lanes):  S0 | S1 | O0 | O1.  P (bf16) is handed to the PV MMA in two 64-key halves
// (split-P) so the PV starts while the softmax still works on the second half.  The first half aliases the
// first 32 columns of its S and is the A operand straight from TMEM.  For long key sequences (kQBufs == 1) the
// second half goes through a 16 KiB shared-memory panel per tile instead: once the first half has arrived the
// whole of S_t sits in the softmax registers, so the NEXT score tile Q_t K_{j+1}^T is issued right after the
// first PV half -- QK^T leaves the softmax -> MMA -> softmax dependency chain and only 64 keys of PV remain
// on it.  K/V tiles (128 keys x 128 dims) alternate through one TMA ring; Q stays resident in smem for the
// segment and its buffer is reused to stage O for the TMA store.
//
// CTA pairs (kCtas == 2, long key sequences): a cluster of two CTAs on one TPC works on a 512-row unit -- CTA r
// owns query rows [q0 + 256 r, +256) as its two tiles -- and every MMA is one tcgen05.mma.cta_group::2 with M = 256
// covering tile t of BOTH CTAs, issued by the leader's warp 8.  Each CTA stages only HALF of every K tile (64 of
// the 128 keys: the N split of Q K^T) and HALF of every V tile (64 of the 128 dims: the N split of P V), so the
// tensor core of an SM reads 80 KiB instead of 112 KiB of shared memory per (query tile, key tile) and TMA writes
// 16 KiB instead of 32 KiB: 109 B/clk against the 128 B/clk an SM delivers (156 B/clk single-CTA, i.e. over the
// limit).  TMA loads of both CTAs count on the LEADER's q_full / kv_full barriers (cp.async.bulk.tensor
// .cta_group::2); s_full / pv_done / o_full / kv_empty are multicast commits that arrive in both CTAs; p_full lives
// in the leader and collects the 4 + 4 softmax warps of both CTAs (remote mbarrier.arrive.release.cluster).
// Softmax, lazy rescale, epilogue, stream-K partials and peer-rank stores stay per CTA; the schedule runs over pairs.
//
// Ordering facts the protocol relies on: tcgen05 MMAs issued by one thread execute in order, so
// S_t(s+1) overwriting the buffer that held P_t(s) is safe once PV_t(s) has been issued before it; the
// softmax warps rescale O_t in place (lazily, only when a row max grows by more than 2^8) after waiting
// for pv_done[t] of the previous step, and PV_t(s) is not issued before they arrive on p_full[t].
// All mbarrier parities are derived from running counters (segments, steps, ring slots) that every role
// advances identically.
#pragma once
#include "ptx.cuh"

namespace uvb {

constexpr int kBlockM = 128;          // query rows per tile (UMMA M)
constexpr int kBlockN = 128;          // keys per K/V smem tile == keys per softmax/MMA step
constexpr int kHeadDim = 128;
constexpr int kQTiles = 2;            // query tiles per CTA
constexpr int kUnitRows = kQTiles * kBlockM;
constexpr int kTileBytes = kBlockN * kHeadDim * 2;   // 32 KiB, one [128 x 128] bf16 tile
constexpr int kHalfTile = kTileBytes / 2;            // one 64-column swizzle panel
constexpr int kFmhaThreads = 384;
constexpr int kRegsSoftmax = 224;                    // 2*128*224 + 128*56 = 64512 <= 65536
constexpr int kRegsOther = 56;
constexpr float kRescaleThreshold = 8.0f;            // lazy rescale: only when max grows by > 2^8

// split-unit workspace slot (one per CTA): O fp32 column-major [128 dims][256 rows], then m[256], l[256]
constexpr int kWsOFloats = kHeadDim * kUnitRows;
constexpr int kWsSlotFloats = kWsOFloats + 2 * kUnitRows;

// kStages K/V ring slots, kQBufs query-block buffers.  Self-attention (hundreds of key tiles per query
// block) uses <4, 1>; cross-attention (4 key tiles per query block) uses <3, 2>: the next block's Q is
// prefetched while the current one is computed and its O tile drains through the other buffer.
template <int kStages, int kQBufs, int kCtas = 1, bool kEarlyS = false>
struct FmhaSmem {
  static_assert(kCtas == 1 || kCtas == 2, "one CTA or a CTA pair");
  static_assert(!kEarlyS || kQBufs == 1, "early S release exists for the long-key variant only");
  static constexpr int kStageBytes = kTileBytes / kCtas;   // a pair CTA stages half of every K / V tile
  static constexpr int kQOff = 0;
  static constexpr int kKvOff = kQBufs * kQTiles * kTileBytes;
  // The second 64-key half of P goes through shared memory (16 KiB per query tile, K-major SWIZZLE_128B like Q) so
  // that the next score tile can be issued before it is consumed -- whenever the panels fit: always for the
  // long-key variant (one query-block buffer), and for the query-block-pipelined short-key variant when it runs as
  // CTA pairs (half-size K/V stages leave room for two Q/O buffers AND the panels)
  static constexpr bool kPSmem = kQBufs == 1 || kCtas == 2;
  static constexpr int kPOff = kKvOff + kStages * kStageBytes;
  // kEarlyS: BOTH 64-key halves of P go through shared memory (two 16 KiB panels per query tile), nothing of P
  // aliases S_t any more, and S_t is handed back to the MMA warp as soon as the softmax holds it in registers
  static constexpr int kPPanels = kPSmem ? (kEarlyS ? 2 : 1) : 0;
  static constexpr int kBarOff = kPOff + kQTiles * kPPanels * kHalfTile;
  // barriers: q_full[B][2] q_empty[B][2] kv_full[S] kv_empty[S] s_full[2] p_full[2][2] pv_done[2] o_full[2] s_free[2]
  // + tmem ptr
  static constexpr int kNumBars = 4 * kQBufs + 2 * kStages + 12;
  static constexpr int kBytes = kBarOff + kNumBars * 8 + 16;
  static constexpr int kDynBytes = kBytes + 1024;  // slack for 1024 B alignment
};

struct FmhaParams {
  CUtensorMap tm_q;   // dims (d, token, head, batch), box (64, 128, 1, 1), SWIZZLE_128B
  CUtensorMap tm_k;
  CUtensorMap tm_kh;  // same tensor as tm_k, box (64, 64, 1, 1): the 64-key half a pair CTA stages
  CUtensorMap tm_v;
  CUtensorMap tm_o;
  const int* k_lens;              // [B] or nullptr (= Lk)
  const float* key_logit_scale;   // [Lk] or nullptr : logits[:, j] *= key_logit_scale[j]
  const float* key_pv_weight;     // [Lk] or nullptr : P[:, j] *= w[j] after the row sum
  const float* out_bias;          // [N*128] or nullptr : added to the normalised output
  float* ws;                      // [gridDim.x][kWsSlotFloats] or nullptr (then units are never split)
  uint32_t* flags;                // [gridDim.x][2], zero on entry and on exit
  unsigned long long* timeline;   // diagnostics (uvb_debug_fmha_timeline) or nullptr: per CTA 32 x u64
  unsigned long long* prof;       // UVB_FMHA_PROFILE builds: per CTA 16 x u64 wait counters, or nullptr
  // Fused Ulysses output exchange: when n_peers > 0 the rows [j*chunk, (j+1)*chunk) of the output belong
  // to rank j and are TMA-stored straight into rank j's [B, chunk, N_total, 128] buffer (tm_o_peer[j],
  // mapped over NVLink) at head offset o_head_off; tm_o is unused.
  int n_peers;
  int chunk;
  int o_head_off;
  int o_heads;                    // total heads of the peers' [B, chunk, o_heads, 128] buffers
  __nv_bfloat16* o_peer_ptr[8];
  CUtensorMap tm_o_peer[8];
  int Lq;
  int Lk;
  int n_qt;                       // query blocks (256 rows, 512 for CTA pairs) per (batch, head)
  int N;                          // heads
  int n_units;                    // B * N * n_qt
  float scale_log2;               // softmax_scale * log2(e)
};

// One contiguous piece of work of a CTA: key tiles [a, b) of `unit`.  `owner` pieces produce the output
// rows (after merging the partials of CTAs g-1, g-2, ... when a > 0); the others write a partial.
struct FmhaSeg {
  int unit, a, b;
  bool owner;
};

// The static schedule; every role of the CTA evaluates it with identical (warp-uniform) results.
struct FmhaSched {
  int G, g, W, n_kv, n_seg;
  long long rem_total;
  FmhaSeg rem0, rem1;   // remainder pieces in processing order (non-owner first)

  __device__ __forceinline__ long long rem_lo(int cta) const { return rem_total * cta / G; }

  // n_workers CTAs (or CTA pairs) walk the unit list; `worker` is this CTA's (pair's) index
  __device__ __forceinline__ void init(const FmhaParams& p, bool split, int n_workers, int worker) {
    G = n_workers;
    g = worker;
    n_kv = (p.Lk + kBlockN - 1) / kBlockN;
    W = p.n_units / G;
    const int R = p.n_units - W * G;
    int n_rem = 0;
    rem_total = 0;
    rem0 = rem1 = FmhaSeg{0, 0, 0, true};
    if (!split) {
      if (g < R) {
        rem0 = FmhaSeg{W * G + g, 0, n_kv, true};
        n_rem = 1;
      }
    } else {
      rem_total = static_cast<long long>(R) * n_kv;
      const long long lo = rem_lo(g), hi = rem_lo(g + 1);
      if (hi > lo) {
        const int u0 = static_cast<int>(lo / n_kv), a0 = static_cast<int>(lo - static_cast<long long>(u0) * n_kv);
        const int len = static_cast<int>(hi - lo);
        if (a0 + len <= n_kv) {
          rem0 = FmhaSeg{W * G + u0, a0, a0 + len, a0 + len == n_kv};
          n_rem = 1;
        } else {
          rem0 = FmhaSeg{W * G + u0 + 1, 0, a0 + len - n_kv, false};   // head of the next unit first
          rem1 = FmhaSeg{W * G + u0, a0, n_kv, true};
          n_rem = 2;
        }
      }
    }
    n_seg = W + n_rem;
  }
  __device__ __forceinline__ FmhaSeg seg(int i) const {
    if (i < W) return FmhaSeg{i * G + g, 0, n_kv, true};
    return i == W ? rem0 : rem1;
  }
};

// exp2 of 2*kPairs scores of one row -> kPairs packed bf16x2 probabilities + partial row sums.  One pair in every
// kPolyEvery goes through the FMA-pipe polynomial instead of MUFU.EX2 (0 = never): the XU pipe does 16
// exp2/clk/SM, exactly the rate at which the tensor pipe consumes a 128x128 tile, so offloading a share
// of them is what lets the softmax keep ahead of the MMAs.
// Lab builds only (-DUVB_FMHA_PROFILE): cycle counters of the barrier waits, per CTA 16 x u64 in FmhaParams::prof
// [0..2] softmax group 0: total, waiting for S, waiting for PV; [3..5] group 1; [6..10] MMA warp: total, waiting for
// P half 0, P half 1, K/V tiles, Q; [11] steps; [12..13] softmax group 0: unit set-up (top of a segment to its first
// step), epilogue (end of the step loop to the end of the segment); [14..15] group 1
#ifdef UVB_FMHA_PROFILE
#define UVB_PROF(acc, stmt)          \
  do {                               \
    const long long t0__ = clock64(); \
    stmt;                            \
    acc += clock64() - t0__;         \
  } while (0)
#else
#define UVB_PROF(acc, stmt) stmt
#endif

template <int kPairs, bool kKeyMod, int kPolyEvery>
__device__ __forceinline__ void softmax_exp(const uint32_t* sr, float scale_log2, float neg_ms,
                                            float2& sum_a, float2& sum_b, uint32_t* pk,
                                            const float* pvw) {
  const float2 sc = make_float2(scale_log2, scale_log2);
  const float2 nm = make_float2(neg_ms, neg_ms);
#pragma unroll
  for (int i = 0; i < kPairs; ++i) {
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

// kPoly: one exp2 pair in every kPoly is evaluated on the FMA pipe instead of MUFU.EX2 (0 = never; UVB_KNOB_FMHA_POLY)
template <int kStages, int kQBufs, int kCtas, bool kKeyMod, bool kEarlyS = false, int kPoly = 0>
__global__ void __launch_bounds__(kFmhaThreads, 1)
fmha_fwd_kernel(const __grid_constant__ FmhaParams p) {
  using SM = FmhaSmem<kStages, kQBufs, kCtas, kEarlyS>;
  constexpr int kStageBytes = SM::kStageBytes;
  static_assert(SM::kDynBytes <= 232448, "shared memory budget (227 KiB)");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  // opaque to the optimiser: otherwise every mbarrier access re-derives the aligned base (S2UR CgaCtaId,
  // SWINHI, ULEA, ... ~12 instructions each) inside the issue-sensitive softmax loop
  asm volatile("" : "+l"(smem));
  uint8_t* smem_q = smem + SM::kQOff;
  uint8_t* smem_kv = smem + SM::kKvOff;
  uint8_t* smem_p = smem + SM::kPOff;
  constexpr bool kPSmem = SM::kPSmem;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SM::kBarOff);
  uint64_t* q_full = bars;                       // [buf][tile]  TMA -> MMA: Q_t of the segment landed
  uint64_t* q_empty = bars + 2 * kQBufs;         // [buf][tile]  softmax -> TMA: Q_t / O staging is free again
  uint64_t* kv_full = bars + 4 * kQBufs;         // [kStages]
  uint64_t* kv_empty = kv_full + kStages;        // [kStages]
  uint64_t* s_full = kv_empty + kStages;         // [tile]
  uint64_t* p_full = s_full + 2;                 // [tile][half]
  uint64_t* pv_done = p_full + 4;                // [tile]
  uint64_t* o_full = pv_done + 2;                // [tile]
  uint64_t* s_free = o_full + 2;                 // [tile]  softmax -> MMA (kEarlyS): S_t is in registers, overwrite it
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(s_free + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t cta_rank = kCtas == 2 ? cluster_ctarank() : 0u;
  const bool leader = cta_rank == 0;

  if (threadIdx.x == 0) {
    for (int i = 0; i < 2 * kQBufs; ++i) {
      mbar_init(&q_full[i], 1);
      mbar_init(&q_empty[i], 1);
    }
    for (int t = 0; t < 2; ++t) {
      mbar_init(&s_full[t], 1);
      mbar_init(&p_full[2 * t], 4 * kCtas);       // one arrive per softmax warp (of both CTAs of a pair)
      mbar_init(&p_full[2 * t + 1], 4 * kCtas);
      mbar_init(&pv_done[t], 1);
      mbar_init(&o_full[t], 1);
      mbar_init(&s_free[t], 4 * kCtas);
    }
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    fence_mbar_init();
  }
  if (warp == 8) {
    if constexpr (kCtas == 2) {
      tmem_alloc_pair(tmem_ptr, 512);
    } else {
      tmem_alloc(tmem_ptr, 512);
      tmem_relinquish();
    }
  }
  if (warp == 9 && lane == 0) {
    tma_prefetch_desc(&p.tm_q);
    tma_prefetch_desc(kCtas == 2 ? &p.tm_kh : &p.tm_k);
    tma_prefetch_desc(&p.tm_v);
    if (p.n_peers == 0) tma_prefetch_desc(&p.tm_o);
  }
  tc_fence_before();
  if constexpr (kCtas == 2) {
    cluster_sync_all();     // the peer's barriers are initialised and its TMEM is allocated
    __syncthreads();        // (subsumed by the cluster barrier; spelled out for compute-sanitizer's racecheck, which
                            // does not take barrier.cluster as ordering the tcgen05.alloc write of tmem_ptr)
  } else {
    __syncthreads();
  }
  tc_fence_after();
  // warp-uniform copy (a plain smem load is not provably uniform and would force ptxas to wrap every
  // tcgen05 instruction in an R2UR.BROADCAST waterfall loop)
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr, 0);

  FmhaSched sch;
  sch.init(p, p.ws != nullptr, static_cast<int>(gridDim.x) / kCtas, static_cast<int>(blockIdx.x) / kCtas);
  // stream-K workspace slots and flags are per CTA: CTA r of pair g exchanges partials with CTA r of other pairs
  const int my_slot = static_cast<int>(blockIdx.x);
  auto slot_of = [&](int worker) { return worker * kCtas + static_cast<int>(cta_rank); };

  // (batch, head, first query row) of a unit and its key-tile range clipped to the batch's key length
  auto decode = [&](const FmhaSeg& sg, int& batch, int& head, int& q_row0, int& k_len, int& ka, int& kb) {
    const int qt = sg.unit % p.n_qt;
    const int bh = sg.unit / p.n_qt;
    head = bh % p.N;
    batch = bh / p.N;
    q_row0 = (qt * kCtas + static_cast<int>(cta_rank)) * kUnitRows;
    k_len = p.Lk;
    if (p.k_lens != nullptr) k_len = min(max(__ldg(p.k_lens + batch), 0), p.Lk);
    const int n_kv_b = (k_len + kBlockN - 1) / kBlockN;
    kb = min(sg.b, n_kv_b);
    ka = min(sg.a, kb);
  };

  if (warp >= 8) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegsOther));
    if (warp == 9) {
      // ===================================== TMA producer =====================================
      // the whole warp runs the loop (converged, uniform operands); one elected lane issues
      int ring = 0;
      // pair: both CTAs' loads report to the LEADER's full barriers (cluster addresses)
      const uint32_t q_full_cl = kCtas == 2 ? map_to_cta(&q_full[0], 0) : 0u;
      const uint32_t kv_full_cl = kCtas == 2 ? map_to_cta(&kv_full[0], 0) : 0u;
      for (int si = 0; si < sch.n_seg; ++si) {
        const FmhaSeg sg = sch.seg(si);
        int batch, head, q_row0, k_len, ka, kb;
        decode(sg, batch, head, q_row0, k_len, ka, kb);
        const int qb = si % kQBufs;                      // query-block buffer of this segment
        const uint32_t qpar = (si / kQBufs) & 1;
        for (int t = 0; t < kQTiles; ++t) {
          uint64_t* full = &q_full[2 * qb + t];
          mbar_wait(&q_empty[2 * qb + t], qpar ^ 1);
          uint8_t* dst = smem_q + (2 * qb + t) * kTileBytes;
          if (elect_one()) {
            if constexpr (kCtas == 2) {
              if (leader) mbar_arrive_expect_tx(full, 2 * kTileBytes);
              const uint32_t fb = q_full_cl + (2 * qb + t) * 8;
              tma_load_4d_pair(dst, &p.tm_q, fb, 0, q_row0 + t * kBlockM, head, batch, kEvictFirst);
              tma_load_4d_pair(dst + kHalfTile, &p.tm_q, fb, 64, q_row0 + t * kBlockM, head, batch, kEvictFirst);
            } else {
              mbar_arrive_expect_tx(full, kTileBytes);
              tma_load_4d_hint(dst, &p.tm_q, full, 0, q_row0 + t * kBlockM, head, batch, kEvictFirst);
              tma_load_4d_hint(dst + kHalfTile, &p.tm_q, full, 64, q_row0 + t * kBlockM, head, batch, kEvictFirst);
            }
          }
          __syncwarp();
        }
        // ring order matches consumption order: K(ka), V(ka), K(ka+1), V(ka+1), ...
        for (int i = 2 * ka; i < 2 * kb; ++i, ++ring) {
          const int stage = ring % kStages;
          mbar_wait(&kv_empty[stage], ((ring / kStages) & 1) ^ 1);
          uint8_t* dst = smem_kv + stage * kStageBytes;
          const int key0 = (i >> 1) * kBlockN;
          if (elect_one()) {
            if constexpr (kCtas == 2) {
              if (leader) mbar_arrive_expect_tx(&kv_full[stage], 2 * kStageBytes);
              const uint32_t fb = kv_full_cl + stage * 8;
              if (i & 1) {
                // V: all 128 keys x this CTA's 64 dims (one panel): the N half of P V
                tma_load_4d_pair(dst, &p.tm_v, fb, 64 * static_cast<int>(cta_rank), key0, head, batch, kEvictLast);
              } else {
                // K: this CTA's 64 keys x 128 dims (two panels of 8 KiB): the N half of Q K^T
                const int kr = key0 + 64 * static_cast<int>(cta_rank);
                tma_load_4d_pair(dst, &p.tm_kh, fb, 0, kr, head, batch, kEvictLast);
                tma_load_4d_pair(dst + kHalfTile / 2, &p.tm_kh, fb, 64, kr, head, batch, kEvictLast);
              }
            } else {
              const CUtensorMap* tm = (i & 1) ? &p.tm_v : &p.tm_k;
              mbar_arrive_expect_tx(&kv_full[stage], kTileBytes);
              tma_load_4d_hint(dst, tm, &kv_full[stage], 0, key0, head, batch, kEvictLast);
              tma_load_4d_hint(dst + kHalfTile, tm, &kv_full[stage], 64, key0, head, batch, kEvictLast);
            }
          }
          __syncwarp();
        }
      }
    } else if (warp == 8 && leader) {
      // ===================================== MMA issuer (pair: leader CTA only) ===============
      // The whole warp runs the control flow converged so every operand is warp-uniform; one elected
      // lane (always the same one) issues the tcgen05.mma / tcgen05.commit instructions.  Descriptors are
      // built once; per instruction only a 64-bit add of a compile-time offset remains.  (Round-1 ncu:
      // with the loop inside `if (lane == 0)` ptxas emitted an ELECT/R2UR.BROADCAST waterfall around each
      // UTCHMMA and descriptor math on the uniform datapath -- ~110 cycles per MMA, the real bottleneck.)
      constexpr uint32_t idesc_qk = umma_idesc_bf16(kBlockM * kCtas, kBlockN, 0, 0);
      constexpr uint32_t idesc_pv = umma_idesc_bf16(kBlockM * kCtas, kHeadDim, 0, 1);
      const uint64_t q_desc = umma_desc_sw128(smem_u32(smem_q), 16, 1024);            // K-major A
      const uint64_t k_desc = umma_desc_sw128(smem_u32(smem_kv), 16, 1024);           // K-major B
      const uint64_t v_desc = umma_desc_sw128(smem_u32(smem_kv), kHalfTile, 1024);    // MN-major B
      constexpr uint64_t kTile16 = kTileBytes >> 4;           // descriptor address units are 16 B
      constexpr uint64_t kStage16 = kStageBytes >> 4;
      constexpr int kKPanel = kHalfTile / kCtas;              // bytes of one 64-dim panel of the staged K rows
      auto mma_ss = [](uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
        if constexpr (kCtas == 2) umma_ss_pair(d, a, b, idesc, acc); else umma_ss(d, a, b, idesc, acc);
      };
      auto mma_ts = [](uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
        if constexpr (kCtas == 2) umma_ts_pair(d, a, b, idesc, acc); else umma_ts(d, a, b, idesc, acc);
      };
      // p_full also collects remote arrives of the peer CTA's softmax warps (default .release.cta arrives: what they
      // publish -- P in the peer's TMEM / shared memory -- is consumed by the peer's half of the pair MMA)
      auto wait_p = [&](uint64_t* bar, uint32_t par) {
        mbar_wait(bar, par);
        tc_fence_after();
      };

      auto wait_full = [&](int r) {
        mbar_wait(&kv_full[r % kStages], (r / kStages) & 1);
        tc_fence_after();
      };
      auto commit = [&](uint64_t* bar) {     // pair: the arrive is multicast to the same barrier of both CTAs
        if (elect_one()) {
          if constexpr (kCtas == 2) tc_commit_pair(bar); else tc_commit(bar);
        }
        __syncwarp();
      };
      // S_t = Q_t K^T : A, B K-major, N = 128 keys; 8 K-steps of 16 dims: panel = kk/4, 32 B per K-step
      // inside the 128 B swizzle atom
      auto issue_qk = [&](int qt, int t, int k_ring) {   // qt = 2 * query buffer + tile
        const uint64_t qa = q_desc + qt * kTile16;
        const uint64_t ka = k_desc + (k_ring % kStages) * kStage16;
        const uint32_t d = tmem_base + t * 128;
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < kHeadDim / 16; ++kk) {
            const uint64_t aoff = ((kk >> 2) * kHalfTile + (kk & 3) * 32) >> 4;
            const uint64_t boff = ((kk >> 2) * kKPanel + (kk & 3) * 32) >> 4;
            mma_ss(d, qa + aoff, ka + boff, idesc_qk, kk > 0 ? 1u : 0u);
          }
        }
        __syncwarp();
      };
      // O_t (+)= P_t V : A = P from TMEM (8 columns per 16 keys), B = V rows, MN-major: 16 keys = 2 KiB
      // per K-step, the two 64-dim panels are kHalfTile apart (LBO), 8-key groups 1 KiB apart (SBO)
      // second half of P from shared memory (kPSmem): A K-major like Q, keys 64..127 of the V tile
      const uint64_t p_desc = umma_desc_sw128(smem_u32(smem_p), 16, 1024);
      auto issue_pv_smem = [&](int t, int v_ring) {
        const uint64_t va = v_desc + (v_ring % kStages) * kStage16;
        const uint64_t pa = p_desc + t * (kHalfTile >> 4);
        const uint32_t d = tmem_base + 256 + t * kHeadDim;
        if (elect_one()) {
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4) {
            mma_ss(d, pa + ((k4 * 32) >> 4), va + (4 + k4) * (2048 >> 4), idesc_pv, 1u);
          }
        }
        __syncwarp();
      };
      // kEarlyS: 64-key half `half` of P_t from its own panel (2 t + half); the very first MMA of a segment overwrites O
      auto issue_pv_panel = [&](int t, int v_ring, int half, bool first_step) {
        const uint64_t va = v_desc + (v_ring % kStages) * kStage16;
        const uint64_t pa = p_desc + (2 * t + half) * (kHalfTile >> 4);
        const uint32_t d = tmem_base + 256 + t * kHeadDim;
        const uint32_t acc0 = (first_step && half == 0) ? 0u : 1u;
        if (elect_one()) {
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4) {
            mma_ss(d, pa + ((k4 * 32) >> 4), va + (4 * half + k4) * (2048 >> 4), idesc_pv, k4 > 0 ? 1u : acc0);
          }
        }
        __syncwarp();
      };
      auto issue_pv = [&](int t, int v_ring, bool first_step, int kk0) {   // 4 K-steps = 64 keys from kk0
        const uint64_t va = v_desc + (v_ring % kStages) * kStage16;
        const uint32_t d = tmem_base + 256 + t * kHeadDim;
        const uint32_t a = tmem_base + t * 128;
        const uint32_t acc0 = (!first_step || kk0 > 0) ? 1u : 0u;
        if (elect_one()) {
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4) {
            const int kk = kk0 + k4;
            mma_ts(d, a + kk * 8, va + kk * (2048 >> 4), idesc_pv, k4 > 0 ? 1u : acc0);
          }
        }
        __syncwarp();
      };

      int ring = 0;     // K/V tiles consumed so far (all segments)
      int gstep = 0;    // softmax/MMA steps so far (all segments)
      [[maybe_unused]] long long pf_p0 = 0, pf_p1 = 0, pf_kv = 0, pf_q = 0;
      [[maybe_unused]] const long long pf_start = clock64();
      // kQBufs == 2 (few key tiles per unit): the next unit's query block sits in the other buffer long before this
      // unit ends, so its first score tiles are issued inside this unit's LAST step, at the point where a middle
      // step issues the scores of its successor.  The unit boundary then looks like a step boundary to the tensor
      // pipe: the softmax groups find S ready when they leave the epilogue instead of waiting for both tiles' last
      // P, two PVs, and two Q K^T (wait-cycle profile of round 2: ~3500 cycles per 4-step unit).  Only behind a unit
      // of >= 2 steps: the buffer the next Q loads into is released one softmax step into this unit.
      bool scores_ahead = false;   // this unit's first score tiles were issued by the previous unit
      for (int si = 0; si < sch.n_seg; ++si) {
        const FmhaSeg sg = sch.seg(si);
        int batch, head, q_row0, k_len, ka, kb;
        decode(sg, batch, head, q_row0, k_len, ka, kb);
        const int n_steps = kb - ka;
        const int qb = si % kQBufs;
        const uint32_t qpar = (si / kQBufs) & 1;
        const bool primed = scores_ahead;
        scores_ahead = false;
        [[maybe_unused]] int nqb = 0;
        [[maybe_unused]] uint32_t nqpar = 0;
        if constexpr (kQBufs == 2 && SM::kPSmem && !kEarlyS) {
          if (n_steps >= 2 && si + 1 < sch.n_seg) {
            int nb, nh, nq0, nkl, nka, nkb;
            decode(sch.seg(si + 1), nb, nh, nq0, nkl, nka, nkb);
            scores_ahead = nkb > nka;
            nqb = (si + 1) % kQBufs;
            nqpar = ((si + 1) / kQBufs) & 1;
          }
        }
        UVB_PROF(pf_q, mbar_wait(&q_full[2 * qb], qpar));
        tc_fence_after();
        if (n_steps == 0) {
          // nothing to attend to (k_lens clipped the range away): the epilogue treats O as zero
          mbar_wait(&q_full[2 * qb + 1], qpar);
          commit(&o_full[0]);
          commit(&o_full[1]);
          continue;
        }
        if (!primed) {
          // prologue: scores of the first step of both tiles
          UVB_PROF(pf_kv, wait_full(ring));
          issue_qk(2 * qb, 0, ring);
          commit(&s_full[0]);
          mbar_wait(&q_full[2 * qb + 1], qpar);
          tc_fence_after();
          issue_qk(2 * qb + 1, 1, ring);
          commit(&s_full[1]);
          commit(&kv_empty[ring % kStages]);   // K tile of step 0 is fully consumed by the prologue
        }

        for (int step = 0; step < n_steps; ++step) {
          const uint32_t par = (gstep + step) & 1;
          const int v_ring = ring + 2 * step + 1;
          const int k_ring = v_ring + 1;          // K tile of the next step (last step: first K tile of the next unit)
          const bool more = step + 1 < n_steps;
          const bool ahead = more || scores_ahead;   // a successor score tile goes out during this step
          UVB_PROF(pf_kv, wait_full(v_ring));
          if constexpr (kEarlyS) {
            // S_t is free as soon as the softmax group holds it in registers (s_free), long before any of P_t exists:
            // the next score tile is issued then, so Q K^T is off the softmax -> MMA -> softmax chain altogether and
            // the (cross-CTA) signalling latencies hide under a whole softmax step.  Both halves of P come through
            // shared memory.
#pragma unroll
            for (int t = 0; t < kQTiles; ++t) {
              if (more) {
                UVB_PROF(pf_p0, wait_p(&s_free[t], par));
                if (t == 0) UVB_PROF(pf_kv, wait_full(k_ring));
                issue_qk(2 * qb + t, t, k_ring);
                commit(&s_full[t]);
              }
              UVB_PROF(pf_p0, wait_p(&p_full[2 * t + 0], par));
              issue_pv_panel(t, v_ring, 0, step == 0);
              UVB_PROF(pf_p1, wait_p(&p_full[2 * t + 1], par));
              issue_pv_panel(t, v_ring, 1, step == 0);
              commit(&pv_done[t]);
              if (!more) commit(&o_full[t]);
            }
            commit(&kv_empty[v_ring % kStages]);
            if (more) commit(&kv_empty[k_ring % kStages]);
            continue;
          }
#pragma unroll
          for (int t = 0; t < kQTiles; ++t) {
            // split-P: the first 64 keys of P_t are signalled while the softmax still works on the rest
            UVB_PROF(pf_p0, wait_p(&p_full[2 * t + 0], par));
            issue_pv(t, v_ring, step == 0, 0);
            if constexpr (kPSmem) {
              // The first half of P has arrived, so all of S_t sits in the softmax registers, and the PV just
              // issued reads P[0, 32) before anything issued after it runs: the next score tile can go out NOW,
              // while the softmax still exponentiates the second half (which it hands over through shared
              // memory, not through S_t).  Only the last 64 keys of PV stay on the softmax -> MMA -> softmax
              // dependency chain; QK^T leaves it.
              if (ahead) {
                if (t == 0) UVB_PROF(pf_kv, wait_full(k_ring));
                if (!more) {
                  UVB_PROF(pf_q, mbar_wait(&q_full[2 * nqb + t], nqpar));
                  tc_fence_after();
                }
                issue_qk(more ? 2 * qb + t : 2 * nqb + t, t, k_ring);
                commit(&s_full[t]);
              }
              UVB_PROF(pf_p1, wait_p(&p_full[2 * t + 1], par));
              issue_pv_smem(t, v_ring);
              commit(&pv_done[t]);
              if (!more) commit(&o_full[t]);
            } else {
              UVB_PROF(pf_p1, wait_p(&p_full[2 * t + 1], par));
              issue_pv(t, v_ring, step == 0, 4);
              commit(&pv_done[t]);
              if (more) {
                if (t == 0) wait_full(k_ring);
                issue_qk(2 * qb + t, t, k_ring);
                commit(&s_full[t]);
              } else {
                commit(&o_full[t]);
              }
            }
          }
          commit(&kv_empty[v_ring % kStages]);
          if (ahead) commit(&kv_empty[k_ring % kStages]);
        }
        ring += 2 * n_steps;
        gstep += n_steps;
      }
#ifdef UVB_FMHA_PROFILE
      if (p.prof != nullptr && lane == 0) {
        unsigned long long* pr = p.prof + blockIdx.x * 16;
        pr[6] = clock64() - pf_start;
        pr[7] = pf_p0;
        pr[8] = pf_p1;
        pr[9] = pf_kv;
        pr[10] = pf_q;
        pr[11] = gstep;
      }
#endif
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRegsSoftmax));
    // ===================================== softmax groups ====================================
    const int t = warp >> 2;
    const int wq = warp & 3;
    const int row = wq * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(wq * 32) << 16;
    const uint32_t tS = tmem_base + lane_addr + t * 128;
    const uint32_t tO = tmem_base + lane_addr + 256 + t * kHeadDim;
    const float scale_log2 = p.scale_log2;
    const int ws_row = t * kBlockM + row;          // row inside the unit
    // pair: p_full / s_free live in the leader CTA
    const uint32_t p_full_cl = kCtas == 2 ? map_to_cta(&p_full[0], 0) : 0u;
    const uint32_t s_free_cl = kCtas == 2 ? map_to_cta(&s_free[0], 0) : 0u;
    bool stored = false;
    [[maybe_unused]] long long pf_s = 0, pf_pv = 0;
    [[maybe_unused]] const long long pf_start = clock64();
    int pending_qb = -1;   // kQBufs == 2: query buffer whose O store was issued but not yet released
    int gstep = 0;
    // With two query buffers the release of a buffer (its O store must have finished READING the staging
    // area) is not needed before the producer loads the block after next, so the storing thread does not
    // wait right after the store: it releases the buffer one softmax step into the next segment.
    auto release_pending = [&]() {
      if (pending_qb >= 0) {
        tma_store_wait_read0();
        mbar_arrive(&q_empty[2 * pending_qb + t]);
        pending_qb = -1;
      }
    };
    if (p.timeline != nullptr && threadIdx.x == 0) {
      unsigned smid;
      asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
      p.timeline[blockIdx.x * 32 + 0] = smid;
      p.timeline[blockIdx.x * 32 + 1] = globaltimer_ns();
    }

    [[maybe_unused]] long long pf_setup = 0, pf_epi = 0, pf_mark = 0;
    for (int si = 0; si < sch.n_seg; ++si) {
#ifdef UVB_FMHA_PROFILE
      const long long pf_top = clock64();
      if (pf_mark != 0) pf_epi += pf_top - pf_mark;
#endif
      if (p.timeline != nullptr && threadIdx.x == 0 && si > 0 && si < 30)
        p.timeline[blockIdx.x * 32 + 1 + si] = globaltimer_ns();
      const FmhaSeg sg = sch.seg(si);
      int batch, head, q_row0, k_len, ka, kb;
      decode(sg, batch, head, q_row0, k_len, ka, kb);
      const int n_steps = kb - ka;
      float m = -INFINITY;  // running row max in raw-logit units
      float l = 0.f;        // running row sum of exp2((s - m) * scale_log2)
#ifdef UVB_FMHA_PROFILE
      pf_setup += clock64() - pf_top;
#endif

      for (int step = 0; step < n_steps; ++step) {
        const uint32_t par = (gstep + step) & 1;
        const int key0 = (ka + step) * kBlockN;
        UVB_PROF(pf_s, mbar_wait(&s_full[t], par));
        tc_fence_after();
        uint32_t sr[kBlockN];
        tmem_ld_x32(tS, sr);
        tmem_ld_x32(tS + 32, sr + 32);
        tmem_ld_x32(tS + 64, sr + 64);
        tmem_ld_x32(tS + 96, sr + 96);
        tmem_wait_ld();
        if constexpr (kEarlyS) {
          // S_t is in registers: hand the TMEM tile back (the last step of a segment has no successor to wait for
          // it, but arriving unconditionally keeps the barrier phase equal to the step count)
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if constexpr (kCtas == 2) {
              mbar_arrive_remote(s_free_cl + t * 8);
            } else {
              mbar_arrive(&s_free[t]);
            }
          }
        }

        if constexpr (kKeyMod) {
          if (p.key_logit_scale != nullptr) {
            const float4* ks = reinterpret_cast<const float4*>(p.key_logit_scale + key0);
#pragma unroll
            for (int c = 0; c < kBlockN / 4; ++c) {
              // entries past Lk are masked below; the host pads the array to a multiple of 128
              const float4 w = __ldg(ks + c);
              sr[4 * c + 0] = __float_as_uint(__uint_as_float(sr[4 * c + 0]) * w.x);
              sr[4 * c + 1] = __float_as_uint(__uint_as_float(sr[4 * c + 1]) * w.y);
              sr[4 * c + 2] = __float_as_uint(__uint_as_float(sr[4 * c + 2]) * w.z);
              sr[4 * c + 3] = __float_as_uint(__uint_as_float(sr[4 * c + 3]) * w.w);
            }
          }
        }
        const int valid = k_len - key0;   // only the last key tile of a sequence can be ragged
        if (valid < kBlockN) {
#pragma unroll
          for (int c = 0; c < kBlockN; ++c) {
            if (c >= valid) sr[c] = 0xff800000u;  // -inf
          }
        }

        // four independent chains of three-input maxima (FMNMX3): 64 instructions for 128 scores
        float mx0 = __uint_as_float(sr[0]), mx1 = __uint_as_float(sr[1]);
        float mx2 = __uint_as_float(sr[2]), mx3 = __uint_as_float(sr[3]);
#pragma unroll
        for (int c = 4; c + 8 <= kBlockN; c += 8) {
          mx0 = fmax3(mx0, __uint_as_float(sr[c + 0]), __uint_as_float(sr[c + 1]));
          mx1 = fmax3(mx1, __uint_as_float(sr[c + 2]), __uint_as_float(sr[c + 3]));
          mx2 = fmax3(mx2, __uint_as_float(sr[c + 4]), __uint_as_float(sr[c + 5]));
          mx3 = fmax3(mx3, __uint_as_float(sr[c + 6]), __uint_as_float(sr[c + 7]));
        }
        mx0 = fmax3(mx0, __uint_as_float(sr[kBlockN - 4]), __uint_as_float(sr[kBlockN - 3]));
        mx1 = fmax3(mx1, __uint_as_float(sr[kBlockN - 2]), __uint_as_float(sr[kBlockN - 1]));
        const float tile_max = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));

        if constexpr (!kPSmem) {
          // Without the shared-memory P panel nothing waits on pv_done unless a rescale happens; one non-blocking
          // test per step keeps the barrier's phases observed (compute-sanitizer synccheck: "missing wait").
          if (step > 0 && wq == 0 && lane == 0) (void)mbar_try_wait(&pv_done[t], par ^ 1);
        }
        if (step == 0) {
          m = tile_max;
        } else {
          const bool need = (tile_max - m) * scale_log2 > kRescaleThreshold;
          if (__any_sync(0xffffffffu, need)) {
            // O_t may still be accumulating PV_t(step-1): wait for it.  PV_t(step) cannot be issued
            // before we arrive on p_full below, so pv_done[t] is at most one phase ahead of us.
            UVB_PROF(pf_pv, mbar_wait(&pv_done[t], par ^ 1));
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
          if (p.key_pv_weight != nullptr) pvw = p.key_pv_weight + key0;
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          if (kEarlyS || (kPSmem && h == 1)) {
            // second half -> shared memory, row `row` of the K-major SWIZZLE_128B panel of this tile (the
            // PV of the previous step must have finished reading the panel: pv_done, long since complete).
            // (Handing the generic->async proxy fence and the arrive to a helper warp behind a named barrier
            // takes ~230 cycles per step off this warp but adds as much to the hand-off latency of the last
            // PV, which is on the dependency chain: measured slower.)
            uint32_t pk[32];
            softmax_exp<32, kKeyMod, kPoly>(sr + 64 * h, scale_log2, neg_ms, sum_a, sum_b, pk,
                                     pvw == nullptr ? nullptr : pvw + 64 * h);
            // the previous step's PV must have finished reading the panel(s) of this tile
            if (gstep + step > 0 && (!kEarlyS || h == 0)) UVB_PROF(pf_pv, mbar_wait(&pv_done[t], par ^ 1));
            uint8_t* prow = smem_p + (kEarlyS ? 2 * t + h : t) * kHalfTile + row * 128;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              *reinterpret_cast<uint4*>(prow + ((c ^ (row & 7)) << 4)) =
                  make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
            }
            fence_proxy_async_smem();
          } else {
            uint32_t pk[32];
            softmax_exp<32, kKeyMod, kPoly>(sr + 64 * h, scale_log2, neg_ms, sum_a, sum_b, pk,
                                                 pvw == nullptr ? nullptr : pvw + 64 * h);
            tmem_st_x32(tS + 32 * h, pk);
            tmem_wait_st();
            tc_fence_before();
          }
          __syncwarp();
          if (lane == 0) {                                   // one barrier per 64-key half (split-P)
            if constexpr (kCtas == 2) {
              mbar_arrive_remote(p_full_cl + (2 * t + h) * 8);
            } else {
              mbar_arrive(&p_full[2 * t + h]);
            }
          }
        }
        l += (sum_a.x + sum_a.y) + (sum_b.x + sum_b.y);
        if constexpr (kQBufs == 2) {
          if (step == 0) release_pending();
        }
      }
      if constexpr (kQBufs == 2) release_pending();     // segments without steps
      gstep += n_steps;
#ifdef UVB_FMHA_PROFILE
      pf_mark = clock64();
#endif

      // ------------------------------------- segment epilogue -------------------------------------
      mbar_wait(&o_full[t], si & 1);
      tc_fence_after();
      const bool have_o = n_steps > 0;     // otherwise TMEM holds stale data and O is zero

      if (!sg.owner) {
        // ---- partial: un-normalised O (fp32, column-major so a warp writes 128 contiguous bytes), m, l
        float* slot = p.ws + static_cast<size_t>(my_slot) * kWsSlotFloats;
        if (have_o) {
#pragma unroll
          for (int c = 0; c < kHeadDim / 32; ++c) {
            uint32_t orr[32];
            tmem_ld_x32(tO + c * 32, orr);
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 32; ++i) __stcg(slot + (c * 32 + i) * kUnitRows + ws_row, __uint_as_float(orr[i]));
          }
        }
        __stcg(slot + kWsOFloats + ws_row, m);
        __stcg(slot + kWsOFloats + kUnitRows + ws_row, have_o ? l : 0.f);
        __threadfence();
        tc_fence_before();
        named_bar_sync(1 + t, kBlockM);
        if (wq == 0 && lane == 0) {
          st_release_gpu(p.flags + 2 * my_slot + t, 1u);
          mbar_arrive(&q_empty[2 * (si % kQBufs) + t]);
        }
        continue;
      }

      // ---- owner: merge the partials of the CTAs that hold key tiles [0, a) of this unit
      float m_tot = have_o ? m : -INFINITY;
      int g_lo = sch.g;   // parts come from CTAs [g_lo, g)
      if (sg.a > 0) {
        const long long unit_lo = static_cast<long long>(sg.unit - sch.W * sch.G) * sch.n_kv;
        while (g_lo > 0 && sch.rem_lo(g_lo) > unit_lo) --g_lo;
        for (int gp = g_lo; gp < sch.g; ++gp) {
          if (sch.rem_lo(gp + 1) == sch.rem_lo(gp)) continue;     // that CTA has no remainder range
          const uint32_t* flag = p.flags + 2 * slot_of(gp) + t;
          if (ld_acquire_gpu(flag) == 0u) {
            const long long t0 = clock64();
            while (ld_acquire_gpu(flag) == 0u) {
              // other CTAs of this launch are co-resident, so a legitimate wait is microseconds; the bound only has
              // to survive a preempted / time-sliced context (~1 min of SM clocks)
              if (clock64() - t0 > 100000000000LL) __trap();
            }
          }
          const float* slot = p.ws + static_cast<size_t>(slot_of(gp)) * kWsSlotFloats;
          if (__ldcg(slot + kWsOFloats + kUnitRows + ws_row) > 0.f)
            m_tot = fmaxf(m_tot, __ldcg(slot + kWsOFloats + ws_row));
        }
      }
      const float alpha = (have_o && sg.a > 0) ? ex2_approx((m - m_tot) * scale_log2) : 1.0f;
      float l_tot = have_o ? l * alpha : 0.f;
      if (sg.a > 0) {
        for (int gp = g_lo; gp < sch.g; ++gp) {
          if (sch.rem_lo(gp + 1) == sch.rem_lo(gp)) continue;
          const float* slot = p.ws + static_cast<size_t>(slot_of(gp)) * kWsSlotFloats;
          const float lp = __ldcg(slot + kWsOFloats + kUnitRows + ws_row);
          if (lp > 0.f) l_tot += lp * ex2_approx((__ldcg(slot + kWsOFloats + ws_row) - m_tot) * scale_log2);
        }
      }
      const float inv = l_tot > 0.f ? 1.0f / l_tot : 0.f;
      const float own_scale = alpha * inv;

      uint8_t* so = smem_q + (2 * (si % kQBufs) + t) * kTileBytes;  // Q_t is dead once o_full[t] has fired
      // Peer mode: a tile whose 128 rows lie in ONE rank's token chunk is TMA-stored into that rank's
      // buffer; a tile that straddles chunks is written row by row with plain stores (each thread owns a
      // row and picks its rank) -- TMA rejects the negative start coordinate a clipped store would need.
      const int r0 = q_row0 + t * kBlockM;
      const int j_first = p.n_peers > 0 ? r0 / p.chunk : 0;
      const bool direct = p.n_peers > 0 && r0 + kBlockM > (j_first + 1) * p.chunk && j_first + 1 < p.n_peers;
      __nv_bfloat16* drow = nullptr;
      if (direct) {
        const int rg = r0 + row;
        if (rg < p.Lq) {
          const int j = rg / p.chunk;
          drow = p.o_peer_ptr[j] + ((static_cast<size_t>(batch) * p.chunk + (rg - j * p.chunk)) * p.o_heads +
                                    p.o_head_off + head) * kHeadDim;
        }
      }
      const float* bias = nullptr;
      if constexpr (kKeyMod) {
        if (p.out_bias != nullptr && l_tot > 0.f) bias = p.out_bias + head * kHeadDim;
      }
#pragma unroll
      for (int c = 0; c < kHeadDim / 32; ++c) {
        uint32_t orr[32];
        if (have_o) {
          tmem_ld_x32(tO + c * 32, orr);
          tmem_wait_ld();
#pragma unroll
          for (int i = 0; i < 32; ++i) orr[i] = __float_as_uint(__uint_as_float(orr[i]) * own_scale);
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) orr[i] = 0u;
        }
        if (sg.a > 0) {
          for (int gp = g_lo; gp < sch.g; ++gp) {
            if (sch.rem_lo(gp + 1) == sch.rem_lo(gp)) continue;
            const float* slot = p.ws + static_cast<size_t>(slot_of(gp)) * kWsSlotFloats;
            const float lp = __ldcg(slot + kWsOFloats + kUnitRows + ws_row);
            if (lp > 0.f) {
              const float w = ex2_approx((__ldcg(slot + kWsOFloats + ws_row) - m_tot) * scale_log2) * inv;
#pragma unroll
              for (int i = 0; i < 32; ++i) {
                orr[i] = __float_as_uint(fmaf(__ldcg(slot + (c * 32 + i) * kUnitRows + ws_row), w,
                                              __uint_as_float(orr[i])));
              }
            }
          }
        }
        uint32_t ob[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float a = __uint_as_float(orr[2 * i]);
          float b = __uint_as_float(orr[2 * i + 1]);
          if constexpr (kKeyMod) {
            if (bias != nullptr) {
              a += __ldg(bias + c * 32 + 2 * i);
              b += __ldg(bias + c * 32 + 2 * i + 1);
            }
          }
          ob[i] = pack_bf16x2(a, b);
        }
        if (direct) {
          if (drow != nullptr) {
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4) {
              *reinterpret_cast<uint4*>(drow + c * 32 + q4 * 8) =
                  make_uint4(ob[4 * q4], ob[4 * q4 + 1], ob[4 * q4 + 2], ob[4 * q4 + 3]);
            }
          }
        } else {
          // 32 columns = 4 x 16-byte chunks of panel (c/2); SWIZZLE_128B: chunk ^= row % 8
          uint8_t* prow = so + (c >> 1) * kHalfTile + row * 128;
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            const int chunk = (c & 1) * 4 + q4;
            *reinterpret_cast<uint4*>(prow + ((chunk ^ (row & 7)) << 4)) =
                make_uint4(ob[4 * q4], ob[4 * q4 + 1], ob[4 * q4 + 2], ob[4 * q4 + 3]);
          }
        }
      }
      fence_proxy_async_smem();
      tc_fence_before();
      named_bar_sync(1 + t, kBlockM);
      if (wq == 0 && lane == 0) {
        if (r0 >= p.Lq) {
          // the second tile of the last query block may lie entirely past Lq: nothing to store
        } else if (p.n_peers == 0) {
          tma_store_4d(&p.tm_o, so, 0, r0, head, batch);
          tma_store_4d(&p.tm_o, so + kHalfTile, 64, r0, head, batch);
        } else if (!direct) {
          // static indices only: the descriptor address must stay a plain param-space address (a
          // dynamically indexed copy would live in local memory, which TMA cannot read).  Rows past the
          // end of the last rank's chunk (Lq not a multiple of 128) are out of bounds and dropped.
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            if (j == j_first) {
              tma_store_4d(&p.tm_o_peer[j], so, 0, r0 - j * p.chunk, p.o_head_off + head, batch);
              tma_store_4d(&p.tm_o_peer[j], so + kHalfTile, 64, r0 - j * p.chunk, p.o_head_off + head, batch);
            }
          }
        }
        tma_store_commit();
        if (sg.a > 0) {   // every thread of the group has read the partials: hand the slots back
          for (int gp = g_lo; gp < sch.g; ++gp) {
            if (sch.rem_lo(gp + 1) != sch.rem_lo(gp)) st_release_gpu(p.flags + 2 * slot_of(gp) + t, 0u);
          }
        }
        if constexpr (kQBufs == 2) {
          pending_qb = si % kQBufs;
        } else {
          tma_store_wait_read0();          // the staging buffer may be overwritten by the next Q_t
          mbar_arrive(&q_empty[si % kQBufs * 2 + t]);
        }
        stored = true;
      }
    }
    if constexpr (kQBufs == 2) release_pending();
    if (stored) tma_store_wait0();
#ifdef UVB_FMHA_PROFILE
    if (p.prof != nullptr && wq == 0 && lane == 0) {
      unsigned long long* pr = p.prof + blockIdx.x * 16 + 3 * t;
      pr[0] = clock64() - pf_start;
      pr[1] = pf_s;
      pr[2] = pf_pv;
      p.prof[blockIdx.x * 16 + 12 + 2 * t] = pf_setup;
      p.prof[blockIdx.x * 16 + 13 + 2 * t] = pf_epi + (pf_mark != 0 ? clock64() - pf_mark : 0);
    }
#endif
    if (p.timeline != nullptr && threadIdx.x == 0)
      p.timeline[blockIdx.x * 32 + 1 + min(sch.n_seg, 30)] = globaltimer_ns();
  }

  // ------------------------------------------ teardown ------------------------------------------
  tc_fence_before();
  if constexpr (kCtas == 2) {
    cluster_sync_all();     // the leader's MMAs read the peer's shared memory / TMEM until the very end
  } else {
    __syncthreads();
  }
  if (warp == 8) {
    __syncwarp();
    if constexpr (kCtas == 2) {
      tmem_dealloc_pair(tmem_base, 512);
    } else {
      tmem_dealloc(tmem_base, 512);
    }
  }
}

}  // namespace uvb
