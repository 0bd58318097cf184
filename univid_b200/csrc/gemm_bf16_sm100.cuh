// y[M, N] = act(x[M, K] . w[N, K]^T + bias[N]) on sm_100a: bf16 operands, fp32 accumulation in TMEM, bf16 result.
// This is nn.Linear under bf16 autocast as the Wan DiT block uses it around the attention hot path (reference:
// models/wan/utils/modules/model.py:119-122 q/k/v/o, :212-214 ffn) with nn.GELU(approximate='tanh') (:213)
// optionally fused into the epilogue.
//
// Persistent, warp-specialised, one CTA per SM (kCtas == 1) or one CTA PAIR per TPC (kCtas == 2, cta_group::2):
//   warp 0      TMA producer: A panel [128 rows x 64 k] and B panel [256/kCtas rows x 64 k] per ring stage
//   warp 1      tcgen05.mma issuer (one elected lane; in a pair only the leader CTA issues) + TMEM allocator
//   warps 2-5   epilogue: TMEM -> registers -> (+bias, round, GELU) -> swizzled smem panel -> TMA store
// Output tile per CTA is 128 x kBN, kBN = 256 or 192 (a pair computes 256 x kBN: each CTA holds its 128 rows of A
// and HALF of the B tile, the tensor cores of both SMs read both halves -- half the shared-memory operand traffic
// per flop).  The host picks kBN per problem so that the number of tiles fills whole waves of the persistent grid
// (N = 1536 on 74 pairs: 10.4 waves of 256-wide tiles, 13.8 of 192-wide ones).  TMEM holds two 128 x kBN fp32
// accumulators so the epilogue of tile i overlaps the MMAs of tile i+1.  Shared memory: kStages x (A 16 KiB +
// B kBN/kCtas x 128 B) ring, two 16 KiB output staging panels.
// Problems whose 128 x 64 tiles fit in one wave of single CTAs (a few hundred rows) run as <1, 64>: latency-bound.
// Tiles are dealt round-robin to the persistent CTAs in an order that walks `group_n` column tiles for each row
// tile before moving down; the host sizes the group so that its slab of B (group_n x kBN x K) stays in L2 while
// A streams past once per group.  Ragged M / N / K edges are handled by TMA (zero fill on load, clipping on store).
//
// Rounding points follow the reference chain Linear(bf16 autocast) -> GELU: the biased accumulator is rounded to
// bf16 first (the Linear's output), GELU is evaluated in fp32 on that value and rounded again.
#pragma once
#include "ptx.cuh"

namespace uvb {

constexpr int kGemmBM = 128;            // rows per CTA tile (UMMA M per CTA)
constexpr int kGemmBK = 64;             // k per ring stage = one 128-byte swizzle panel
constexpr int kGemmThreads = 192;
constexpr int kGemmPanelBytes = kGemmBM * 128;   // 16 KiB: 128 rows x 64 bf16

enum GemmAct { kActNone = 0, kActGeluTanh = 1 };

template <int kCtas, int kBN>
struct GemmSmem {
  static_assert(kBN % 64 == 0 && kBN <= 256 && (kBN / kCtas) % 16 == 0, "tile width");
  static constexpr int kABytes = kGemmPanelBytes;
  static constexpr int kBBytes = (kBN / kCtas) * 128;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kFixedBytes = 2 * kGemmPanelBytes + 256 * 4 + 256 + 1024;   // staging, bias, barriers, slack
  static constexpr int kStages = (232448 - kFixedBytes) / kStageBytes < 8 ? (232448 - kFixedBytes) / kStageBytes : 8;
  static constexpr int kCOff = kStages * kStageBytes;            // 2 output staging panels
  static constexpr int kBiasOff = kCOff + 2 * kGemmPanelBytes;   // 256 fp32: bias of the current column tile
  static constexpr int kBarOff = kBiasOff + 256 * 4;
  // barriers: full[S] empty[S] tmem_full[2] tmem_empty[2] + tmem ptr
  static constexpr int kNumBars = 2 * kStages + 4;
  static constexpr int kBytes = kBarOff + kNumBars * 8 + 16;
  static constexpr int kDynBytes = kBytes + 1024;
};

struct GemmParams {
  CUtensorMap tm_a;   // x  : dims (K, M), box (64, 128),        SWIZZLE_128B
  CUtensorMap tm_b;   // w  : dims (K, N), box (64, kBN/kCtas),  SWIZZLE_128B
  CUtensorMap tm_c;   // y  : dims (N, M), box (64, 128),        SWIZZLE_128B
  const float* bias;  // [N] or nullptr
  int M, N, K;
  int act;
  int n_m;            // row tiles of (128 * kCtas) rows
  int n_n;            // column tiles of kBN
  int group_n;        // column tiles walked per row tile (rasterisation)
  // Column-group scatter (Ulysses: the v projection stored straight into the peers' exchange buffers, the
  // chunk(...).contiguous() pack of distributed/util.py:27 fused into the GEMM epilogue): when n_peers > 0 the
  // output columns [j * peer_cols, (j + 1) * peer_cols) go to tm_c_peer[j] (a [M, peer_cols] matrix in rank j's
  // buffer, mapped over NVLink) instead of tm_c.  peer_cols is a multiple of 64, so a staging panel never straddles.
  int n_peers;
  int peer_cols;
  CUtensorMap tm_c_peer[8];
};

// ---- 2-D TMA and cluster helpers local to the GEMM ----
__device__ __forceinline__ void tma_load_2d_hint(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0,
                                                 int c1, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint "
      "[%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "l"(policy)
      : "memory");
}
// CTA-pair variant: the data lands in THIS CTA's shared memory, the transaction bytes are counted on the
// mbarrier at cluster address `bar_cluster` (the leader CTA's barrier)
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* map, uint32_t bar_cluster,
                                                 int c0, int c1, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint "
      "[%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(smem_dst)), "l"(map), "r"(bar_cluster), "r"(c0), "r"(c1), "l"(policy)
      : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(map), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
template <int kCtas>
__device__ __forceinline__ void gemm_tmem_alloc(uint32_t* smem_dst) {
  if constexpr (kCtas == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
                 "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  } else {
    tmem_alloc(smem_dst, 512);
    tmem_relinquish();
  }
}
template <int kCtas>
__device__ __forceinline__ void gemm_tmem_dealloc(uint32_t taddr) {
  if constexpr (kCtas == 2) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(512) : "memory");
  } else {
    tmem_dealloc(taddr, 512);
  }
}
template <int kCtas>
__device__ __forceinline__ void gemm_umma(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
  if constexpr (kCtas == 2) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    umma_ss(d_tmem, a_desc, b_desc, idesc, accumulate);
  }
}
// arrive on `bar` (same smem offset in every CTA of the pair) once all MMAs issued so far have completed
template <int kCtas>
__device__ __forceinline__ void gemm_commit(uint64_t* bar) {
  if constexpr (kCtas == 2) {
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
        ::"r"(smem_u32(bar)), "h"(static_cast<uint16_t>(3))
        : "memory");
  } else {
    tc_commit(bar);
  }
}

// tile index -> (row tile, column tile): groups of group_n column tiles, row-major inside a group
__device__ __forceinline__ void gemm_tile_coords(int t, int n_m, int n_n, int group_n, int& tm, int& tn) {
  const int per_group = n_m * group_n;
  const int g = t / per_group;
  const int r = t - g * per_group;
  const int w = min(group_n, n_n - g * group_n);
  tm = r / w;
  tn = g * group_n + (r - tm * w);
}

__device__ __forceinline__ float gelu_tanh_f32(float u) {
  // 0.5 u (1 + tanh(c (u + 0.044715 u^3))) == u / (1 + exp(-2 c (u + 0.044715 u^3)))
  const float inner = u * fmaf(0.044715f * u, u, 1.0f);
  const float e = ex2_approx(inner * (-2.0f * 0.7978845608028654f * 1.4426950408889634f));
  return __fdividef(u, 1.0f + e);
}

template <int kCtas, int kBN>
__global__ void __launch_bounds__(kGemmThreads, 1) gemm_bf16_kernel(const __grid_constant__ GemmParams p) {
  using SM = GemmSmem<kCtas, kBN>;
  constexpr int kStages = SM::kStages;
  static_assert(SM::kDynBytes <= 232448, "shared memory budget (227 KiB)");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  asm volatile("" : "+l"(smem));   // keep the aligned base opaque (see fmha_fwd_sm100.cuh)
  uint8_t* smem_c = smem + SM::kCOff;
  float* smem_bias = reinterpret_cast<float*>(smem + SM::kBiasOff);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SM::kBarOff);
  uint64_t* full = bars;                       // [kStages]  TMA -> MMA (pair: only the leader's are used)
  uint64_t* empty = bars + kStages;            // [kStages]  MMA -> TMA (every CTA its own)
  uint64_t* tmem_full = empty + kStages;       // [2]        MMA -> epilogue (every CTA its own)
  uint64_t* tmem_empty = tmem_full + 2;        // [2]        epilogue -> MMA (pair: the leader's collect both CTAs)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t cta_rank = kCtas == 2 ? cluster_ctarank() : 0u;
  const bool leader = cta_rank == 0;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 4 * kCtas);     // one arrive per epilogue warp of every CTA of the pair
    }
    fence_mbar_init();
  }
  if (warp == 1) gemm_tmem_alloc<kCtas>(tmem_ptr);
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tm_a);
    tma_prefetch_desc(&p.tm_b);
    if (p.n_peers == 0) tma_prefetch_desc(&p.tm_c);
  }
  tc_fence_before();
  if constexpr (kCtas == 2) {
    cluster_sync_all();
    __syncthreads();        // subsumed by the cluster barrier; spelled out for compute-sanitizer's racecheck
  } else {
    __syncthreads();
  }
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr, 0);

  const int n_tiles = p.n_m * p.n_n;
  const int n_workers = static_cast<int>(gridDim.x) / kCtas;      // CTAs (or pairs) walking the tile list
  const int worker = static_cast<int>(blockIdx.x) / kCtas;
  const int n_kb = (p.K + kGemmBK - 1) / kGemmBK;

  if (warp == 0) {
    // ========================================= TMA producer =========================================
    int ring = 0;
    const uint32_t full0 = kCtas == 2 ? map_to_cta(&full[0], 0) : 0u;   // leader's full[0], cluster address
    for (int t = worker; t < n_tiles; t += n_workers) {
      int tm, tn;
      gemm_tile_coords(t, p.n_m, p.n_n, p.group_n, tm, tn);
      const int row0 = (tm * kCtas + static_cast<int>(cta_rank)) * kGemmBM;
      const int col0 = tn * kBN + static_cast<int>(cta_rank) * (kBN / kCtas);
      for (int kb = 0; kb < n_kb; ++kb, ++ring) {
        const int stage = ring % kStages;
        mbar_wait(&empty[stage], ((ring / kStages) & 1) ^ 1);
        uint8_t* sa = smem + stage * SM::kStageBytes;
        uint8_t* sb = sa + SM::kABytes;
        if (elect_one()) {
          if constexpr (kCtas == 2) {
            if (leader) mbar_arrive_expect_tx(&full[stage], 2 * SM::kStageBytes);
            const uint32_t fb = full0 + stage * 8;
            tma_load_2d_pair(sa, &p.tm_a, fb, kb * kGemmBK, row0, kEvictNormal);
            tma_load_2d_pair(sb, &p.tm_b, fb, kb * kGemmBK, col0, kEvictLast);
          } else {
            mbar_arrive_expect_tx(&full[stage], SM::kStageBytes);
            tma_load_2d_hint(sa, &p.tm_a, &full[stage], kb * kGemmBK, row0, kEvictNormal);
            tma_load_2d_hint(sb, &p.tm_b, &full[stage], kb * kGemmBK, col0, kEvictLast);
          }
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ========================================= MMA issuer ===========================================
    if (leader) {
      constexpr uint32_t idesc = umma_idesc_bf16(kGemmBM * kCtas, kBN, 0, 0);
      const uint64_t a_desc = umma_desc_sw128(smem_u32(smem), 16, 1024);
      const uint64_t b_desc = umma_desc_sw128(smem_u32(smem + SM::kABytes), 16, 1024);
      int ring = 0, it = 0;
      for (int t = worker; t < n_tiles; t += n_workers, ++it) {
        const int buf = it & 1;
        mbar_wait(&tmem_empty[buf], ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d = tmem_base + buf * 256;
        for (int kb = 0; kb < n_kb; ++kb, ++ring) {
          const int stage = ring % kStages;
          mbar_wait(&full[stage], (ring / kStages) & 1);
          tc_fence_after();
          if (elect_one()) {
            const uint64_t so = static_cast<uint64_t>(stage * SM::kStageBytes) >> 4;
#pragma unroll
            for (int k4 = 0; k4 < kGemmBK / 16; ++k4) {
              gemm_umma<kCtas>(d, a_desc + so + ((k4 * 32) >> 4), b_desc + so + ((k4 * 32) >> 4), idesc,
                               (kb > 0 || k4 > 0) ? 1u : 0u);
            }
            gemm_commit<kCtas>(&empty[stage]);
            if (kb == n_kb - 1) gemm_commit<kCtas>(&tmem_full[buf]);
          }
          __syncwarp();
        }
      }
    }
  } else {
    // ========================================= epilogue =============================================
    const int wq = warp & 3;                       // TMEM lane quadrant this warp may access
    const int row = wq * 32 + lane;                // row inside the CTA tile
    const int et = threadIdx.x - 64;               // 0..127
    const uint32_t lane_addr = static_cast<uint32_t>(wq * 32) << 16;
    const uint32_t te0 = kCtas == 2 ? map_to_cta(&tmem_empty[0], 0) : 0u;
    const bool gelu = p.act == kActGeluTanh;
    int it = 0;
    int slot = 0;                                  // staging panels used so far (they alternate)
    for (int t = worker; t < n_tiles; t += n_workers, ++it) {
      int tm, tn;
      gemm_tile_coords(t, p.n_m, p.n_n, p.group_n, tm, tn);
      const int row0 = (tm * kCtas + static_cast<int>(cta_rank)) * kGemmBM;
      const int col0 = tn * kBN;
      const int buf = it & 1;
      // every epilogue thread passed the last barrier of the previous tile after its last bias read
      float* sbias = smem_bias;
      for (int c = et; c < kBN; c += 128) {
        const int n = col0 + c;
        sbias[c] = (p.bias != nullptr && n < p.N) ? __ldg(p.bias + n) : 0.f;
      }
      mbar_wait(&tmem_full[buf], (it >> 1) & 1);
      tc_fence_after();
      const uint32_t tacc = tmem_base + lane_addr + buf * 256;
#pragma unroll 1
      for (int j = 0; j < kBN / 64; ++j, ++slot) {   // 64 output columns = one staging panel
        uint8_t* panel = smem_c + (slot & 1) * kGemmPanelBytes;
        // the TMA store that last read this panel (two slots ago) must be done reading: thread 0 waits (every slot
        // commits exactly one bulk group, empty or not), the barrier publishes it (it also orders the bias writes
        // of this tile before their first use)
        if (et == 0) tma_store_wait_read1();
        named_bar_sync(1, 128);
        if (col0 + j * 64 < p.N) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            uint32_t acc[32];
            tmem_ld_x32(tacc + j * 64 + h * 32, acc);
            tmem_wait_ld();
            uint32_t ob[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const float2 b2 = *reinterpret_cast<const float2*>(sbias + j * 64 + h * 32 + 2 * i);
              float v0 = __uint_as_float(acc[2 * i]) + b2.x;
              float v1 = __uint_as_float(acc[2 * i + 1]) + b2.y;
              uint32_t pk = pack_bf16x2(v0, v1);
              if (gelu) {
                v0 = gelu_tanh_f32(__uint_as_float(pk << 16));
                v1 = gelu_tanh_f32(__uint_as_float(pk & 0xffff0000u));
                pk = pack_bf16x2(v0, v1);
              }
              ob[i] = pk;
            }
            uint8_t* prow = panel + row * 128;
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4) {
              const int chunk = h * 4 + q4;
              *reinterpret_cast<uint4*>(prow + ((chunk ^ (row & 7)) << 4)) =
                  make_uint4(ob[4 * q4], ob[4 * q4 + 1], ob[4 * q4 + 2], ob[4 * q4 + 3]);
            }
          }
        }
        if (j == kBN / 64 - 1) {
          // accumulator fully read: hand the TMEM buffer back to the MMA issuer
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if constexpr (kCtas == 2) {
              mbar_arrive_cluster(te0 + buf * 8);
            } else {
              mbar_arrive(&tmem_empty[buf]);
            }
          }
        }
        fence_proxy_async_smem();
        named_bar_sync(2, 128);
        if (et == 0) {
          if (col0 + j * 64 < p.N && row0 < p.M) {
            if (p.n_peers == 0) {
              tma_store_2d(&p.tm_c, panel, col0 + j * 64, row0);
            } else {
              // static indices only: the descriptor must stay a plain param-space address
              const int c = col0 + j * 64;
              const int peer = c / p.peer_cols;
#pragma unroll
              for (int q = 0; q < 8; ++q) {
                if (q == peer) tma_store_2d(&p.tm_c_peer[q], panel, c - q * p.peer_cols, row0);
              }
            }
          }
          tma_store_commit();
        }
      }
    }
    if (et == 0) tma_store_wait0();
  }

  // ------------------------------------------ teardown ------------------------------------------
  tc_fence_before();
  if constexpr (kCtas == 2) {
    cluster_sync_all();
  } else {
    __syncthreads();
  }
  if (warp == 1) {
    __syncwarp();
    gemm_tmem_dealloc<kCtas>(tmem_base);
  }
}

}  // namespace uvb
