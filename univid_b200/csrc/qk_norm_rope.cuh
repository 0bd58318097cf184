// Fused q/k WanRMSNorm (over the full model width) + 3-D RoPE prologue, HBM-bound.
// Reference math: WanRMSNorm.forward  models/wan/utils/modules/model.py:77-85
//                 rope_apply          models/wan/utils/modules/model.py:38-66
//                 SP rope_apply       models/wan/distributed/sequence_parallel.py:23-61 (tok_offset)
// Rounding points mirror the reference under bf16 autocast: the normalised value is rounded to the
// input dtype (`.type_as(x)`), multiplied by the fp32 weight, rotated, and rounded once to bf16
// (the cast the reference performs at flash_attention() entry, attention.py:59-83).
//
// One warp owns one token row of q or of k.  Lane l holds the 16-byte vectors v = l + 32*i of the row, so every
// vector of a lane has the same offset inside its head (d0 = 8*(l%16)): the lane needs just four
// (cos, sin) pairs per token, shared by all heads and by q and k.  All loads of a row are issued
// before the first use (VPL independent 128-bit loads per lane), the sum of squares is reduced with
// warp shuffles, and the row is written back with 128-bit stores.
#pragma once
#include "ptx.cuh"

namespace uvb {

constexpr int kMaxBatchGrid = 8;
constexpr int kMaxPeers = 8;     // ranks of one NVSwitch box

struct NormRopeParams {
  const void* q_in;        // [B, L, dim]  (InT)
  const void* k_in;        // [B, L, dim]  or nullptr
  const float* wq;         // [dim]
  const float* wk;         // [dim]
  const float2* cos_sin;   // [1024][64] (cos, sin) of the concatenated (f | h | w) bands, or nullptr
  const float* row_scale;  // [L] or nullptr: x <- row_scale[l] * x + pre_bias (applied to k only)
  const float* pre_bias;   // [dim] or nullptr
  __nv_bfloat16* q_out;
  __nv_bfloat16* k_out;
  int B, L, N;             // dim = N * 128
  int grid[kMaxBatchGrid][3];   // (f, h, w) per sample
  int tok_offset;          // global index of local token 0 (sequence-parallel shard)
  float eps;
  // output addressing: elem(b, l, n, d) = b*out_sb + l*out_sl + (n / hpg)*out_sg + (n % hpg)*128 + d
  int hpg;                 // heads per group (N for plain [B,L,N,128])
  long long out_sb, out_sl, out_sg;
  // Ulysses peer stores: when n_peers > 0 head group j is written through {q,k}_peer[j] (a pointer into
  // rank j's exchange buffer, mapped over NVLink) instead of {q,k}_out + j*out_sg
  int n_peers;
  __nv_bfloat16* q_peer[kMaxPeers];
  __nv_bfloat16* k_peer[kMaxPeers];
};

template <typename InT>
struct RowVec;  // 8 consecutive elements

template <>
struct RowVec<__nv_bfloat16> {
  uint4 raw;
  __device__ __forceinline__ void load(const __nv_bfloat16* p) {
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(raw.x), "=r"(raw.y), "=r"(raw.z), "=r"(raw.w)
                 : "l"(p));
  }
  __device__ __forceinline__ void unpack(float* f) const {
    const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      f[2 * i] = __uint_as_float(w[i] << 16);
      f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
  }
  // `.type_as(x)`: round the normalised values to the input dtype.  Done two at a time through the
  // packed F2FP convert + shifts: the scalar F2F.BF16.F32 runs on the quarter-rate XU pipe and made
  // the kernel issue-bound (profiles/r01: 34 % XU, 46 % DRAM).
  static __device__ __forceinline__ void round_in2(float& a, float& b) {
    const uint32_t pk = pack_bf16x2(a, b);
    a = __uint_as_float(pk << 16);
    b = __uint_as_float(pk & 0xffff0000u);
  }
};

template <>
struct RowVec<float> {
  float4 a, b;
  __device__ __forceinline__ void load(const float* p) {
    a = __ldg(reinterpret_cast<const float4*>(p));
    b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  }
  __device__ __forceinline__ void unpack(float* f) const {
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w;
    f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
  }
  static __device__ __forceinline__ void round_in2(float&, float&) {}
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Normalise + rotate one row and store it.  The row stays PACKED in registers (VPL x 16 bytes per
// lane for bf16) and is unpacked twice -- once for the sum of squares, once for the output -- which
// keeps the register footprint small enough for 24-32 resident warps per SM.
//
// WPR warps cooperate on one row (lane index `lane` in [0, 32*WPR)); their partial sums of squares
// meet in shared memory behind a named barrier private to the row.  kPre selects the affine
// pre-map x <- rscale * x + pre_bias (text-weighted context rows).
template <typename InT, int VPL, int WPR>
__device__ __forceinline__ void load_row(RowVec<InT> (&v)[VPL], const InT* __restrict__ in, int lane) {
#pragma unroll
  for (int i = 0; i < VPL; ++i) v[i].load(in + (lane + 32 * WPR * i) * 8);
}

template <typename InT, int VPL, int WPR, bool kPre>
__device__ __forceinline__ void norm_rope_row(RowVec<InT> (&v)[VPL], const float* __restrict__ w,
                                              __nv_bfloat16* __restrict__ out_row,
                                              __nv_bfloat16* const* peers, long long out_off, int hpg,
                                              long long out_sg, int dim, float eps, bool rotate,
                                              const float (&cs)[8], float rscale,
                                              const float* __restrict__ pre_bias, int lane,
                                              float* red, int bar_id) {
  constexpr int kStride = 32 * WPR;

  auto fetch = [&](int i, float* x) {
    v[i].unpack(x);
    if constexpr (kPre) {
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(pre_bias + (lane + kStride * i) * 8));
      const float4 b1 = __ldg(reinterpret_cast<const float4*>(pre_bias + (lane + kStride * i) * 8) + 1);
      const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int e = 0; e < 8; e += 2) {
        x[e] = fmaf(rscale, x[e], bb[e]);
        x[e + 1] = fmaf(rscale, x[e + 1], bb[e + 1]);
        RowVec<InT>::round_in2(x[e], x[e + 1]);
      }
    }
  };

  // packed f32x2 arithmetic (FFMA2 / FMUL2): the kernel is co-limited by instruction issue
  float2 ss2 = make_float2(0.f, 0.f);
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    float x[8];
    fetch(i, x);
#pragma unroll
    for (int e = 0; e < 8; e += 2) {
      const float2 xx = make_float2(x[e], x[e + 1]);
      ss2 = ffma2(xx, xx, ss2);
    }
  }
  float ss = warp_sum(ss2.x + ss2.y);
  if constexpr (WPR > 1) {
    if ((lane & 31) == 0) red[lane >> 5] = ss;
    named_bar_sync(bar_id, kStride);
    ss = 0.f;
#pragma unroll
    for (int i = 0; i < WPR; ++i) ss += red[i];
  }
  // w == nullptr: qk_norm disabled (nn.Identity, model.py:123-124) -> rotation only
  const bool normed = w != nullptr;
  const float rinv = normed ? rsqrtf(ss / static_cast<float>(dim) + eps) : 1.0f;
  const bool flat = hpg * 128 == dim && peers == nullptr;   // one head group: output rows are dense [N, 128]

#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int vec = lane + kStride * i;
    float y[8];
    fetch(i, y);
    if (normed) {
      const float4 w0 = __ldg(reinterpret_cast<const float4*>(w + vec * 8));
      const float4 w1 = __ldg(reinterpret_cast<const float4*>(w + vec * 8) + 1);
      const float ww[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
      const float2 r2 = make_float2(rinv, rinv);
#pragma unroll
      for (int e = 0; e < 8; e += 2) {
        float2 t = fmul2(make_float2(y[e], y[e + 1]), r2);
        RowVec<InT>::round_in2(t.x, t.y);
        t = fmul2(t, make_float2(ww[e], ww[e + 1]));
        y[e] = t.x;
        y[e + 1] = t.y;
      }
    }
    uint32_t o[4];
#pragma unroll
    for (int pr = 0; pr < 4; ++pr) {
      float a = y[2 * pr], b = y[2 * pr + 1];
      if (rotate) {
        const float c = cs[2 * pr], s = cs[2 * pr + 1];
        const float ra = a * c - b * s;
        const float rb = fmaf(a, s, b * c);
        a = ra;
        b = rb;
      }
      o[pr] = pack_bf16x2(a, b);
    }
    __nv_bfloat16* dst;
    if (flat) {
      dst = out_row + vec * 8;                 // [.., N, 128] rows are dense
    } else {
      const int n = vec >> 4;                  // head index
      const int d0 = (vec & 15) * 8;           // offset inside the head
      if (peers != nullptr) {
        dst = peers[n / hpg] + out_off + (n % hpg) * 128 + d0;      // rank (n / hpg)'s buffer, over NVLink
      } else {
        dst = out_row + static_cast<long long>(n / hpg) * out_sg + (n % hpg) * 128 + d0;
      }
    }
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "r"(o[0]),
                 "r"(o[1]), "r"(o[2]), "r"(o[3])
                 : "memory");
  }
}

// Generic-width variant (dim = 256 * nvec_per_lane not known at compile time): two passes over the
// row, the second one served by L1/L2.
template <typename InT>
__device__ __forceinline__ void norm_rope_row_generic(const InT* __restrict__ in,
                                                      const float* __restrict__ w,
                                                      __nv_bfloat16* __restrict__ out_row,
                                                      __nv_bfloat16* const* peers, long long out_off,
                                                      int hpg, long long out_sg, int dim, float eps,
                                                      bool rotate, const float (&cs)[8],
                                                      float rscale,
                                                      const float* __restrict__ pre_bias, int lane) {
  const int nvec = dim / 8;
  float ss = 0.f;
  for (int vec = lane; vec < nvec; vec += 32) {
    RowVec<InT> v;
    v.load(in + vec * 8);
    float x[8];
    v.unpack(x);
    if (pre_bias != nullptr) {
#pragma unroll
      for (int e = 0; e < 8; e += 2) {
        x[e] = fmaf(rscale, x[e], pre_bias[vec * 8 + e]);
        x[e + 1] = fmaf(rscale, x[e + 1], pre_bias[vec * 8 + e + 1]);
        RowVec<InT>::round_in2(x[e], x[e + 1]);
      }
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      ss = fmaf(x[e], x[e], ss);
    }
  }
  ss = warp_sum(ss);
  const bool normed = w != nullptr;
  const float rinv = normed ? rsqrtf(ss / static_cast<float>(dim) + eps) : 1.0f;
  for (int vec = lane; vec < nvec; vec += 32) {
    RowVec<InT> v;
    v.load(in + vec * 8);
    float x[8];
    v.unpack(x);
    uint32_t o[4];
#pragma unroll
    for (int pr = 0; pr < 4; ++pr) {
      float a = x[2 * pr], b = x[2 * pr + 1];
      if (pre_bias != nullptr) {
        a = fmaf(rscale, a, pre_bias[vec * 8 + 2 * pr]);
        b = fmaf(rscale, b, pre_bias[vec * 8 + 2 * pr + 1]);
        RowVec<InT>::round_in2(a, b);
      }
      if (normed) {
        a *= rinv;
        b *= rinv;
        RowVec<InT>::round_in2(a, b);
        a *= w[vec * 8 + 2 * pr];
        b *= w[vec * 8 + 2 * pr + 1];
      }
      if (rotate) {
        const float c = cs[2 * pr], s = cs[2 * pr + 1];
        const float ra = a * c - b * s;
        const float rb = fmaf(a, s, b * c);
        a = ra;
        b = rb;
      }
      o[pr] = pack_bf16x2(a, b);
    }
    const int n = vec >> 4;
    const int d0 = (vec & 15) * 8;
    __nv_bfloat16* dst = peers != nullptr
                             ? peers[n / hpg] + out_off + (n % hpg) * 128 + d0
                             : out_row + static_cast<long long>(n / hpg) * out_sg + (n % hpg) * 128 + d0;
    *reinterpret_cast<uint4*>(dst) = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

constexpr int kNormRopeWarps = 8;

// One group of WPR warps per (token row, tensor): even groups take q, odd groups k, so the groups of a
// token sit next to each other and share the token's (cos, sin) lines in L1.
// VPL * WPR = dim / 256 (16-byte vectors per lane); VPL == 0 selects the generic two-pass path.
template <typename InT, int VPL, int WPR, bool kPeers>
__global__ void __launch_bounds__(kNormRopeWarps * 32, sizeof(InT) == 2 ? 3 : 2)
qk_norm_rope_kernel(const __grid_constant__ NormRopeParams p) {
  __shared__ float red[kNormRopeWarps];
  const int warp = threadIdx.x >> 5;
  const int group = warp / WPR;                       // row group inside the CTA
  const int lane = threadIdx.x - group * (32 * WPR);  // lane inside the group
  const long long unit = static_cast<long long>(blockIdx.x) * (kNormRopeWarps / WPR) + group;
  const long long row = unit >> 1;
  const bool is_k = (unit & 1) != 0;
  if (row >= static_cast<long long>(p.B) * p.L) return;
  const InT* src = static_cast<const InT*>(is_k ? p.k_in : p.q_in);
  if (src == nullptr) return;
  const int b = static_cast<int>(row / p.L);
  const int l = static_cast<int>(row % p.L);
  const int dim = p.N * 128;

  // (cos, sin) for pairs jj = 4*(lane%16) .. +3 of this token
  float cs[8];
  bool rotate = false;
  if (p.cos_sin != nullptr) {
    const int gb = b < kMaxBatchGrid ? b : kMaxBatchGrid - 1;
    const int gh = p.grid[gb][1], gw = p.grid[gb][2];
    const int tok = p.tok_offset + l;
    rotate = tok < p.grid[gb][0] * gh * gw;   // padding tokens pass through unrotated (model.py:62)
    if (rotate) {
      const int pf = tok / (gh * gw), ph = (tok / gw) % gh, pw = tok % gw;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int jj = 4 * (lane & 15) + i;
        const int pos = jj < 22 ? pf : (jj < 43 ? ph : pw);   // bands: 22 | 21 | 21 (model.py:43)
        const float2 v = __ldg(p.cos_sin + pos * 64 + jj);
        cs[2 * i] = v.x;
        cs[2 * i + 1] = v.y;
      }
    }
  }
  if (!rotate) {
#pragma unroll
    for (int i = 0; i < 8; ++i) cs[i] = 0.f;
  }

  const long long in_off = row * dim;
  const long long out_off = static_cast<long long>(b) * p.out_sb + static_cast<long long>(l) * p.out_sl;
  const float* w = is_k ? p.wk : p.wq;
  __nv_bfloat16* dst = (is_k ? p.k_out : p.q_out) + out_off;
  __nv_bfloat16* const* peers = kPeers ? (is_k ? p.k_peer : p.q_peer) : nullptr;
  const float* pre_bias = is_k ? p.pre_bias : nullptr;
  const float rscale = (is_k && p.row_scale != nullptr) ? p.row_scale[l] : 1.0f;
  if constexpr (VPL > 0) {
    RowVec<InT> v[VPL];
    load_row<InT, VPL, WPR>(v, src + in_off, lane);
    if (pre_bias != nullptr) {
      norm_rope_row<InT, VPL, WPR, true>(v, w, dst, peers, out_off, p.hpg, p.out_sg, dim, p.eps,
                                         rotate, cs, rscale, pre_bias, lane, red + group * WPR, 1 + group);
    } else {
      norm_rope_row<InT, VPL, WPR, false>(v, w, dst, peers, out_off, p.hpg, p.out_sg, dim, p.eps,
                                          rotate, cs, rscale, nullptr, lane, red + group * WPR, 1 + group);
    }
  } else {
    norm_rope_row_generic<InT>(src + in_off, w, dst, peers, out_off, p.hpg, p.out_sg, dim, p.eps, rotate, cs,
                               rscale, pre_bias, lane);
  }
}

// Self-attention form (q AND k given, no affine pre-map): one group of WPR warps per TOKEN.  The q row and the
// k row are both loaded before either is processed (2 x VPL independent 128-bit loads per lane in flight:
// a third more bytes in flight per SM than the one-row-per-warp kernel at two CTAs per SM), and the token's
// (cos, sin) pairs and index arithmetic are shared by the two rows.
template <typename InT, int VPL, int WPR, bool kPeers>
__global__ void __launch_bounds__(kNormRopeWarps * 32, 2)
qk_norm_rope_pair_kernel(const __grid_constant__ NormRopeParams p) {
  __shared__ float red[2][kNormRopeWarps];
  const int warp = threadIdx.x >> 5;
  const int group = warp / WPR;
  const int lane = threadIdx.x - group * (32 * WPR);
  const long long row = static_cast<long long>(blockIdx.x) * (kNormRopeWarps / WPR) + group;
  if (row >= static_cast<long long>(p.B) * p.L) return;
  const int b = p.B == 1 ? 0 : static_cast<int>(row / p.L);
  const int l = static_cast<int>(row - static_cast<long long>(b) * p.L);
  const int dim = p.N * 128;
  const long long in_off = row * dim;

  RowVec<InT> vq[VPL], vk[VPL];
  load_row<InT, VPL, WPR>(vq, static_cast<const InT*>(p.q_in) + in_off, lane);
  load_row<InT, VPL, WPR>(vk, static_cast<const InT*>(p.k_in) + in_off, lane);

  float cs[8];
  bool rotate = false;
  if (p.cos_sin != nullptr) {
    const int gb = b < kMaxBatchGrid ? b : kMaxBatchGrid - 1;
    const int gh = p.grid[gb][1], gw = p.grid[gb][2];
    const int tok = p.tok_offset + l;
    rotate = tok < p.grid[gb][0] * gh * gw;
    if (rotate) {
      const int pf = tok / (gh * gw), ph = (tok / gw) % gh, pw = tok % gw;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int jj = 4 * (lane & 15) + i;
        const int pos = jj < 22 ? pf : (jj < 43 ? ph : pw);
        const float2 v = __ldg(p.cos_sin + pos * 64 + jj);
        cs[2 * i] = v.x;
        cs[2 * i + 1] = v.y;
      }
    }
  }
  if (!rotate) {
#pragma unroll
    for (int i = 0; i < 8; ++i) cs[i] = 0.f;
  }
  const long long out_off = static_cast<long long>(b) * p.out_sb + static_cast<long long>(l) * p.out_sl;
  norm_rope_row<InT, VPL, WPR, false>(vq, p.wq, p.q_out + out_off, kPeers ? p.q_peer : nullptr, out_off, p.hpg,
                                      p.out_sg, dim, p.eps, rotate, cs, 1.0f, nullptr, lane, red[0] + group * WPR,
                                      1 + group);
  norm_rope_row<InT, VPL, WPR, false>(vk, p.wk, p.k_out + out_off, kPeers ? p.k_peer : nullptr, out_off, p.hpg,
                                      p.out_sg, dim, p.eps, rotate, cs, 1.0f, nullptr, lane, red[1] + group * WPR,
                                      1 + group);
}

// ----------------------------------------------------------------------------------------------
// Streaming form of the self-attention prologue (bf16 q AND k, no affine pre-map): persistent CTAs, one per SM.
// Round-1 profile of the token-pair kernel above: 110 registers -> 2 CTAs per SM, 23 % warps active, DRAM busy
// 55 %, ~600 warp instructions per row: every CTA alternates between a load phase (nothing to compute) and a
// compute phase (nothing in flight), and under the power-capped clocks of a denoise step the instruction count
// itself is a co-limiter.  Here the two concerns are decoupled:
//   warp 8       producer: one elected lane streams the rows through a ring of shared-memory stages with 1-D bulk
//                copies (cp.async.bulk, complete_tx on an mbarrier) -- a stage holds the q rows and the k rows of
//                R = 8 / WPR consecutive tokens, which are contiguous in [B, L, dim], i.e. TWO copies per stage;
//                ~150-200 KB per SM are in flight at any time regardless of what the compute warps do;
//   warps 0-15   consumers: WPR warps per token; the row is read from shared memory and unpacked ONCE (registers are
//                plentiful at one CTA per SM), the norm weights sit in shared memory, the token's (cos, sin) pairs are
//                fetched once for q and k.
// The arithmetic (order of operations, rounding points) is that of norm_rope_row: results are bit-identical.
// ----------------------------------------------------------------------------------------------
constexpr int kStreamConsumerWarps = 16;
constexpr int kStreamThreads = (kStreamConsumerWarps + 1) * 32;

__device__ __forceinline__ void bulk_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
      ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)), "l"(kEvictFirst)
      : "memory");
}

// kHasK = false: q only (the query prologue of cross-attention: norm, no rotation partner) -- a stage holds q rows only
template <int VPL, int WPR, bool kHasK = true>
struct StreamSmem {
  static constexpr int kDim = 256 * VPL * WPR;
  static constexpr int kRowBytes = kDim * 2;
  static constexpr int kRows = kStreamConsumerWarps / WPR;          // tokens per stage
  static constexpr int kStageBytes = (kHasK ? 2 : 1) * kRows * kRowBytes;   // q rows | k rows
  static constexpr int kWeightBytes = 2 * kDim * 4;                 // wq | wk, fp32
  static constexpr int kFixed = kWeightBytes + 1024;                // + barriers, reduction scratch, alignment slack
  static constexpr int kStagesMax = (232448 - kFixed) / kStageBytes;
  static constexpr int kStages = kStagesMax < 6 ? kStagesMax : 6;
  static_assert(kStages >= 2, "at least two stages");
  static constexpr int kWOff = kStages * kStageBytes;
  static constexpr int kBarOff = kWOff + kWeightBytes;              // full[kStages], empty[kStages], red[2][8]
  static constexpr int kDynBytes = kBarOff + 2 * kStages * 8 + 2 * kStreamConsumerWarps * 4 + 128;
};

// one row (already in shared memory) -> normalised, rotated, bf16, stored
template <int VPL, int WPR, bool kPeers>
__device__ __forceinline__ void stream_row(const uint8_t* __restrict__ srow, const float* __restrict__ w_s,
                                           bool normed, __nv_bfloat16* __restrict__ out_row,
                                           __nv_bfloat16* const* peers, long long out_off, int hpg, long long out_sg,
                                           float eps, bool rotate, const float (&cs)[8], int lane, float* red,
                                           int bar_id) {
  constexpr int kStride = 32 * WPR;
  constexpr int kDim = 256 * VPL * WPR;
  float x[VPL][8];
  float2 ss2 = make_float2(0.f, 0.f);
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const uint4 raw = *reinterpret_cast<const uint4*>(srow + (lane + kStride * i) * 16);
    const uint32_t wd[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      x[i][2 * e] = __uint_as_float(wd[e] << 16);
      x[i][2 * e + 1] = __uint_as_float(wd[e] & 0xffff0000u);
      const float2 xx = make_float2(x[i][2 * e], x[i][2 * e + 1]);
      ss2 = ffma2(xx, xx, ss2);
    }
  }
  float ss = warp_sum(ss2.x + ss2.y);
  if constexpr (WPR > 1) {
    if ((lane & 31) == 0) red[lane >> 5] = ss;
    named_bar_sync(bar_id, kStride);
    ss = 0.f;
#pragma unroll
    for (int i = 0; i < WPR; ++i) ss += red[i];
  }
  const float rinv = normed ? rsqrtf(ss / static_cast<float>(kDim) + eps) : 1.0f;
  const bool flat = hpg * 128 == kDim && !kPeers;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int vec = lane + kStride * i;
    float y[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) y[e] = x[i][e];
    if (normed) {
      const float4 w0 = *reinterpret_cast<const float4*>(w_s + vec * 8);
      const float4 w1 = *reinterpret_cast<const float4*>(w_s + vec * 8 + 4);
      const float ww[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
      const float2 r2 = make_float2(rinv, rinv);
#pragma unroll
      for (int e = 0; e < 8; e += 2) {
        float2 t = fmul2(make_float2(y[e], y[e + 1]), r2);
        RowVec<__nv_bfloat16>::round_in2(t.x, t.y);
        t = fmul2(t, make_float2(ww[e], ww[e + 1]));
        y[e] = t.x;
        y[e + 1] = t.y;
      }
    }
    uint32_t o[4];
#pragma unroll
    for (int pr = 0; pr < 4; ++pr) {
      float a = y[2 * pr], b = y[2 * pr + 1];
      if (rotate) {
        const float c = cs[2 * pr], sn = cs[2 * pr + 1];
        const float ra = a * c - b * sn;
        const float rb = fmaf(a, sn, b * c);
        a = ra;
        b = rb;
      }
      o[pr] = pack_bf16x2(a, b);
    }
    __nv_bfloat16* dst;
    if (flat) {
      dst = out_row + vec * 8;
    } else {
      const int n = vec >> 4;
      const int d0 = (vec & 15) * 8;
      if constexpr (kPeers) {
        dst = peers[n / hpg] + out_off + (n % hpg) * 128 + d0;
      } else {
        dst = out_row + static_cast<long long>(n / hpg) * out_sg + (n % hpg) * 128 + d0;
      }
    }
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "r"(o[0]),
                 "r"(o[1]), "r"(o[2]), "r"(o[3])
                 : "memory");
  }
}

template <int VPL, int WPR, bool kPeers, bool kHasK = true>
__global__ void __launch_bounds__(kStreamThreads, 1)
qk_norm_rope_stream_kernel(const __grid_constant__ NormRopeParams p) {
  using SM = StreamSmem<VPL, WPR, kHasK>;
  constexpr int kStages = SM::kStages;
  constexpr int kRows = SM::kRows;
  extern __shared__ uint8_t stream_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(stream_smem_raw) + 127) &
                                             ~static_cast<uintptr_t>(127));
  float* w_s = reinterpret_cast<float*>(smem + SM::kWOff);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + SM::kBarOff);
  uint64_t* empty = full + kStages;
  float* red = reinterpret_cast<float*>(empty + kStages);      // [2][8]
  const int warp = threadIdx.x >> 5;
  const long long rows = static_cast<long long>(p.B) * p.L;
  const long long n_chunks = (rows + kRows - 1) / kRows;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], kStreamConsumerWarps);
    }
    fence_mbar_init();
  }
  // norm weights -> shared memory (nullptr = qk_norm disabled: never read)
  for (int i = threadIdx.x; i < SM::kDim; i += kStreamThreads) {
    w_s[i] = p.wq != nullptr ? __ldg(p.wq + i) : 1.0f;
    w_s[SM::kDim + i] = p.wk != nullptr ? __ldg(p.wk + i) : 1.0f;
  }
  __syncthreads();

  if (warp == kStreamConsumerWarps) {
    // ------------------------------------------ producer ------------------------------------------
    const __nv_bfloat16* qg = static_cast<const __nv_bfloat16*>(p.q_in);
    const __nv_bfloat16* kg = static_cast<const __nv_bfloat16*>(p.k_in);
    int it = 0;
    for (long long c = blockIdx.x; c < n_chunks; c += gridDim.x, ++it) {
      const int slot = it % kStages;
      mbar_wait(&empty[slot], ((it / kStages) & 1) ^ 1);
      if (elect_one()) {
        const long long r0 = c * kRows;
        const long long nr = rows - r0 < kRows ? rows - r0 : kRows;
        const uint32_t bytes = static_cast<uint32_t>(nr) * SM::kRowBytes;
        uint8_t* dst = smem + slot * SM::kStageBytes;
        mbar_arrive_expect_tx(&full[slot], (kHasK ? 2 : 1) * bytes);
        bulk_load_1d(dst, qg + r0 * SM::kDim, bytes, &full[slot]);
        if constexpr (kHasK) bulk_load_1d(dst + kRows * SM::kRowBytes, kg + r0 * SM::kDim, bytes, &full[slot]);
      }
      __syncwarp();
    }
  } else {
    // ------------------------------------------ consumers -----------------------------------------
    const int group = warp / WPR;                              // token inside the stage
    const int lane = threadIdx.x - group * (32 * WPR);
    const bool normed_q = p.wq != nullptr, normed_k = p.wk != nullptr;
    int it = 0;
    for (long long c = blockIdx.x; c < n_chunks; c += gridDim.x, ++it) {
      const int slot = it % kStages;
      const long long row = c * kRows + group;
      mbar_wait(&full[slot], (it / kStages) & 1);
      if (row < rows) {
        const int b = p.B == 1 ? 0 : static_cast<int>(row / p.L);
        const int l = static_cast<int>(row - static_cast<long long>(b) * p.L);
        float cs[8];
        bool rotate = false;
        if (p.cos_sin != nullptr) {
          const int gb = b < kMaxBatchGrid ? b : kMaxBatchGrid - 1;
          const int gh = p.grid[gb][1], gw = p.grid[gb][2];
          const int tok = p.tok_offset + l;
          rotate = tok < p.grid[gb][0] * gh * gw;
          if (rotate) {
            const int pf = tok / (gh * gw), ph = (tok / gw) % gh, pw = tok % gw;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int jj = 4 * (lane & 15) + i;
              const int pos = jj < 22 ? pf : (jj < 43 ? ph : pw);
              const float2 v = __ldg(p.cos_sin + pos * 64 + jj);
              cs[2 * i] = v.x;
              cs[2 * i + 1] = v.y;
            }
          }
        }
        if (!rotate) {
#pragma unroll
          for (int i = 0; i < 8; ++i) cs[i] = 0.f;
        }
        const long long out_off = static_cast<long long>(b) * p.out_sb + static_cast<long long>(l) * p.out_sl;
        const uint8_t* sq = smem + slot * SM::kStageBytes + group * SM::kRowBytes;
        const uint8_t* sk = sq + kRows * SM::kRowBytes;
        stream_row<VPL, WPR, kPeers>(sq, w_s, normed_q, p.q_out + out_off, p.q_peer, out_off, p.hpg, p.out_sg, p.eps,
                                     rotate, cs, lane, red + group * WPR, 1 + group);
        if constexpr (kHasK) {
          stream_row<VPL, WPR, kPeers>(sk, w_s + SM::kDim, normed_k, p.k_out + out_off, p.k_peer, out_off, p.hpg,
                                       p.out_sg, p.eps, rotate, cs, lane, red + kStreamConsumerWarps + group * WPR,
                                       1 + group);
        }
      }
      // both rows of this warp's token have been read out of the stage
      __syncwarp();
      if ((threadIdx.x & 31) == 0) mbar_arrive(&empty[slot]);
    }
  }
}

// Head-group scatter of an un-normalised tensor (v) into the Ulysses send layout; pure copy.
struct HeadScatterParams {
  const __nv_bfloat16* in;   // [B, L, N, 128]
  __nv_bfloat16* out;
  int B, L, N, hpg;
  long long out_sb, out_sl, out_sg;
  int n_peers;                              // > 0: group j goes through peer[j] (see NormRopeParams)
  __nv_bfloat16* peer[kMaxPeers];
};

__global__ void __launch_bounds__(256) head_scatter_kernel(const __grid_constant__ HeadScatterParams p) {
  const long long nvec_row = static_cast<long long>(p.N) * 16;   // 16-byte vectors per token
  const long long total = static_cast<long long>(p.B) * p.L * nvec_row;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long row = i / nvec_row;
    const int vec = static_cast<int>(i % nvec_row);
    const int b = static_cast<int>(row / p.L), l = static_cast<int>(row % p.L);
    const int n = vec >> 4, d0 = (vec & 15) * 8;
    const uint4 val = __ldg(reinterpret_cast<const uint4*>(p.in) + i);
    __nv_bfloat16* base = p.n_peers > 0 ? p.peer[n / p.hpg]
                                        : p.out + static_cast<long long>(n / p.hpg) * p.out_sg;
    __nv_bfloat16* dst = base + b * p.out_sb + l * p.out_sl + (n % p.hpg) * 128 + d0;
    *reinterpret_cast<uint4*>(dst) = val;
  }
}

// ----------------------------------------------------------------------------------------------
// Cross-GPU hand-off flags of the fused Ulysses exchange.  A producer kernel's remote stores are followed
// (in stream order) by sp_signal_kernel, which publishes `value` into a flag word in every peer's buffer;
// the consumer runs sp_wait_kernel before the kernel that reads what the peers wrote.  Values only grow
// (one epoch per exchange), so nothing is ever reset.
// ----------------------------------------------------------------------------------------------
struct SpSignalParams {
  uint32_t* flag[kMaxPeers];   // flag word of this rank inside each peer's flag array
  int n;
  uint32_t value;
};

__global__ void sp_signal_kernel(const __grid_constant__ SpSignalParams p) {
  if (threadIdx.x < p.n) {
    __threadfence_system();
    st_release_sys(p.flag[threadIdx.x], p.value);
  }
}

// Spins until *f >= value (wrap-safe).  A rank may legitimately be seconds or minutes late (rank-0-only VAE decode or
// save, offload_model reloads, lazy initialisation, a profiler pause), so the bound is generous and configurable like
// a collective's watchdog: timeout_ns = 0 waits for ever; otherwise the kernel traps (reported by the host as a launch
// failure) after that much wall time (%globaltimer) without the flag arriving.  Default 600 s
// (UVB_KNOB_SP_WAIT_TIMEOUT_S), the default torch.distributed gives an NCCL collective.
__device__ __forceinline__ void sp_wait_flag(const uint32_t* f, uint32_t value, unsigned long long timeout_ns) {
  if (static_cast<int32_t>(ld_acquire_sys(f) - value) < 0) {
    const unsigned long long t0 = globaltimer_ns();
    unsigned backoff = 64;
    while (static_cast<int32_t>(ld_acquire_sys(f) - value) < 0) {
      if (timeout_ns != 0 && globaltimer_ns() - t0 > timeout_ns) __trap();
      __nanosleep(backoff);
      if (backoff < 2048) backoff <<= 1;      // late peers are polled every ~2 us, not hammered over NVLink
    }
  }
}

__global__ void sp_wait_kernel(const uint32_t* flags, int n, uint32_t value, unsigned long long timeout_ns) {
  if (threadIdx.x < n) sp_wait_flag(flags + threadIdx.x, value, timeout_ns);
  __syncthreads();
  __threadfence_system();
}

// sp_signal_kernel followed by sp_wait_kernel in ONE launch (the exchange needs the pair twice per layer)
__global__ void sp_signal_wait_kernel(const __grid_constant__ SpSignalParams p, const uint32_t* flags,
                                      unsigned long long timeout_ns) {
  if (threadIdx.x < p.n) {
    __threadfence_system();
    st_release_sys(p.flag[threadIdx.x], p.value);
    sp_wait_flag(flags + threadIdx.x, p.value, timeout_ns);
  }
  __syncthreads();
  __threadfence_system();
}

}  // namespace uvb
