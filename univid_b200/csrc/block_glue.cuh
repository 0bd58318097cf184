// Fused elementwise glue of a WanAttentionBlock around the attention hot path,
// HBM-bound.  Reference: WanAttentionBlock.forward, models/wan/utils/modules/model.py:219-259, and
// WanLayerNorm.forward :88-98.  Per token row, in one pass over the fp32 residual stream:
//
//     x' = x + y * gate                      (gated residual, model.py:247 / :252 / :257; y is the bf16
//                                             output of the o / ffn projection; gate == NULL means 1)
//     h  = LayerNorm(x') [* ln_w + ln_b]     (norm1 / norm2 without affine, norm3 with, eps inside rsqrt)
//     h  = h * (1 + scale) + shift           (adaLN modulation, model.py:244 / :255; NULL = no modulation)
//     store x' (fp32) and h (bf16)
//
// The reference evaluates the same expression in fp32 (x is the fp32 residual stream, the modulation e is
// asserted fp32, model.py:237-240) and the consuming nn.Linear rounds h to bf16 under autocast, so rounding h
// once here is the same rounding point.  Eagerly the block moves ~116 bytes per element through HBM for
// this glue (LayerNorm, addcmul, three autocast casts, mul, add, ...); fused it is 6-12.
//
// One group of WPR warps owns a row; lane l holds the 8-float vectors l + 32*WPR*i, the row stays in
// registers (fp32), mean and variance are reduced with warp shuffles (+ shared memory across the warps of
// a group), every global access is 128 bits wide.
#pragma once
#include "ptx.cuh"

namespace uvb {

struct BlockGlueParams {
  const float* x_in;            // [rows, dim] fp32 residual stream
  const __nv_bfloat16* y;       // [rows, dim] bf16 branch output, or nullptr (no residual update)
  const float* gate;            // modulation chunk, element (b, l, c) at gate + b*mod_sb + l*mod_sl + c; nullptr = 1
  float* x_out;                 // [rows, dim] fp32 (may alias x_in), or nullptr when y == nullptr
  const float* ln_w;            // [dim] LayerNorm affine weight or nullptr
  const float* ln_b;            // [dim] LayerNorm affine bias or nullptr
  const float* scale;           // modulation chunk (same addressing as gate) or nullptr
  const float* shift;           // modulation chunk or nullptr
  __nv_bfloat16* h_out;         // [rows, dim] bf16 normalised (+modulated) rows, or nullptr (residual only)
  long long rows;               // B * L
  int L;                        // rows per batch sample
  int dim;
  long long mod_sb, mod_sl;     // batch / token strides (elements) of the modulation chunks; mod_sl = 0 broadcasts
  const int* mod_index;         // [rows] or nullptr: modulation row of each token (instead of its position l) --
                                // per-token timesteps with few distinct values keep one row per value
  float eps;
  int ln_round_bf16;            // 1: round the LayerNorm result to bf16 before the modulation -- WanLayerNorm returns
                                // its input's dtype (model.py:98), which is bf16 for the first block of the DiT (the
                                // patch embedding runs under autocast); the caller widens that x to fp32 (exact)
};

constexpr int kGlueWarps = 8;

__device__ __forceinline__ float glue_warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// VPL vectors of 8 floats per lane, WPR warps per row: dim = 256 * VPL * WPR.
template <int VPL, int WPR>
__global__ void __launch_bounds__(kGlueWarps * 32) block_glue_kernel(const __grid_constant__ BlockGlueParams p) {
  constexpr int kStride = 32 * WPR;
  __shared__ float red[kGlueWarps][2];
  const int warp = threadIdx.x >> 5;
  const int group = warp / WPR;
  const int lane = threadIdx.x - group * kStride;          // lane inside the row group
  const long long row = static_cast<long long>(blockIdx.x) * (kGlueWarps / WPR) + group;
  if (row >= p.rows) return;                               // whole groups leave together (named barriers are per group)
  const int b = static_cast<int>(row / p.L);
  const int l = static_cast<int>(row - static_cast<long long>(b) * p.L);
  const long long mod_off = b * p.mod_sb + (p.mod_index != nullptr ? __ldg(p.mod_index + row) : l) * p.mod_sl;
  const float* xr = p.x_in + row * p.dim;

  float x[VPL][8];
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int c = (lane + kStride * i) * 8;
    float4 a0, a1;
    // plain (coherent) streaming loads: x_out may alias x_in (in-place residual update), and .nc requires the
    // memory to be read-only for the lifetime of the kernel
    asm volatile("ld.global.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(a0.x), "=f"(a0.y), "=f"(a0.z), "=f"(a0.w) : "l"(xr + c) : "memory");
    asm volatile("ld.global.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(a1.x), "=f"(a1.y), "=f"(a1.z), "=f"(a1.w) : "l"(xr + c + 4) : "memory");
    x[i][0] = a0.x; x[i][1] = a0.y; x[i][2] = a0.z; x[i][3] = a0.w;
    x[i][4] = a1.x; x[i][5] = a1.y; x[i][6] = a1.z; x[i][7] = a1.w;
  }

  if (p.y != nullptr) {
    const __nv_bfloat16* yr = p.y + row * p.dim;
    float* xo = p.x_out + row * p.dim;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int c = (lane + kStride * i) * 8;
      uint4 raw;
      asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
                   : "=r"(raw.x), "=r"(raw.y), "=r"(raw.z), "=r"(raw.w) : "l"(yr + c));
      const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
      float g[8];
      if (p.gate != nullptr) {
        const float4 g0 = __ldg(reinterpret_cast<const float4*>(p.gate + mod_off + c));
        const float4 g1 = __ldg(reinterpret_cast<const float4*>(p.gate + mod_off + c) + 1);
        g[0] = g0.x; g[1] = g0.y; g[2] = g0.z; g[3] = g0.w; g[4] = g1.x; g[5] = g1.y; g[6] = g1.z; g[7] = g1.w;
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) g[e] = 1.0f;
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        // x + y * gate: the product is rounded to fp32 before the add, like the two eager kernels
        x[i][2 * e] = __fadd_rn(x[i][2 * e], __fmul_rn(__uint_as_float(w[e] << 16), g[2 * e]));
        x[i][2 * e + 1] = __fadd_rn(x[i][2 * e + 1], __fmul_rn(__uint_as_float(w[e] & 0xffff0000u), g[2 * e + 1]));
      }
      asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(xo + c), "f"(x[i][0]),
                   "f"(x[i][1]), "f"(x[i][2]), "f"(x[i][3]) : "memory");
      asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(xo + c + 4), "f"(x[i][4]),
                   "f"(x[i][5]), "f"(x[i][6]), "f"(x[i][7]) : "memory");
    }
  }
  if (p.h_out == nullptr) return;

  // mean, then variance around the mean (two passes over registers: no cancellation)
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
#pragma unroll
    for (int e = 0; e < 8; ++e) s += x[i][e];
  }
  s = glue_warp_sum(s);
  if constexpr (WPR > 1) {
    if ((lane & 31) == 0) red[warp][0] = s;
    named_bar_sync(1 + group, kStride);
    s = 0.f;
#pragma unroll
    for (int i = 0; i < WPR; ++i) s += red[group * WPR + i][0];
  }
  const float mean = s / static_cast<float>(p.dim);
  float v = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float d = x[i][e] - mean;
      v = fmaf(d, d, v);
    }
  }
  v = glue_warp_sum(v);
  if constexpr (WPR > 1) {
    if ((lane & 31) == 0) red[warp][1] = v;
    named_bar_sync(1 + group, kStride);
    v = 0.f;
#pragma unroll
    for (int i = 0; i < WPR; ++i) v += red[group * WPR + i][1];
  }
  const float rstd = rsqrtf(v / static_cast<float>(p.dim) + p.eps);

  __nv_bfloat16* hr = p.h_out + row * p.dim;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int c = (lane + kStride * i) * 8;
    float h[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) h[e] = (x[i][e] - mean) * rstd;
    if (p.ln_w != nullptr) {
      const float4 w0 = __ldg(reinterpret_cast<const float4*>(p.ln_w + c));
      const float4 w1 = __ldg(reinterpret_cast<const float4*>(p.ln_w + c) + 1);
      const float ww[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
      for (int e = 0; e < 8; ++e) h[e] *= ww[e];
    }
    if (p.ln_b != nullptr) {
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.ln_b + c));
      const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.ln_b + c) + 1);
      const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int e = 0; e < 8; ++e) h[e] += bb[e];
    }
    if (p.ln_round_bf16) {
#pragma unroll
      for (int e = 0; e < 8; ++e) h[e] = __bfloat162float(__float2bfloat16_rn(h[e]));
    }
    if (p.scale != nullptr) {
      const float4 s0 = __ldg(reinterpret_cast<const float4*>(p.scale + mod_off + c));
      const float4 s1 = __ldg(reinterpret_cast<const float4*>(p.scale + mod_off + c) + 1);
      const float4 t0 = __ldg(reinterpret_cast<const float4*>(p.shift + mod_off + c));
      const float4 t1 = __ldg(reinterpret_cast<const float4*>(p.shift + mod_off + c) + 1);
      const float sc[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
      const float sh[8] = {t0.x, t0.y, t0.z, t0.w, t1.x, t1.y, t1.z, t1.w};
#pragma unroll
      for (int e = 0; e < 8; ++e) h[e] = fmaf(h[e], 1.0f + sc[e], sh[e]);
    }
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(hr + c),
                 "r"(pack_bf16x2(h[0], h[1])), "r"(pack_bf16x2(h[2], h[3])), "r"(pack_bf16x2(h[4], h[5])),
                 "r"(pack_bf16x2(h[6], h[7]))
                 : "memory");
  }
}

}  // namespace uvb
