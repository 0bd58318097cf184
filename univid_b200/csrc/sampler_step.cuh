// One fused pass for the sampler update between two DiT forwards:
//   classifier-free-guidance combine          models/wan/textimage2video.py:385-386
//   FlowUniPCMultistepScheduler.step          models/wan/utils/fm_solvers_unipc.py:657-741
//     convert_model_output (flow -> x0)       :320-323
//     UniC corrector (order 1 / 2)            :549-628
//     UniP predictor (order 1 / 2)            :395-486
// All scalar coefficients (sigma ratios, expm1 terms, the 2x2 solve of the corrector) are computed on the host in
// the reference's own fp32 arithmetic; the kernel is the elementwise part.  Every operation is an explicitly
// rounded IEEE fp32 op in the reference's order (no FMA contraction), so the result is bit-identical to the
// reference evaluated op by op in fp32.  Pure HBM streaming: 6 reads + 3 writes of 4 bytes per latent element
// (the eager chain is ~25 launches and ~100 bytes per element).
//
// history_bf16: the product runs the scheduler inside torch.amp.autocast('cuda', bf16) (textimage2video.py:330-331),
// under which the reference's torch.einsum over the history terms (:471, :614) runs in bf16 on the GPU: its operands
// (rho, D1) are rounded to bf16, the product is rounded to bf16, and in the predictor the bf16 result meets the 0-dim
// fp32 (CPU) tensor alpha_t * B_h in a multiply whose result is bf16 (type promotion keeps the dimensioned operand's
// dtype; the scalar itself enters in fp32).  With history_bf16 != 0 the kernel reproduces exactly those roundings; with 0 it is the
// fp32 chain the reference computes outside autocast (and on the CPU).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace uvb {

struct SamplerStepParams {
  const float* cond;     // model output (or its conditional branch when uncond != nullptr)
  const float* uncond;   // unconditional branch or nullptr
  const float* x;        // sample entering the step
  const float* last;     // last_sample (corrector) or nullptr
  const float* m0;       // newest stored x0 prediction or nullptr
  const float* m1;       // the one before or nullptr
  float* m_out;          // x0 prediction of this step
  float* xc_out;         // corrected sample (= x when there is no corrector): the next step's last_sample
  float* x_next;         // prev_sample
  long long n;
  float guide, sigma;
  int corr_order;        // 0 = none
  float c_a, c_b, c_ab, c_rk, c_rho0, c_rho_last;
  int pred_order;        // 1 or 2
  float p_a, p_b, p_ab, p_rk, p_rho0;
  int history_bf16;      // reproduce the bf16 einsum of the history terms under the product's autocast
};

__device__ __forceinline__ float round_bf16(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }

__device__ __forceinline__ void sampler_step_elem(const SamplerStepParams& p, float vc, float vu, float x, float last,
                                                  float m0, float m1, float& m_t, float& xs, float& xn) {
  const float v = p.uncond != nullptr ? __fadd_rn(vu, __fmul_rn(p.guide, __fsub_rn(vc, vu))) : vc;
  m_t = __fsub_rn(x, __fmul_rn(p.sigma, v));
  xs = x;
  if (p.corr_order > 0) {
    const float xt = __fsub_rn(__fmul_rn(p.c_a, last), __fmul_rn(p.c_b, m0));
    float inner = __fmul_rn(p.c_rho_last, __fsub_rn(m_t, m0));
    if (p.corr_order == 2) {
      const float d1 = __fdiv_rn(__fsub_rn(m1, m0), p.c_rk);
      const float hist = p.history_bf16 ? round_bf16(__fmul_rn(round_bf16(p.c_rho0), round_bf16(d1)))   // bf16 einsum
                                        : __fmul_rn(p.c_rho0, d1);
      inner = __fadd_rn(hist, inner);
    }
    xs = __fsub_rn(xt, __fmul_rn(p.c_ab, inner));
  }
  const float pt = __fsub_rn(__fmul_rn(p.p_a, xs), __fmul_rn(p.p_b, m_t));
  xn = pt;
  if (p.pred_order == 2) {
    const float d1 = __fdiv_rn(__fsub_rn(m0, m_t), p.p_rk);
    if (p.history_bf16) {
      // pred_res = einsum(rho, D1) in bf16; alpha_t * B_h is a 0-dim CPU tensor (the reference keeps its sigmas on the
      // host, :228-229), i.e. a scalar operand: it enters the multiply in fp32 and the RESULT takes pred_res' dtype
      const float pred = round_bf16(__fmul_rn(round_bf16(p.p_rho0), round_bf16(d1)));
      xn = __fsub_rn(pt, round_bf16(__fmul_rn(p.p_ab, pred)));
    } else {
      xn = __fsub_rn(pt, __fmul_rn(p.p_ab, __fmul_rn(p.p_rho0, d1)));
    }
  }
}

__global__ void __launch_bounds__(256) sampler_step_kernel(const SamplerStepParams p) {
  const long long n4 = p.n >> 2;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  const long long tid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
  for (long long i = tid; i < n4; i += stride) {
    const float4 vc = __ldg(reinterpret_cast<const float4*>(p.cond) + i);
    const float4 vu = p.uncond != nullptr ? __ldg(reinterpret_cast<const float4*>(p.uncond) + i) : z;
    const float4 x = __ldg(reinterpret_cast<const float4*>(p.x) + i);
    const float4 la = p.corr_order > 0 ? __ldg(reinterpret_cast<const float4*>(p.last) + i) : z;
    const float4 m0 = (p.corr_order > 0 || p.pred_order == 2) ? __ldg(reinterpret_cast<const float4*>(p.m0) + i) : z;
    const float4 m1 = p.corr_order == 2 ? __ldg(reinterpret_cast<const float4*>(p.m1) + i) : z;
    float4 mt, xs, xn;
    sampler_step_elem(p, vc.x, vu.x, x.x, la.x, m0.x, m1.x, mt.x, xs.x, xn.x);
    sampler_step_elem(p, vc.y, vu.y, x.y, la.y, m0.y, m1.y, mt.y, xs.y, xn.y);
    sampler_step_elem(p, vc.z, vu.z, x.z, la.z, m0.z, m1.z, mt.z, xs.z, xn.z);
    sampler_step_elem(p, vc.w, vu.w, x.w, la.w, m0.w, m1.w, mt.w, xs.w, xn.w);
    reinterpret_cast<float4*>(p.m_out)[i] = mt;
    reinterpret_cast<float4*>(p.xc_out)[i] = xs;
    reinterpret_cast<float4*>(p.x_next)[i] = xn;
  }
  for (long long i = (n4 << 2) + tid; i < p.n; i += stride) {     // ragged tail
    float mt, xs, xn;
    sampler_step_elem(p, p.cond[i], p.uncond != nullptr ? p.uncond[i] : 0.f, p.x[i], p.corr_order > 0 ? p.last[i] : 0.f,
                      (p.corr_order > 0 || p.pred_order == 2) ? p.m0[i] : 0.f, p.corr_order == 2 ? p.m1[i] : 0.f, mt, xs, xn);
    p.m_out[i] = mt;
    p.xc_out[i] = xs;
    p.x_next[i] = xn;
  }
}

}  // namespace uvb
