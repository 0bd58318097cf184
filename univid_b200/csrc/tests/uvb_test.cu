// Stand-alone GPU check + timing of libunivid_b200.so through its C ABI (no torch).
// Usage: uvb_test <case> [args]   -- one case per process so a trapped kernel cannot poison the next.
//   fmha  B Lq Lk N klen keymod iters     compare with a naive fp32 kernel (if Lq*Lk small) and time
//   prol  B L N rope dtype iters          compare with a double-precision host reference and time
//   gemm  M N K act iters                 uvb_linear_bf16 against a naive fp32-accumulate kernel, and time
// Test infrastructure only; nothing here is part of the product library.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../../../include/univid_b200.h"

#ifdef UVB_FMHA_PROFILE
extern "C" void uvb_debug_fmha_profile(void* device_buffer);
#endif

#define CK(x)                                                                              \
  do {                                                                                     \
    cudaError_t e = (x);                                                                   \
    if (e != cudaSuccess) {                                                                \
      printf("CUDA error %s at %s:%d: %s\n", #x, __FILE__, __LINE__, cudaGetErrorString(e)); \
      exit(2);                                                                             \
    }                                                                                      \
  } while (0)

static uint32_t g_seed = 12345;
static float frand() {  // uniform(-1, 1)
  g_seed = g_seed * 1664525u + 1013904223u;
  return ((g_seed >> 8) & 0xFFFFFF) / 8388608.0f - 1.0f;
}
static float nrand() {  // ~N(0,1) (sum of uniforms)
  float s = 0;
  for (int i = 0; i < 6; ++i) s += frand();
  return s * 0.70710678f;
}

// one warp per query row, online softmax in fp32
__global__ void naive_attn(const __nv_bfloat16* q, const __nv_bfloat16* k, const __nv_bfloat16* v,
                           float* o, const int* k_lens, const float* kls, const float* pvw,
                           const float* bias, int B, int Lq, int Lk, int N, float scale) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5);
  if (row >= (long long)B * Lq * N) return;
  const int n = row % N;
  const int lq = (row / N) % Lq;
  const int b = row / ((long long)N * Lq);
  const int klen = k_lens ? min(k_lens[b], Lk) : Lk;
  const __nv_bfloat16* qp = q + (((long long)b * Lq + lq) * N + n) * 128 + lane * 4;
  float qv[4], acc[4] = {0, 0, 0, 0};
  for (int i = 0; i < 4; ++i) qv[i] = __bfloat162float(qp[i]);
  float m = -INFINITY, l = 0;
  for (int j = 0; j < klen; ++j) {
    const __nv_bfloat16* kp = k + (((long long)b * Lk + j) * N + n) * 128 + lane * 4;
    const __nv_bfloat16* vp = v + (((long long)b * Lk + j) * N + n) * 128 + lane * 4;
    float s = 0;
    for (int i = 0; i < 4; ++i) s += qv[i] * __bfloat162float(kp[i]);
    for (int o2 = 16; o2 > 0; o2 >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o2);
    s *= scale;
    if (kls) s *= kls[j];
    const float mn = fmaxf(m, s);
    const float f = __expf(m - mn);
    const float pj = __expf(s - mn);
    l = l * f + pj;
    const float pw = pvw ? pj * pvw[j] : pj;
    for (int i = 0; i < 4; ++i) acc[i] = acc[i] * f + pw * __bfloat162float(vp[i]);
    m = mn;
  }
  float* op = o + (((long long)b * Lq + lq) * N + n) * 128 + lane * 4;
  for (int i = 0; i < 4; ++i) {
    float r = klen > 0 ? acc[i] / l : 0.f;
    if (bias && klen > 0) r += bias[n * 128 + lane * 4 + i];
    op[i] = r;
  }
}

static int run_fmha(int argc, char** argv) {
  if (argc < 9) {
    printf("fmha B Lq Lk N klen keymod iters\n");
    return 2;
  }
  const int B = atoi(argv[2]), Lq = atoi(argv[3]), Lk = atoi(argv[4]), N = atoi(argv[5]);
  const int klen = atoi(argv[6]), keymod = atoi(argv[7]), iters = atoi(argv[8]);
  const size_t nq = (size_t)B * Lq * N * 128, nk = (size_t)B * Lk * N * 128;
  std::vector<__nv_bfloat16> hq(nq), hk(nk), hv(nk);
  for (auto& x : hq) x = __float2bfloat16(nrand());
  for (auto& x : hk) x = __float2bfloat16(nrand());
  for (auto& x : hv) x = __float2bfloat16(nrand());
  // diagnostic input modes: 1 = K zero (uniform P: isolates the PV path), 2 = V ones (isolates
  // normalisation), 3 = V[j][d] = (j % 128 == d) (isolates the QK^T / P path)
  const int inmode = getenv("UVB_INMODE") ? atoi(getenv("UVB_INMODE")) : 0;
  if (inmode == 1) for (auto& x : hk) x = __float2bfloat16(0.f);
  if (inmode == 2) for (auto& x : hv) x = __float2bfloat16(1.f);
  if (inmode == 3)
    for (size_t i = 0; i < nk; ++i) hv[i] = __float2bfloat16(((i / ((size_t)N * 128)) % Lk) % 128 == i % 128 ? 1.f : 0.f);
  __nv_bfloat16 *dq, *dk, *dv, *dout;
  float* dref;
  CK(cudaMalloc(&dq, nq * 2));
  CK(cudaMalloc(&dk, nk * 2));
  CK(cudaMalloc(&dv, nk * 2));
  CK(cudaMalloc(&dout, nq * 2));
  CK(cudaMalloc(&dref, nq * 4));
  CK(cudaMemcpy(dq, hq.data(), nq * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dk, hk.data(), nk * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dv, hv.data(), nk * 2, cudaMemcpyHostToDevice));
  CK(cudaMemset(dout, 0x7f, nq * 2));  // poison (bf16 NaN-ish pattern 0x7f7f is a large finite)
  int* dklens = nullptr;
  if (klen >= 0) {
    std::vector<int> hl(B);
    for (int b = 0; b < B; ++b) hl[b] = b == 0 ? klen : (klen * (b + 1)) % (Lk + 1);
    CK(cudaMalloc(&dklens, B * 4));
    CK(cudaMemcpy(dklens, hl.data(), B * 4, cudaMemcpyHostToDevice));
  }
  float *dkls = nullptr, *dpvw = nullptr, *dbias = nullptr;
  if (keymod) {
    const int Lp = (Lk + 127) / 128 * 128;
    std::vector<float> a(Lp, 1.f), w(Lp, 1.f), bb(N * 128);
    for (int j = 0; j < Lk; ++j) {
      a[j] = j < Lk / 4 ? 1.3f : 1.0f;
      w[j] = j < Lk / 4 ? 1.2f : 1.0f;
    }
    for (auto& x : bb) x = 0.1f * frand();
    CK(cudaMalloc(&dkls, Lp * 4));
    CK(cudaMalloc(&dpvw, Lp * 4));
    CK(cudaMalloc(&dbias, N * 128 * 4));
    CK(cudaMemcpy(dkls, a.data(), Lp * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dpvw, w.data(), Lp * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dbias, bb.data(), N * 128 * 4, cudaMemcpyHostToDevice));
  }
  const float scale = 1.0f / sqrtf(128.f);
  // split-unit workspace (zero-filled once); UVB_TEST_NOWS=1 exercises the no-workspace schedule
  void* dws = nullptr;
  int64_t ws_bytes = 0;
  if (!(getenv("UVB_TEST_NOWS") && atoi(getenv("UVB_TEST_NOWS")))) {
    ws_bytes = uvb_fmha_workspace_bytes();
    if (ws_bytes <= 0) {
      printf("FAIL uvb_fmha_workspace_bytes -> %lld\n", (long long)ws_bytes);
      return 1;
    }
    CK(cudaMalloc(&dws, ws_bytes));
    CK(cudaMemset(dws, 0, ws_bytes));
  }
  auto launch = [&]() {
    return keymod ? uvb_xattn_fwd_bf16(dq, dk, dv, dout, dklens, dkls, dpvw, dbias, B, Lq, Lk, N,
                                       nullptr, nullptr, nullptr, nullptr, scale, dws, ws_bytes, nullptr)
                  : uvb_fmha_fwd_bf16(dq, dk, dv, dout, dklens, B, Lq, Lk, N, nullptr, nullptr,
                                      nullptr, nullptr, scale, dws, ws_bytes, nullptr);
  };
  int rc = launch();
  if (rc != 0) {
    printf("FAIL launch rc=%d: %s\n", rc, uvb_last_error());
    return 1;
  }
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("FAIL kernel: %s\n", cudaGetErrorString(e));
    return 1;
  }
  int status = 0;
  const double work = (double)B * Lq * (double)Lk * N;
  if (work <= 4e9) {
    const long long rows = (long long)B * Lq * N;
    naive_attn<<<(unsigned)((rows + 3) / 4), 128>>>(dq, dk, dv, dref, dklens, dkls, dpvw, dbias, B,
                                                   Lq, Lk, N, scale);
    CK(cudaDeviceSynchronize());
    std::vector<__nv_bfloat16> ho(nq);
    std::vector<float> hr(nq);
    CK(cudaMemcpy(ho.data(), dout, nq * 2, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(hr.data(), dref, nq * 4, cudaMemcpyDeviceToHost));
    double maxabs = 0, dot = 0, na = 0, nb = 0;
    size_t worst = 0;
    int nan = 0;
    for (size_t i = 0; i < nq; ++i) {
      const double a = __bfloat162float(ho[i]), r = hr[i];
      if (!(a == a)) ++nan;
      const double d = fabs(a - r);
      if (d > maxabs) {
        maxabs = d;
        worst = i;
      }
      dot += a * r;
      na += a * a;
      nb += r * r;
    }
    const double cosv = dot / (sqrt(na) * sqrt(nb) + 1e-30);
    const bool ok = nan == 0 && maxabs <= 2e-2 && cosv >= 0.9999;
    printf("%s fmha B=%d Lq=%d Lk=%d N=%d klen=%d keymod=%d: max_abs=%.3e cos=%.7f nan=%d", ok ? "PASS" : "FAIL",
           B, Lq, Lk, N, klen, keymod, maxabs, cosv, nan);
    if (!ok) {
      const size_t r = worst / 128;
      printf("  worst at row=%zu (lq=%zu n=%zu) d=%zu got=%f ref=%f", r, (r / N) % Lq, r % N, worst % 128,
             __bfloat162float(ho[worst]), hr[worst]);
      status = 1;
    }
    printf("\n");
    if (!ok) {
      // error map by 128x? blocks to localise descriptor/layout mistakes
      printf("  per-(qtile,dpanel) max err for b=0,n=0:\n");
      for (int qt = 0; qt < (Lq + 127) / 128 && qt < 12; ++qt) {
        printf("   qtile %d:", qt);
        for (int dp = 0; dp < 8; ++dp) {
          double mx = 0;
          for (int r2 = qt * 128; r2 < min(Lq, qt * 128 + 128); ++r2)
            for (int d = dp * 16; d < dp * 16 + 16; ++d) {
              const size_t i = ((size_t)r2 * N + 0) * 128 + d;
              mx = fmax(mx, fabs(__bfloat162float(ho[i]) - hr[i]));
            }
          printf(" %.2e", mx);
        }
        printf("\n");
      }
      printf("  first row got: ");
      for (int d = 0; d < 8; ++d) printf("%.4f ", __bfloat162float(ho[d]));
      printf("\n  first row ref: ");
      for (int d = 0; d < 8; ++d) printf("%.4f ", hr[d]);
      printf("\n");
    }
  }
  if (getenv("UVB_TEST_TIMELINE") && atoi(getenv("UVB_TEST_TIMELINE"))) {
    // per-CTA timeline of one (warm) launch: unit time per SM and drift between SMs
    unsigned long long* dtl;
    const int nsm = 148;
    CK(cudaMalloc(&dtl, nsm * 32 * 8));
    launch();
    CK(cudaDeviceSynchronize());
    CK(cudaMemset(dtl, 0, nsm * 32 * 8));
    uvb_debug_fmha_timeline(dtl);
    launch();
    CK(cudaDeviceSynchronize());
    uvb_debug_fmha_timeline(nullptr);
    std::vector<unsigned long long> tl(nsm * 32);
    CK(cudaMemcpy(tl.data(), dtl, nsm * 32 * 8, cudaMemcpyDeviceToHost));
    unsigned long long t0 = ~0ull;
    for (int g = 0; g < nsm; ++g)
      if (tl[g * 32 + 1] && tl[g * 32 + 1] < t0) t0 = tl[g * 32 + 1];
    printf("TIMELINE cta smid start_us | piece durations (us) ... | end_us\n");
    for (int g = 0; g < nsm; ++g) {
      if (!tl[g * 32 + 1]) continue;
      printf("TL %3d %3llu %7.1f |", g, tl[g * 32], (tl[g * 32 + 1] - t0) * 1e-3);
      int k = 2;
      for (; k < 32 && tl[g * 32 + k]; ++k) printf(" %6.1f", (tl[g * 32 + k] - tl[g * 32 + k - 1]) * 1e-3);
      printf(" | %7.1f\n", (tl[g * 32 + k - 1] - t0) * 1e-3);
    }
  }
#ifdef UVB_FMHA_PROFILE
  {
    // wait-cycle counters of one warm launch (library built with -DUVB_FMHA_PROFILE)
    unsigned long long* dpr;
    const int nsm = 148;
    CK(cudaMalloc(&dpr, nsm * 16 * 8));
    launch();
    CK(cudaDeviceSynchronize());
    CK(cudaMemset(dpr, 0, nsm * 16 * 8));
    uvb_debug_fmha_profile(dpr);
    launch();
    CK(cudaDeviceSynchronize());
    uvb_debug_fmha_profile(nullptr);
    std::vector<unsigned long long> pr(nsm * 16);
    CK(cudaMemcpy(pr.data(), dpr, nsm * 16 * 8, cudaMemcpyDeviceToHost));
    static const char* names[12] = {"sm0_total", "sm0_wait_S", "sm0_wait_PV", "sm1_total", "sm1_wait_S", "sm1_wait_PV",
                                    "mma_total", "mma_wait_P0", "mma_wait_P1", "mma_wait_KV", "mma_wait_Q", "steps"};
    for (int par = 0; par < 2; ++par) {       // even CTAs (leaders of pairs) and odd CTAs separately
      double sum[12] = {0};
      int n = 0;
      for (int g = par; g < nsm; g += 2) {
        if (!pr[g * 16 + 0]) continue;
        ++n;
        for (int i = 0; i < 12; ++i) sum[i] += (double)pr[g * 16 + i];
      }
      if (!n) continue;
      printf("PROF %s CTAs (%d):", par ? "odd " : "even", n);
      const double steps = sum[11] > 0 ? sum[11] / n : 0;
      for (int i = 0; i < 12; ++i) printf(" %s=%.0f", names[i], sum[i] / n);
      if (steps > 0)
        printf("\n     per step: sm0 %.0f (S %.0f, PV %.0f)  sm1 %.0f (S %.0f, PV %.0f)  mma %.0f (P0 %.0f P1 %.0f KV %.0f)",
               sum[0] / n / steps, sum[1] / n / steps, sum[2] / n / steps, sum[3] / n / steps, sum[4] / n / steps,
               sum[5] / n / steps, sum[6] / n / steps, sum[7] / n / steps, sum[8] / n / steps, sum[9] / n / steps);
      double extra[4] = {0};
      for (int g = par; g < nsm; g += 2) {
        if (!pr[g * 16 + 0]) continue;
        for (int i = 0; i < 4; ++i) extra[i] += (double)pr[g * 16 + 12 + i];
      }
      if (steps > 0)
        printf("\n     per step: sm0 set-up %.0f epilogue %.0f  sm1 set-up %.0f epilogue %.0f", extra[0] / n / steps,
               extra[1] / n / steps, extra[2] / n / steps, extra[3] / n / steps);
      printf("\n");
    }
  }
#endif
  if (iters > 0) {
    // timing: inputs here exceed L2 only for the big shapes; flush L2 between iterations anyway
    char* flush;
    const size_t fb = 256u << 20;
    CK(cudaMalloc(&flush, fb));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    for (int i = 0; i < 3; ++i) launch();
    CK(cudaDeviceSynchronize());
    double best = 1e30, tot = 0;
    for (int i = 0; i < iters; ++i) {
      CK(cudaMemsetAsync(flush, i, fb));
      CK(cudaEventRecord(e0));
      launch();
      CK(cudaEventRecord(e1));
      CK(cudaEventSynchronize(e1));
      float ms;
      CK(cudaEventElapsedTime(&ms, e0, e1));
      best = fmin(best, ms);
      tot += ms;
    }
    double eff_lk = klen >= 0 ? klen : Lk;
    const double flop = 4.0 * B * Lq * eff_lk * N * 128;
    printf("TIME fmha B=%d Lq=%d Lk=%d N=%d: best %.3f ms (%.1f TFLOP/s)  mean %.3f ms (%.1f TFLOP/s)\n", B,
           Lq, Lk, N, best, flop / best * 1e-9, tot / iters, flop / (tot / iters) * 1e-9);
  }
  return status;
}

static int run_prol(int argc, char** argv) {
  if (argc < 8) {
    printf("prol B L N rope dtype iters\n");
    return 2;
  }
  const int B = atoi(argv[2]), L = atoi(argv[3]), N = atoi(argv[4]), rope = atoi(argv[5]);
  const int dtype = atoi(argv[6]), iters = atoi(argv[7]);
  const int dim = N * 128;
  const size_t n = (size_t)B * L * dim;
  std::vector<float> hq(n), hk(n), wq(dim), wk(dim);
  for (auto& x : hq) x = nrand();
  for (auto& x : hk) x = nrand() * 0.5f;
  for (auto& x : wq) x = 1.f + 0.1f * frand();
  for (auto& x : wk) x = 1.f + 0.1f * frand();
  std::vector<__nv_bfloat16> bq(n), bk(n);
  if (dtype == UVB_BF16) {
    for (size_t i = 0; i < n; ++i) {
      bq[i] = __float2bfloat16(hq[i]);
      bk[i] = __float2bfloat16(hk[i]);
      hq[i] = __bfloat162float(bq[i]);
      hk[i] = __bfloat162float(bk[i]);
    }
  }
  // grid: pick (f, h, w) with f*h*w <= L, leaving a few padding tokens when possible
  int gf = 1, gh = 1, gw = 1;
  {
    gw = 13;
    gh = 7;
    gf = L / (gw * gh);
    if (gf < 1) {
      gf = 1;
      gh = 1;
      gw = L > 3 ? L - 3 : L;
    }
    if (gf > 1024) gf = 1024;
  }
  std::vector<float> cs(1024 * 64 * 2);
  for (int pos = 0; pos < 1024; ++pos)
    for (int j = 0; j < 64; ++j) {
      double ang;
      if (j < 22) ang = pos * pow(10000.0, -(2.0 * j) / 44.0);
      else if (j < 43) ang = pos * pow(10000.0, -(2.0 * (j - 22)) / 42.0);
      else ang = pos * pow(10000.0, -(2.0 * (j - 43)) / 42.0);
      cs[(pos * 64 + j) * 2] = (float)cos(ang);
      cs[(pos * 64 + j) * 2 + 1] = (float)sin(ang);
    }
  void *dq, *dk;
  float *dwq, *dwk, *dcs;
  __nv_bfloat16 *oq, *ok;
  const size_t esz = dtype == UVB_BF16 ? 2 : 4;
  CK(cudaMalloc(&dq, n * esz));
  CK(cudaMalloc(&dk, n * esz));
  CK(cudaMalloc(&dwq, dim * 4));
  CK(cudaMalloc(&dwk, dim * 4));
  CK(cudaMalloc(&dcs, cs.size() * 4));
  CK(cudaMalloc(&oq, n * 2));
  CK(cudaMalloc(&ok, n * 2));
  if (dtype == UVB_BF16) {
    CK(cudaMemcpy(dq, bq.data(), n * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dk, bk.data(), n * 2, cudaMemcpyHostToDevice));
  } else {
    CK(cudaMemcpy(dq, hq.data(), n * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dk, hk.data(), n * 4, cudaMemcpyHostToDevice));
  }
  CK(cudaMemcpy(dwq, wq.data(), dim * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dwk, wk.data(), dim * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dcs, cs.data(), cs.size() * 4, cudaMemcpyHostToDevice));
  std::vector<int32_t> grid(3 * B);
  for (int b = 0; b < B; ++b) {
    grid[3 * b] = gf;
    grid[3 * b + 1] = gh;
    grid[3 * b + 2] = gw;
  }
  const float eps = 1e-6f;
  auto launch = [&]() {
    return uvb_qk_norm_rope(dq, dk, dtype, dwq, dwk, rope ? dcs : nullptr, nullptr, nullptr, oq, ok, B, L,
                            N, rope ? grid.data() : nullptr, 0, eps, N, (int64_t)L * dim, dim, 0, nullptr);
  };
  int rc = launch();
  if (rc != 0) {
    printf("FAIL launch rc=%d: %s\n", rc, uvb_last_error());
    return 1;
  }
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("FAIL kernel: %s\n", cudaGetErrorString(e));
    return 1;
  }
  int status = 0;
  {
    std::vector<__nv_bfloat16> gq(n), gk(n);
    CK(cudaMemcpy(gq.data(), oq, n * 2, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(gk.data(), ok, n * 2, cudaMemcpyDeviceToHost));
    double maxerr = 0;
    long long bad = 0;
    const size_t rows = (size_t)B * L;
    const size_t step = rows > 4096 ? rows / 4096 : 1;
    for (size_t r = 0; r < rows; r += step) {
      const int l = r % L;
      for (int which = 0; which < 2; ++which) {
        const float* x = (which ? hk.data() : hq.data()) + r * dim;
        const float* w = which ? wk.data() : wq.data();
        const __nv_bfloat16* g = (which ? gk.data() : gq.data()) + r * dim;
        double ss = 0;
        for (int i = 0; i < dim; ++i) ss += (double)x[i] * x[i];
        const double rinv = 1.0 / sqrt(ss / dim + eps);
        const bool rot = rope && l < gf * gh * gw;
        const int pf = l / (gh * gw), ph = (l / gw) % gh, pw = l % gw;
        for (int i = 0; i < dim; i += 2) {
          double a = x[i] * rinv, b = x[i + 1] * rinv;
          if (dtype == UVB_BF16) {
            a = __bfloat162float(__float2bfloat16((float)a));
            b = __bfloat162float(__float2bfloat16((float)b));
          }
          a *= w[i];
          b *= w[i + 1];
          if (rot) {
            const int jj = (i % 128) / 2;
            const int pos = jj < 22 ? pf : (jj < 43 ? ph : pw);
            const double c = cs[(pos * 64 + jj) * 2], s = cs[(pos * 64 + jj) * 2 + 1];
            const double ra = a * c - b * s, rb = a * s + b * c;
            a = ra;
            b = rb;
          }
          const double ea = fabs(__bfloat162float(g[i]) - a), eb = fabs(__bfloat162float(g[i + 1]) - b);
          // a 1-ulp flip (2^-7 relative) of the intermediate `.type_as` rounding of BOTH inputs moves a
          // rotated output by up to (|cos| + |sin|) <= 1.42 ulp of the pair magnitude, plus half an ulp
          // for the final rounding: 2 ulp in total
          const double mag = fmax(fmax(fabs(a), fabs(b)), fmax(fabs(x[i] * rinv * w[i]), fabs(x[i + 1] * rinv * w[i + 1])));
          const double tol_a = 2.0 * 0.0078125 * mag + 1e-3, tol_b = tol_a;
          if (ea > tol_a || eb > tol_b) ++bad;
          maxerr = fmax(maxerr, fmax(ea, eb));
        }
      }
    }
    const bool okk = bad == 0;
    printf("%s prol B=%d L=%d N=%d rope=%d dtype=%d grid=(%d,%d,%d): max_abs=%.3e bad=%lld\n", okk ? "PASS" : "FAIL",
           B, L, N, rope, dtype, gf, gh, gw, maxerr, bad);
    if (!okk) status = 1;
  }
  if (iters > 0) {
    char* flush;
    const size_t fb = 256u << 20;
    CK(cudaMalloc(&flush, fb));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    for (int i = 0; i < 3; ++i) launch();
    CK(cudaDeviceSynchronize());
    double best = 1e30, tot = 0;
    for (int i = 0; i < iters; ++i) {
      CK(cudaMemsetAsync(flush, i, fb));
      CK(cudaEventRecord(e0));
      launch();
      CK(cudaEventRecord(e1));
      CK(cudaEventSynchronize(e1));
      float ms;
      CK(cudaEventElapsedTime(&ms, e0, e1));
      best = fmin(best, ms);
      tot += ms;
    }
    const double bytes = 2.0 * n * (esz + 2);
    printf("TIME prol B=%d L=%d N=%d dtype=%d: best %.3f ms (%.0f GB/s)  mean %.3f ms (%.0f GB/s)\n", B, L, N,
           dtype, best, bytes / best * 1e-6, tot / iters, bytes / (tot / iters) * 1e-6);
  }
  return status;
}

// one thread per output element: fp32 accumulation in k order, bias, bf16 rounding, optional tanh-GELU in fp32
__global__ void naive_linear(const __nv_bfloat16* x, const __nv_bfloat16* w, const float* bias, float* y, int M,
                             int N, int K, int act) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)M * N) return;
  const int n = idx % N;
  const long long m = idx / N;
  const __nv_bfloat16* xr = x + m * K;
  const __nv_bfloat16* wr = w + (long long)n * K;
  float acc = 0.f;
  for (int k = 0; k < K; ++k) acc += __bfloat162float(xr[k]) * __bfloat162float(wr[k]);
  if (bias) acc += bias[n];
  float v = __bfloat162float(__float2bfloat16(acc));
  if (act == 1) v = 0.5f * v * (1.0f + tanhf(0.7978845608028654f * (v + 0.044715f * v * v * v)));
  y[idx] = v;
}

static int run_gemm(int argc, char** argv) {
  if (argc < 7) {
    printf("gemm M N K act iters\n");
    return 2;
  }
  const int M = atoi(argv[2]), N = atoi(argv[3]), K = atoi(argv[4]), act = atoi(argv[5]), iters = atoi(argv[6]);
  const size_t nx = (size_t)M * K, nw = (size_t)N * K, ny = (size_t)M * N;
  std::vector<__nv_bfloat16> hx(nx), hw(nw);
  std::vector<float> hb(N);
  const float ws = 1.0f / sqrtf((float)K);
  for (auto& v : hx) v = __float2bfloat16(nrand());
  for (auto& v : hw) v = __float2bfloat16(nrand() * ws);
  for (auto& v : hb) v = __bfloat162float(__float2bfloat16(0.5f * frand()));
  __nv_bfloat16 *dx, *dw, *dy;
  float *db, *dref;
  CK(cudaMalloc(&dx, nx * 2));
  CK(cudaMalloc(&dw, nw * 2));
  CK(cudaMalloc(&dy, ny * 2));
  CK(cudaMalloc(&db, N * 4));
  CK(cudaMemcpy(dx, hx.data(), nx * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dw, hw.data(), nw * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(db, hb.data(), N * 4, cudaMemcpyHostToDevice));
  CK(cudaMemset(dy, 0x7f, ny * 2));
  auto launch = [&]() { return uvb_linear_bf16(dx, dw, db, dy, M, N, K, K, K, N, act, nullptr); };
  int rc = launch();
  if (rc != 0) {
    printf("FAIL launch rc=%d: %s\n", rc, uvb_last_error());
    return 1;
  }
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("FAIL kernel: %s\n", cudaGetErrorString(e));
    return 1;
  }
  int status = 0;
  if ((double)M * N * K <= 6e11) {
    CK(cudaMalloc(&dref, ny * 4));
    naive_linear<<<(unsigned)((ny + 255) / 256), 256>>>(dx, dw, db, dref, M, N, K, act);
    CK(cudaDeviceSynchronize());
    std::vector<__nv_bfloat16> ho(ny);
    std::vector<float> hr(ny);
    CK(cudaMemcpy(ho.data(), dy, ny * 2, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(hr.data(), dref, ny * 4, cudaMemcpyDeviceToHost));
    double maxrel = 0, dot = 0, na = 0, nb = 0;
    size_t worst = 0, exact = 0, bad = 0;
    for (size_t i = 0; i < ny; ++i) {
      const double a = __bfloat162float(ho[i]), r = hr[i];
      const double rr = __bfloat162float(__float2bfloat16((float)r));
      if (a == rr) ++exact;
      const double d = fabs(a - r) / fmax(1.0, fabs(r));
      if (!(d <= 1.6e-2)) ++bad;
      if (!(d <= maxrel)) {
        maxrel = d;
        worst = i;
      }
      dot += a * r;
      na += a * a;
      nb += r * r;
    }
    const double cosv = dot / (sqrt(na) * sqrt(nb) + 1e-30);
    const bool ok = bad == 0 && cosv >= 0.99999;
    printf("%s gemm M=%d N=%d K=%d act=%d: max_err=%.3e cos=%.8f bitexact=%.4f bad=%zu", ok ? "PASS" : "FAIL", M, N, K,
           act, maxrel, cosv, (double)exact / ny, bad);
    if (!ok) {
      printf("  worst at m=%zu n=%zu got=%f ref=%f\n", worst / N, worst % N, __bfloat162float(ho[worst]), hr[worst]);
      printf("  bad elements per (128-row, 64-col) block, first 8 x 8 blocks:\n");
      for (int bm = 0; bm < 8 && bm * 128 < M; ++bm) {
        printf("   rows %4d:", bm * 128);
        for (int bn = 0; bn < 8 && bn * 64 < N; ++bn) {
          int cnt = 0;
          for (int r = bm * 128; r < (bm + 1) * 128 && r < M; ++r)
            for (int c = bn * 64; c < (bn + 1) * 64 && c < N; ++c) {
              const double a = __bfloat162float(ho[(size_t)r * N + c]), rf = hr[(size_t)r * N + c];
              if (!(fabs(a - rf) / fmax(1.0, fabs(rf)) <= 1.6e-2)) ++cnt;
            }
          printf(" %5d", cnt);
        }
        printf("\n");
      }
      status = 1;
    }
    printf("\n");
  }
  if (iters > 0 && status == 0) {
    // rotate over several input/weight sets? x (M*K) is >= L2 only for the big shapes; flush L2 between iterations
    const size_t fb = 256u << 20;
    void* flush;
    CK(cudaMalloc(&flush, fb));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    for (int i = 0; i < 3; ++i) launch();
    CK(cudaDeviceSynchronize());
    float best = 1e30f, tot = 0;
    for (int i = 0; i < iters; ++i) {
      CK(cudaMemsetAsync(flush, i, fb));
      CK(cudaEventRecord(e0));
      launch();
      CK(cudaEventRecord(e1));
      CK(cudaEventSynchronize(e1));
      float ms;
      CK(cudaEventElapsedTime(&ms, e0, e1));
      best = fmin(best, ms);
      tot += ms;
    }
    const double flop = 2.0 * M * (double)N * K;
    printf("TIME gemm M=%d N=%d K=%d act=%d: best %.3f ms (%.1f TFLOP/s)  mean %.3f ms (%.1f TFLOP/s)\n", M, N, K, act,
           best, flop / best * 1e-9, tot / iters, flop / (tot / iters) * 1e-9);
  }
  return status;
}

// UVB_KNOBS="fmha_pair=0,gemm_ctas=1": the TEST binary maps its environment onto uvb_set_knob (the library itself
// never reads the environment)
static void apply_knobs() {
  const char* e = getenv("UVB_KNOBS");
  if (e == nullptr) return;
  static const char* names[UVB_KNOB_COUNT] = {"fmha_pair", "fmha_split", "gemm_ctas", "gemm_bn", "gemm_small",
                                              "prologue_pair", "fmha_poly", "sp_wait_timeout_s",
                                              "xattn_pair"};
  std::vector<char> buf(e, e + strlen(e) + 1);
  for (char* tok = strtok(buf.data(), ","); tok != nullptr; tok = strtok(nullptr, ",")) {
    char* eq = strchr(tok, '=');
    if (eq == nullptr) continue;
    *eq = 0;
    int found = -1;
    for (int i = 0; i < UVB_KNOB_COUNT; ++i)
      if (!strcmp(tok, names[i])) found = i;
    if (found < 0 || uvb_set_knob(found, atoi(eq + 1)) != 0) {
      printf("bad knob '%s'\n", tok);
      exit(2);
    }
  }
}

int main(int argc, char** argv) {
  if (argc < 2) {
    printf("usage: uvb_test fmha|prol|gemm ...\n");
    return 2;
  }
  apply_knobs();
  if (!strcmp(argv[1], "fmha")) return run_fmha(argc, argv);
  if (!strcmp(argv[1], "prol")) return run_prol(argc, argv);
  if (!strcmp(argv[1], "gemm")) return run_gemm(argc, argv);
  printf("unknown case %s\n", argv[1]);
  return 2;
}
