// C ABI of libunivid_b200.so (see include/univid_b200.h).  Host side only: argument checks,
// TMA tensor-map construction and kernel launches.  No torch types, no device allocation, no sync.
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/univid_b200.h"
#include "fmha_fwd_sm100.cuh"
#include "qk_norm_rope.cuh"
#include "block_glue.cuh"
#include "gemm_bf16_sm100.cuh"
#include "sampler_step.cuh"

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

#define UVB_CUDA(call)                                                              \
  do {                                                                              \
    cudaError_t e__ = (call);                                                       \
    if (e__ != cudaSuccess)                                                         \
      return fail(UVB_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e__));   \
  } while (0)

using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                   const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                   const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// libcuda is resolved at run time through the runtime, so the library loads (and exports its
// symbols) on hosts without a driver.
EncodeTiledFn get_encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) ==
            cudaSuccess &&
        q == cudaDriverEntryPointSuccess) {
      fn = reinterpret_cast<EncodeTiledFn>(p);
    }
  }
  return fn;
}

// Explicit tuning knobs (uvb_set_knob): process-wide, changed only by an API call -- the library never reads
// the environment.  Defaults are the shipped configuration.
int g_knobs[UVB_KNOB_COUNT] = {
    /* UVB_KNOB_FMHA_PAIR     */ 1,   // CTA-pair attention kernel for long key sequences
    /* UVB_KNOB_FMHA_SPLIT    */ 1,   // stream-K split of the remainder units
    /* UVB_KNOB_GEMM_CTAS     */ 2,   // 2 = CTA pairs (cta_group::2), 1 = single CTAs
    /* UVB_KNOB_GEMM_BN       */ 0,   // 0 = per-problem choice, 192 | 256 pins the tile width
    /* UVB_KNOB_GEMM_SMALL    */ 1,   // single-wave 128x64 tiles for small problems
    /* UVB_KNOB_PROLOGUE_PAIR */ 2,   // q and k both given: 2 streaming kernel (bulk-copy ring), 1 token-pair kernel, 0 row kernel
    /* UVB_KNOB_FMHA_POLY     */ 0,   // CTA-pair attention kernel: 1 exp2 pair in every n on the FMA pipe (0, 2, 3, 4)
    /* UVB_KNOB_SP_WAIT_TIMEOUT_S */ 600,   // seconds a rank waits for a peer's hand-off flag before trapping; 0 = for ever
    /* UVB_KNOB_XATTN_PAIR    */ 1,   // CTA-pair variant of the short-key (cross-attention) kernel
};

int check_device() {
  static int ok_dev = -1;
  int dev = 0;
  UVB_CUDA(cudaGetDevice(&dev));
  if (dev == ok_dev) return UVB_OK;
  int major = 0;
  UVB_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  if (major != 10)
    return fail(UVB_ERR_UNSUPPORTED, "device %d has compute capability %d.x; sm_100 required", dev,
                major);
  ok_dev = dev;
  return UVB_OK;
}

// [B, L, N, 128] bf16 viewed as (d, token, head, batch); box = one 64-column panel of a 128-token tile
int make_tile_map(CUtensorMap* tm, const void* base, int B, int L, int N, const int64_t* strides,
                  int box_rows = 128) {
  EncodeTiledFn enc = get_encode_tiled();
  if (enc == nullptr) return fail(UVB_ERR_CUDA, "cuTensorMapEncodeTiled is not available");
  int64_t sb = static_cast<int64_t>(L) * N * 128, sl = static_cast<int64_t>(N) * 128, sh = 128;
  if (strides != nullptr) {
    sb = strides[0];
    sl = strides[1];
    sh = strides[2];
  }
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0)
    return fail(UVB_ERR_INVALID, "tensor base pointer must be 16-byte aligned");
  if (sl % 8 != 0 || sh % 8 != 0 || sb % 8 != 0 || sl <= 0 || sh <= 0 || sb <= 0)
    return fail(UVB_ERR_INVALID, "strides must be positive multiples of 8 elements");
  const cuuint64_t gdim[4] = {128, static_cast<cuuint64_t>(L), static_cast<cuuint64_t>(N),
                              static_cast<cuuint64_t>(B)};
  const cuuint64_t gstr[3] = {static_cast<cuuint64_t>(sl) * 2, static_cast<cuuint64_t>(sh) * 2,
                              static_cast<cuuint64_t>(sb) * 2};
  const cuuint32_t box[4] = {64, static_cast<cuuint32_t>(box_rows), 1, 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), gdim, gstr,
                   box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(UVB_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", static_cast<int>(r));
  return UVB_OK;
}

// row-major bf16 matrix [rows, cols] with leading dimension ld (elements) as a 2-D TMA map, dims (cols, rows)
int make_matrix_map(CUtensorMap* tm, const void* base, int rows, int cols, int64_t ld, int box_rows) {
  EncodeTiledFn enc = get_encode_tiled();
  if (enc == nullptr) return fail(UVB_ERR_CUDA, "cuTensorMapEncodeTiled is not available");
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0)
    return fail(UVB_ERR_INVALID, "matrix base pointer must be 16-byte aligned");
  if (ld < cols || ld % 8 != 0) return fail(UVB_ERR_INVALID, "leading dimension %lld must be >= %d and a multiple of 8",
                                             static_cast<long long>(ld), cols);
  const cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  const cuuint64_t gstr[1] = {static_cast<cuuint64_t>(ld) * 2};
  const cuuint32_t box[2] = {64, static_cast<cuuint32_t>(box_rows)};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(UVB_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", static_cast<int>(r));
  return UVB_OK;
}

// column-group scatter destinations of a GEMM launch (see GemmParams::n_peers)
struct GemmPeers {
  void* const* ptrs = nullptr;
  int n = 0;
  int64_t ld = 0;
};

template <int kCtas, int kBN>
int launch_gemm(uvb::GemmParams& p, const void* x, const void* w, void* y, int64_t ldx, int64_t ldw, int64_t ldy,
                int workers, cudaStream_t stream, const GemmPeers& peers = GemmPeers()) {
  using SM = uvb::GemmSmem<kCtas, kBN>;
  int rc;
  if ((rc = make_matrix_map(&p.tm_a, x, p.M, p.K, ldx, uvb::kGemmBM)) != UVB_OK) return rc;
  if ((rc = make_matrix_map(&p.tm_b, w, p.N, p.K, ldw, kBN / kCtas)) != UVB_OK) return rc;
  if (peers.n == 0) {
    if ((rc = make_matrix_map(&p.tm_c, y, p.M, p.N, ldy, uvb::kGemmBM)) != UVB_OK) return rc;
  } else {
    p.n_peers = peers.n;
    p.peer_cols = p.N / peers.n;
    for (int j = 0; j < peers.n; ++j) {
      if ((rc = make_matrix_map(&p.tm_c_peer[j], peers.ptrs[j], p.M, p.peer_cols, peers.ld, uvb::kGemmBM)) != UVB_OK)
        return rc;
    }
  }
  p.n_m = (p.M + uvb::kGemmBM * kCtas - 1) / (uvb::kGemmBM * kCtas);
  p.n_n = (p.N + kBN - 1) / kBN;
  const long long tiles = static_cast<long long>(p.n_m) * p.n_n;
  if (tiles > 0x7fffffffLL) return fail(UVB_ERR_INVALID, "too many output tiles");
  // Rasterisation: the slab of B a group of column tiles covers (group_n x kBN x K bf16) should stay in the
  // 126 MB L2 while the row tiles stream past, so A is read from DRAM once per group: as few groups as an
  // 80 MB slab budget allows, all of the same width.
  const double slab_tile = static_cast<double>(kBN) * p.K * 2.0;
  int fit = static_cast<int>(80e6 / slab_tile);
  if (fit < 1) fit = 1;
  const int groups = (p.n_n + fit - 1) / fit;
  p.group_n = (p.n_n + groups - 1) / groups;
  auto kern = uvb::gemm_bf16_kernel<kCtas, kBN>;
  static int attr_dev = -1;         // per instantiation: the attribute is set once per device
  int dev = 0;
  UVB_CUDA(cudaGetDevice(&dev));
  if (dev != attr_dev) {
    UVB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SM::kDynBytes));
    attr_dev = dev;
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.blockDim = dim3(uvb::kGemmThreads);
  cfg.dynamicSmemBytes = SM::kDynBytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  if (kCtas == 2) {
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
  }
  long long n = workers;
  if (n > tiles) n = tiles;
  cfg.gridDim = dim3(static_cast<unsigned>(n * kCtas));
  UVB_CUDA(cudaLaunchKernelEx(&cfg, kern, p));
  return UVB_OK;
}

// persistent workers (CTAs or CTA pairs) the device can hold at once
template <int kCtas>
int gemm_workers(int sms, int* out) {
  if (kCtas == 1) {
    *out = sms;
    return UVB_OK;
  }
  static int cached_dev = -1, cached = 0;
  int dev = 0;
  UVB_CUDA(cudaGetDevice(&dev));
  if (dev != cached_dev) {
    using SM = uvb::GemmSmem<2, 256>;
    auto kern = uvb::gemm_bf16_kernel<2, 256>;
    UVB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SM::kDynBytes));
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.blockDim = dim3(uvb::kGemmThreads);
    cfg.dynamicSmemBytes = SM::kDynBytes;
    cfg.gridDim = dim3(static_cast<unsigned>(sms / 2 * 2));
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int n = 0;
    UVB_CUDA(cudaOccupancyMaxActiveClusters(&n, kern, &cfg));      // one pair per TPC
    if (n <= 0) return fail(UVB_ERR_UNSUPPORTED, "no CTA pair of the GEMM kernel fits on this device");
    cached = n < sms / 2 ? n : sms / 2;
    cached_dev = dev;
  }
  *out = cached;
  return UVB_OK;
}

// Tile width: the persistent grid runs ceil(tiles / workers) waves; 192-wide tiles cost ~3 % more per flop
// (less operand reuse) but often fill the last wave better (N = 1536: 10.4 -> 13.8 waves).
template <int kCtas>
int pick_tile_n(const uvb::GemmParams& p, int workers) {
  const int forced = g_knobs[UVB_KNOB_GEMM_BN];
  if (forced == 192 || forced == 256) return forced;
  const long long n_m = (p.M + uvb::kGemmBM * kCtas - 1) / (uvb::kGemmBM * kCtas);
  auto cost = [&](int bn, double penalty) {
    const long long tiles = n_m * ((p.N + bn - 1) / bn);
    const long long waves = (tiles + workers - 1) / workers;
    return static_cast<double>(waves) * bn * penalty;
  };
  return cost(192, 1.03) < cost(256, 1.0) ? 192 : 256;
}

unsigned long long* g_timeline = nullptr;   // diagnostics: see uvb_debug_fmha_timeline
unsigned long long* g_prof = nullptr;       // UVB_FMHA_PROFILE lab builds: see uvb_debug_fmha_profile

constexpr int kShortKeyTiles = 16;   // <= 2048 keys: query-block-pipelined variant
constexpr size_t kWsFlagBytes = 4096;   // flags [sms][2] u32 live at the start of the workspace

int sm_count(int* out) {
  static int cached_dev = -1, cached = 0;
  int dev = 0;
  UVB_CUDA(cudaGetDevice(&dev));
  if (dev != cached_dev) {
    UVB_CUDA(cudaDeviceGetAttribute(&cached, cudaDevAttrMultiProcessorCount, dev));
    cached_dev = dev;
  }
  *out = cached;
  return UVB_OK;
}

size_t fmha_ws_bytes(int sms) {
  return kWsFlagBytes + static_cast<size_t>(sms) * uvb::kWsSlotFloats * sizeof(float);
}

// K/V ring slots of 16 KiB in the CTA-pair attention kernels: 8 when half of P stays in TMEM, 6 when both halves of
// P go through shared memory (kEarlyS: two more 16 KiB panels per query tile)
template <bool kEarlyS>
constexpr int pair_stages() { return kEarlyS ? 6 : 8; }

// CTA pairs of the attention kernel the device can hold at once (one per TPC); also sets the kernel's
// shared-memory attribute for the current device.  0 pairs = fall back to single CTAs.
template <auto kKernel, int kDynBytes>
int pair_workers_of(int sms, int* out) {
  static int cached_dev = -1, cached = 0;
  int dev = 0;
  UVB_CUDA(cudaGetDevice(&dev));
  if (dev != cached_dev) {
    auto kern = kKernel;
    UVB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kDynBytes));
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.blockDim = dim3(uvb::kFmhaThreads);
    cfg.dynamicSmemBytes = kDynBytes;
    cfg.gridDim = dim3(static_cast<unsigned>(sms / 2 * 2));
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int n = 0;
    UVB_CUDA(cudaOccupancyMaxActiveClusters(&n, kern, &cfg));
    cached = n < sms / 2 ? n : sms / 2;
    if (cached < 0) cached = 0;
    cached_dev = dev;
  }
  *out = cached;
  return UVB_OK;
}

template <bool kEarlyS>
int fmha_pair_workers(int sms, int* out) {
  return pair_workers_of<uvb::fmha_fwd_kernel<pair_stages<kEarlyS>(), 1, 2, false, kEarlyS>,
                         uvb::FmhaSmem<pair_stages<kEarlyS>(), 1, 2, kEarlyS>::kDynBytes>(sms, out);
}

// short-key (cross-attention) variant as CTA pairs: two Q/O buffers, four 16 KiB K/V stages, P panels
constexpr int kShortPairStages = 4;
template <bool kKeyMod>
int xattn_pair_workers(int sms, int* out) {
  return pair_workers_of<uvb::fmha_fwd_kernel<kShortPairStages, 2, 2, kKeyMod>,
                         uvb::FmhaSmem<kShortPairStages, 2, 2>::kDynBytes>(sms, out);
}

template <bool kKeyMod>
int launch_fmha(const void* q, const void* k, const void* v, void* o, const int32_t* k_lens,
                const float* key_logit_scale, const float* key_pv_weight, const float* out_bias,
                int B, int Lq, int Lk, int N, const int64_t* qs, const int64_t* ks,
                const int64_t* vs, const int64_t* os, float scale, void* workspace,
                int64_t workspace_bytes, void* stream, void* const* o_peers = nullptr, int n_peers = 0,
                int head_offset = 0, int total_heads = 0) {
  if (q == nullptr || k == nullptr || v == nullptr || (o == nullptr && n_peers == 0))
    return fail(UVB_ERR_INVALID, "null tensor pointer");
  if (n_peers < 0 || n_peers > uvb::kMaxPeers || (n_peers > 0 && (o_peers == nullptr || Lq % n_peers != 0)))
    return fail(UVB_ERR_INVALID, "bad peer list (n_peers=%d, Lq=%d)", n_peers, Lq);
  if (B <= 0 || Lq <= 0 || Lk <= 0 || N <= 0 || B > 65535 || N > 65535)
    return fail(UVB_ERR_INVALID, "bad shape B=%d Lq=%d Lk=%d N=%d", B, Lq, Lk, N);
  if (!(scale > 0.f)) return fail(UVB_ERR_INVALID, "softmax scale must be positive");
  int rc = check_device();
  if (rc != UVB_OK) return rc;
  int sms = 0;
  if ((rc = sm_count(&sms)) != UVB_OK) return rc;
  if (sms * 2 * sizeof(uint32_t) > kWsFlagBytes) return fail(UVB_ERR_UNSUPPORTED, "%d SMs", sms);

  uvb::FmhaParams p;
  memset(&p, 0, sizeof(p));
  if ((rc = make_tile_map(&p.tm_q, q, B, Lq, N, qs)) != UVB_OK) return rc;
  if ((rc = make_tile_map(&p.tm_k, k, B, Lk, N, ks)) != UVB_OK) return rc;
  if ((rc = make_tile_map(&p.tm_kh, k, B, Lk, N, ks, 64)) != UVB_OK) return rc;
  if ((rc = make_tile_map(&p.tm_v, v, B, Lk, N, vs)) != UVB_OK) return rc;
  if (n_peers == 0) {
    if ((rc = make_tile_map(&p.tm_o, o, B, Lq, N, os)) != UVB_OK) return rc;
  } else {
    // rank j owns output rows [j*chunk, (j+1)*chunk): its buffer is [B, chunk, total_heads, 128]
    if (total_heads < head_offset + N) return fail(UVB_ERR_INVALID, "head_offset + N > total_heads");
    p.n_peers = n_peers;
    p.chunk = Lq / n_peers;
    p.o_head_off = head_offset;
    p.o_heads = total_heads;
    for (int j = 0; j < n_peers; ++j) {
      if (o_peers[j] == nullptr) return fail(UVB_ERR_INVALID, "null peer pointer %d", j);
      p.o_peer_ptr[j] = static_cast<__nv_bfloat16*>(o_peers[j]);
      if ((rc = make_tile_map(&p.tm_o_peer[j], o_peers[j], B, p.chunk, total_heads, nullptr)) != UVB_OK) return rc;
    }
  }
  p.k_lens = k_lens;
  p.key_logit_scale = key_logit_scale;
  p.key_pv_weight = key_pv_weight;
  p.out_bias = out_bias;
  p.Lq = Lq;
  p.Lk = Lk;
  p.N = N;
  const long long n_kv = (Lk + uvb::kBlockN - 1) / uvb::kBlockN;
  // Few key tiles per query block (cross-attention: 4): the per-block prologue/epilogue dominates, so use
  // the variant that prefetches the next block's Q and drains O through a second buffer (3 ring slots);
  // long key sequences keep the deeper K/V ring instead -- and, for plain attention, run as CTA pairs
  // (cta_group::2: 512-row units, half of every K / V tile per CTA).
  const bool short_keys = n_kv <= kShortKeyTiles;
  int pairs = 0;
  const int pair_mode = g_knobs[UVB_KNOB_FMHA_PAIR];       // 0 single CTAs, 1 pairs, 2 pairs with early S release
  if (short_keys && g_knobs[UVB_KNOB_XATTN_PAIR] != 0) {
    if ((rc = xattn_pair_workers<kKeyMod>(sms, &pairs)) != UVB_OK) return rc;
  }
  if (!short_keys && !kKeyMod && pair_mode != 0) {
#ifdef UVB_LAB_VARIANTS
    if (pair_mode == 2) {
      if ((rc = fmha_pair_workers<true>(sms, &pairs)) != UVB_OK) return rc;
    } else
#endif
    {
      if (pair_mode != 1 || g_knobs[UVB_KNOB_FMHA_POLY] != 0) {
#ifndef UVB_LAB_VARIANTS
        return fail(UVB_ERR_UNSUPPORTED, "UVB_KNOB_FMHA_PAIR=%d / UVB_KNOB_FMHA_POLY=%d are lab variants (measured and "
                    "rejected, DESIGN.md sec. 4 S); build with -DUVB_LAB_VARIANTS to run them", pair_mode,
                    g_knobs[UVB_KNOB_FMHA_POLY]);
#endif
      }
      if ((rc = fmha_pair_workers<false>(sms, &pairs)) != UVB_OK) return rc;
    }
  }
  const bool pair = pairs > 0;
  const int unit_rows = uvb::kUnitRows * (pair ? 2 : 1);
  const int workers = pair ? pairs : sms;            // CTAs (or CTA pairs) walking the unit list
  p.n_qt = (Lq + unit_rows - 1) / unit_rows;
  const long long units = static_cast<long long>(B) * N * p.n_qt;
  if (units > 0x7fffffffLL) return fail(UVB_ERR_INVALID, "too many query blocks");
  p.n_units = static_cast<int>(units);
  p.scale_log2 = scale * 1.4426950408889634f;
  p.timeline = g_timeline;
  p.prof = g_prof;

  // Persistent grid: one CTA per SM; with a workspace the remainder units are cut into equal key ranges
  // (then even fewer units than SMs keep every SM busy).
  long long grid_w = units < workers ? units : workers;
  // Splitting pays when a query block spans many key tiles; with a handful (cross-attention: 4) a partial
  // costs almost as much as a whole block (Q load, 128 KiB partial through L2, merge), so those never split.
  if (workspace != nullptr && g_knobs[UVB_KNOB_FMHA_SPLIT] != 0 && !short_keys) {
    if ((reinterpret_cast<uintptr_t>(workspace) & 255) != 0)
      return fail(UVB_ERR_INVALID, "workspace must be 256-byte aligned");
    if (workspace_bytes < static_cast<int64_t>(fmha_ws_bytes(sms)))
      return fail(UVB_ERR_INVALID, "workspace too small: %lld < %zu bytes (uvb_fmha_workspace_bytes)",
                  static_cast<long long>(workspace_bytes), fmha_ws_bytes(sms));
    p.flags = static_cast<uint32_t*>(workspace);
    p.ws = reinterpret_cast<float*>(static_cast<char*>(workspace) + kWsFlagBytes);
    const long long iters = units * n_kv;
    grid_w = iters < workers ? iters : workers;
  }

  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.blockDim = dim3(uvb::kFmhaThreads);
  cfg.stream = static_cast<cudaStream_t>(stream);
  cudaLaunchAttribute attr[1];
  if (pair) {
    cfg.gridDim = dim3(static_cast<unsigned>(2 * grid_w));
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    // (the kernels' max dynamic shared memory was set by *_pair_workers)
    if (short_keys) {
      cfg.dynamicSmemBytes = uvb::FmhaSmem<kShortPairStages, 2, 2>::kDynBytes;
      UVB_CUDA(cudaLaunchKernelEx(&cfg, uvb::fmha_fwd_kernel<kShortPairStages, 2, 2, kKeyMod>, p));
      return UVB_OK;
    }
#ifdef UVB_LAB_VARIANTS
    if (pair_mode == 2) {
      cfg.dynamicSmemBytes = uvb::FmhaSmem<pair_stages<true>(), 1, 2, true>::kDynBytes;
      UVB_CUDA(cudaLaunchKernelEx(&cfg, uvb::fmha_fwd_kernel<pair_stages<true>(), 1, 2, false, true>, p));
      return UVB_OK;
    }
#endif
    constexpr int kSt = pair_stages<false>();
    cfg.dynamicSmemBytes = uvb::FmhaSmem<kSt, 1, 2, false>::kDynBytes;
    void (*kern)(uvb::FmhaParams) = uvb::fmha_fwd_kernel<kSt, 1, 2, false, false, 0>;
#ifdef UVB_LAB_VARIANTS
    switch (g_knobs[UVB_KNOB_FMHA_POLY]) {
      case 2: kern = uvb::fmha_fwd_kernel<kSt, 1, 2, false, false, 2>; break;
      case 3: kern = uvb::fmha_fwd_kernel<kSt, 1, 2, false, false, 3>; break;
      case 4: kern = uvb::fmha_fwd_kernel<kSt, 1, 2, false, false, 4>; break;
      default: break;
    }
    if (g_knobs[UVB_KNOB_FMHA_POLY] != 0)     // the attribute is per instantiation
      UVB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, cfg.dynamicSmemBytes));
#endif
    UVB_CUDA(cudaLaunchKernelEx(&cfg, kern, p));
    return UVB_OK;
  }
  void (*kern)(uvb::FmhaParams) = short_keys ? uvb::fmha_fwd_kernel<3, 2, 1, kKeyMod> : uvb::fmha_fwd_kernel<4, 1, 1, kKeyMod>;
  const int smem = short_keys ? uvb::FmhaSmem<3, 2>::kDynBytes : uvb::FmhaSmem<4, 1>::kDynBytes;
  UVB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  cfg.dynamicSmemBytes = smem;
  cfg.gridDim = dim3(static_cast<unsigned>(grid_w));
  UVB_CUDA(cudaLaunchKernelEx(&cfg, kern, p));
  return UVB_OK;
}

template <typename InT, bool kPeers>
int launch_norm_rope_t(const uvb::NormRopeParams& p, cudaStream_t stream) {
  const int dim = p.N * 128;
  const dim3 block(uvb::kNormRopeWarps * 32);
  // q and k together (self-attention): one warp group per token, both rows in flight (UVB_KNOB_PROLOGUE_PAIR = 0
  // keeps the one-row-per-group kernel)
  // bf16 rows at the product's widths without the affine pre-map: the streaming kernel (persistent CTAs, bulk-copy
  // ring) -- q and k together (self-attention) or q alone (the query prologue of cross-attention)
  if (g_knobs[UVB_KNOB_PROLOGUE_PAIR] == 2 && p.q_in != nullptr && p.pre_bias == nullptr && p.row_scale == nullptr &&
      sizeof(InT) == 2 && (dim == 1536 || dim == 3072 || dim == 5120)) {
    int sms = 0;
    int rc = sm_count(&sms);
    if (rc != UVB_OK) return rc;
    auto launch = [&](auto kern, int dyn_bytes, int rows_per_stage) -> int {
      UVB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn_bytes));
      const long long chunks = (static_cast<long long>(p.B) * p.L + rows_per_stage - 1) / rows_per_stage;
      const unsigned grid = static_cast<unsigned>(chunks < sms ? chunks : sms);
      kern<<<grid, uvb::kStreamThreads, dyn_bytes, stream>>>(p);
      UVB_CUDA(cudaGetLastError());
      return UVB_OK;
    };
    if (p.k_in != nullptr) {
      switch (dim) {
        case 1536: return launch(uvb::qk_norm_rope_stream_kernel<6, 1, kPeers>, uvb::StreamSmem<6, 1>::kDynBytes, uvb::StreamSmem<6, 1>::kRows);
        case 3072: return launch(uvb::qk_norm_rope_stream_kernel<6, 2, kPeers>, uvb::StreamSmem<6, 2>::kDynBytes, uvb::StreamSmem<6, 2>::kRows);
        default: return launch(uvb::qk_norm_rope_stream_kernel<5, 4, kPeers>, uvb::StreamSmem<5, 4>::kDynBytes, uvb::StreamSmem<5, 4>::kRows);
      }
    } else if (static_cast<long long>(p.B) * p.L >= 4096) {      // a few hundred context rows stay on the row kernel
      switch (dim) {
        case 1536: return launch(uvb::qk_norm_rope_stream_kernel<6, 1, kPeers, false>, uvb::StreamSmem<6, 1, false>::kDynBytes, uvb::StreamSmem<6, 1, false>::kRows);
        case 3072: return launch(uvb::qk_norm_rope_stream_kernel<6, 2, kPeers, false>, uvb::StreamSmem<6, 2, false>::kDynBytes, uvb::StreamSmem<6, 2, false>::kRows);
        default: return launch(uvb::qk_norm_rope_stream_kernel<5, 4, kPeers, false>, uvb::StreamSmem<5, 4, false>::kDynBytes, uvb::StreamSmem<5, 4, false>::kRows);
      }
    }
  }
  const bool pair = g_knobs[UVB_KNOB_PROLOGUE_PAIR] != 0 && p.q_in != nullptr && p.k_in != nullptr && p.pre_bias == nullptr &&
                    p.row_scale == nullptr && sizeof(InT) == 2;
  if (pair) {
    const long long tokens = static_cast<long long>(p.B) * p.L;
    auto blocks = [&](int wpr) {
      const int per_cta = uvb::kNormRopeWarps / wpr;
      return static_cast<unsigned>((tokens + per_cta - 1) / per_cta);
    };
    bool done = true;
    switch (dim) {
      case 1536: uvb::qk_norm_rope_pair_kernel<InT, 6, 1, kPeers><<<blocks(1), block, 0, stream>>>(p); break;
      case 2048: uvb::qk_norm_rope_pair_kernel<InT, 8, 1, kPeers><<<blocks(1), block, 0, stream>>>(p); break;
      case 3072: uvb::qk_norm_rope_pair_kernel<InT, 6, 2, kPeers><<<blocks(2), block, 0, stream>>>(p); break;
      case 4096: uvb::qk_norm_rope_pair_kernel<InT, 8, 2, kPeers><<<blocks(2), block, 0, stream>>>(p); break;
      case 5120: uvb::qk_norm_rope_pair_kernel<InT, 5, 4, kPeers><<<blocks(4), block, 0, stream>>>(p); break;
      default: done = false; break;
    }
    if (done) {
      UVB_CUDA(cudaGetLastError());
      return UVB_OK;
    }
  }
  const long long units = 2LL * p.B * p.L;   // one warp group per (row, q|k)
  auto blocks = [&](int wpr) {
    const int per_cta = uvb::kNormRopeWarps / wpr;
    return static_cast<unsigned>((units + per_cta - 1) / per_cta);
  };
  switch (dim) {   // VPL * WPR = dim / 256
    case 1536: uvb::qk_norm_rope_kernel<InT, 6, 1, kPeers><<<blocks(1), block, 0, stream>>>(p); break;
    case 2048: uvb::qk_norm_rope_kernel<InT, 8, 1, kPeers><<<blocks(1), block, 0, stream>>>(p); break;
    case 3072: uvb::qk_norm_rope_kernel<InT, 6, 2, kPeers><<<blocks(2), block, 0, stream>>>(p); break;
    case 4096: uvb::qk_norm_rope_kernel<InT, 8, 2, kPeers><<<blocks(2), block, 0, stream>>>(p); break;
    case 5120: uvb::qk_norm_rope_kernel<InT, 5, 4, kPeers><<<blocks(4), block, 0, stream>>>(p); break;
    default: uvb::qk_norm_rope_kernel<InT, 0, 1, kPeers><<<blocks(1), block, 0, stream>>>(p); break;
  }
  UVB_CUDA(cudaGetLastError());
  return UVB_OK;
}

template <typename InT>
int launch_norm_rope(const uvb::NormRopeParams& p, cudaStream_t stream) {
  return p.n_peers > 0 ? launch_norm_rope_t<InT, true>(p, stream) : launch_norm_rope_t<InT, false>(p, stream);
}

int linear_impl(const void* x, const void* w, const float* bias, void* y, int M, int N, int K, int64_t ldx, int64_t ldw,
                int64_t ldy, int act, void* stream, const GemmPeers& peers) {
  if (x == nullptr || w == nullptr) return fail(UVB_ERR_INVALID, "null matrix pointer");
  if (M <= 0 || N <= 0 || K <= 0) return fail(UVB_ERR_INVALID, "bad shape M=%d N=%d K=%d", M, N, K);
  if (N % 8 != 0 || K % 8 != 0) return fail(UVB_ERR_INVALID, "N=%d and K=%d must be multiples of 8", N, K);
  if (act != UVB_ACT_NONE && act != UVB_ACT_GELU_TANH) return fail(UVB_ERR_INVALID, "bad activation %d", act);
  if (bias != nullptr && (reinterpret_cast<uintptr_t>(bias) & 3) != 0) return fail(UVB_ERR_INVALID, "bias alignment");
  int rc = check_device();
  if (rc != UVB_OK) return rc;
  int sms = 0;
  if ((rc = sm_count(&sms)) != UVB_OK) return rc;
  uvb::GemmParams p;
  memset(&p, 0, sizeof(p));
  p.bias = bias;
  p.M = M;
  p.N = N;
  p.K = K;
  p.act = act;
  const int ctas = g_knobs[UVB_KNOB_GEMM_CTAS] == 1 ? 1 : 2;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int workers = 0;
  // Small problems (the 512 context rows of the cross-attention k / v projections, single-row probes): when 128 x 64
  // tiles of single CTAs fit in ONE wave, latency is what counts -- many small tiles instead of a dozen 256-wide pairs.
  const bool allow_small = g_knobs[UVB_KNOB_GEMM_SMALL] != 0;
  const long long small_tiles = static_cast<long long>((M + uvb::kGemmBM - 1) / uvb::kGemmBM) * ((N + 63) / 64);
  if (allow_small && small_tiles <= sms) return launch_gemm<1, 64>(p, x, w, y, ldx, ldw, ldy, sms, st, peers);
  if (ctas == 2) {
    if ((rc = gemm_workers<2>(sms, &workers)) != UVB_OK) return rc;
    return pick_tile_n<2>(p, workers) == 192 ? launch_gemm<2, 192>(p, x, w, y, ldx, ldw, ldy, workers, st, peers)
                                             : launch_gemm<2, 256>(p, x, w, y, ldx, ldw, ldy, workers, st, peers);
  }
  if ((rc = gemm_workers<1>(sms, &workers)) != UVB_OK) return rc;
  return pick_tile_n<1>(p, workers) == 192 ? launch_gemm<1, 192>(p, x, w, y, ldx, ldw, ldy, workers, st, peers)
                                           : launch_gemm<1, 256>(p, x, w, y, ldx, ldw, ldy, workers, st, peers);
}

}  // namespace

extern "C" {

int uvb_version(void) { return 110; }

int uvb_set_knob(int knob, int value) {
  if (knob < 0 || knob >= UVB_KNOB_COUNT) return fail(UVB_ERR_INVALID, "unknown knob %d", knob);
  g_knobs[knob] = value;
  return UVB_OK;
}

int uvb_get_knob(int knob) {
  if (knob < 0 || knob >= UVB_KNOB_COUNT) return fail(UVB_ERR_INVALID, "unknown knob %d", knob);
  return g_knobs[knob];
}

void uvb_debug_fmha_timeline(void* device_buffer) {
  g_timeline = static_cast<unsigned long long*>(device_buffer);
}

#ifdef UVB_FMHA_PROFILE
// lab builds only (not declared in the public header): per-CTA wait counters, 16 x u64 per CTA
void uvb_debug_fmha_profile(void* device_buffer) { g_prof = static_cast<unsigned long long*>(device_buffer); }
#endif

int64_t uvb_fmha_workspace_bytes(void) {
  int sms = 0;
  if (sm_count(&sms) != UVB_OK) return -1;
  return static_cast<int64_t>(fmha_ws_bytes(sms));
}

const char* uvb_last_error(void) { return g_err; }

int uvb_qk_norm_rope(const void* q_in, const void* k_in, int in_dtype, const float* wq,
                     const float* wk, const float* cos_sin, const float* row_scale,
                     const float* pre_bias, void* q_out, void* k_out, int B, int L, int N,
                     const int32_t* grid_fhw, int tok_offset, float eps, int hpg, int64_t out_sb,
                     int64_t out_sl, int64_t out_sg, void* stream) {
  return uvb_qk_norm_rope_sp(q_in, k_in, in_dtype, wq, wk, cos_sin, row_scale, pre_bias, q_out, k_out,
                             nullptr, nullptr, 0, B, L, N, grid_fhw, tok_offset, eps, hpg, out_sb, out_sl,
                             out_sg, stream);
}

int uvb_qk_norm_rope_sp(const void* q_in, const void* k_in, int in_dtype, const float* wq,
                        const float* wk, const float* cos_sin, const float* row_scale,
                        const float* pre_bias, void* q_out, void* k_out, void* const* q_peers,
                        void* const* k_peers, int n_peers, int B, int L, int N,
                        const int32_t* grid_fhw, int tok_offset, float eps, int hpg, int64_t out_sb,
                        int64_t out_sl, int64_t out_sg, void* stream) {
  if (q_in == nullptr && k_in == nullptr) return fail(UVB_ERR_INVALID, "q_in and k_in are both null");
  if (n_peers < 0 || n_peers > uvb::kMaxPeers) return fail(UVB_ERR_INVALID, "n_peers=%d", n_peers);
  if (n_peers > 0) {
    if (hpg <= 0 || N != hpg * n_peers) return fail(UVB_ERR_INVALID, "N=%d != hpg=%d * n_peers=%d", N, hpg, n_peers);
    if ((q_in != nullptr && q_peers == nullptr) || (k_in != nullptr && k_peers == nullptr))
      return fail(UVB_ERR_INVALID, "missing peer pointer list");
  } else if ((q_in != nullptr && q_out == nullptr) || (k_in != nullptr && k_out == nullptr)) {
    return fail(UVB_ERR_INVALID, "missing output pointer");
  }
  if (B <= 0 || L <= 0 || N <= 0) return fail(UVB_ERR_INVALID, "bad shape B=%d L=%d N=%d", B, L, N);
  if (in_dtype != UVB_BF16 && in_dtype != UVB_F32) return fail(UVB_ERR_INVALID, "bad in_dtype %d", in_dtype);
  if (hpg <= 0 || N % hpg != 0) return fail(UVB_ERR_INVALID, "hpg=%d must divide N=%d", hpg, N);
  if (cos_sin != nullptr && grid_fhw == nullptr)
    return fail(UVB_ERR_INVALID, "grid_fhw is required when cos_sin is given");
  if (cos_sin != nullptr && B > uvb::kMaxBatchGrid)
    return fail(UVB_ERR_UNSUPPORTED, "B=%d > %d samples per call with RoPE", B, uvb::kMaxBatchGrid);
  if (out_sb % 8 != 0 || out_sl % 8 != 0 || out_sg % 8 != 0)
    return fail(UVB_ERR_INVALID, "output strides must be multiples of 8 elements");
  if ((row_scale != nullptr) != (pre_bias != nullptr) && row_scale != nullptr)
    return fail(UVB_ERR_INVALID, "row_scale requires pre_bias");
  int rc = check_device();
  if (rc != UVB_OK) return rc;

  uvb::NormRopeParams p;
  memset(&p, 0, sizeof(p));
  p.q_in = q_in;
  p.k_in = k_in;
  p.wq = wq;
  p.wk = wk;
  p.cos_sin = reinterpret_cast<const float2*>(cos_sin);
  p.row_scale = row_scale;
  p.pre_bias = pre_bias;
  p.q_out = static_cast<__nv_bfloat16*>(q_out);
  p.k_out = static_cast<__nv_bfloat16*>(k_out);
  p.B = B;
  p.L = L;
  p.N = N;
  if (cos_sin != nullptr) {
    for (int b = 0; b < B; ++b) {
      for (int i = 0; i < 3; ++i) {
        const int g = grid_fhw[3 * b + i];
        if (g <= 0 || g > 1024) return fail(UVB_ERR_INVALID, "grid size %d out of (0, 1024]", g);
        p.grid[b][i] = g;
      }
    }
  }
  p.tok_offset = tok_offset;
  p.eps = eps;
  p.hpg = hpg;
  p.out_sb = out_sb;
  p.out_sl = out_sl;
  p.out_sg = out_sg;
  p.n_peers = n_peers;
  for (int j = 0; j < n_peers; ++j) {
    p.q_peer[j] = q_in != nullptr ? static_cast<__nv_bfloat16*>(q_peers[j]) : nullptr;
    p.k_peer[j] = k_in != nullptr ? static_cast<__nv_bfloat16*>(k_peers[j]) : nullptr;
    if ((q_in != nullptr && p.q_peer[j] == nullptr) || (k_in != nullptr && p.k_peer[j] == nullptr))
      return fail(UVB_ERR_INVALID, "null peer pointer %d", j);
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  return in_dtype == UVB_BF16 ? launch_norm_rope<__nv_bfloat16>(p, st) : launch_norm_rope<float>(p, st);
}

int uvb_head_scatter_bf16(const void* v_in, void* v_out, int B, int L, int N, int hpg,
                          int64_t out_sb, int64_t out_sl, int64_t out_sg, void* stream) {
  return uvb_head_scatter_sp(v_in, v_out, nullptr, 0, B, L, N, hpg, out_sb, out_sl, out_sg, stream);
}

int uvb_head_scatter_sp(const void* v_in, void* v_out, void* const* peers, int n_peers, int B, int L,
                        int N, int hpg, int64_t out_sb, int64_t out_sl, int64_t out_sg, void* stream) {
  if (v_in == nullptr || (v_out == nullptr && n_peers == 0)) return fail(UVB_ERR_INVALID, "null pointer");
  if (B <= 0 || L <= 0 || N <= 0 || hpg <= 0 || N % hpg != 0)
    return fail(UVB_ERR_INVALID, "bad shape B=%d L=%d N=%d hpg=%d", B, L, N, hpg);
  if (n_peers < 0 || n_peers > uvb::kMaxPeers || (n_peers > 0 && (peers == nullptr || N != hpg * n_peers)))
    return fail(UVB_ERR_INVALID, "bad peer list (n_peers=%d)", n_peers);
  int rc = check_device();
  if (rc != UVB_OK) return rc;
  uvb::HeadScatterParams p;
  p.in = static_cast<const __nv_bfloat16*>(v_in);
  p.out = static_cast<__nv_bfloat16*>(v_out);
  p.B = B;
  p.L = L;
  p.N = N;
  p.hpg = hpg;
  p.out_sb = out_sb;
  p.out_sl = out_sl;
  p.out_sg = out_sg;
  p.n_peers = n_peers;
  for (int j = 0; j < n_peers; ++j) {
    if (peers[j] == nullptr) return fail(UVB_ERR_INVALID, "null peer pointer %d", j);
    p.peer[j] = static_cast<__nv_bfloat16*>(peers[j]);
  }
  const long long total = static_cast<long long>(B) * L * N * 16;
  long long blocks = (total + 255) / 256;
  if (blocks > 148LL * 16) blocks = 148LL * 16;
  uvb::head_scatter_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
  UVB_CUDA(cudaGetLastError());
  return UVB_OK;
}

int uvb_fmha_fwd_bf16(const void* q, const void* k, const void* v, void* o, const int32_t* k_lens,
                      int B, int Lq, int Lk, int N, const int64_t* q_strides,
                      const int64_t* k_strides, const int64_t* v_strides, const int64_t* o_strides,
                      float scale, void* workspace, int64_t workspace_bytes, void* stream) {
  return launch_fmha<false>(q, k, v, o, k_lens, nullptr, nullptr, nullptr, B, Lq, Lk, N, q_strides,
                            k_strides, v_strides, o_strides, scale, workspace, workspace_bytes, stream);
}

int uvb_block_glue(const float* x_in, const void* y, const float* gate, float* x_out, const float* ln_w,
                   const float* ln_b, const float* scale, const float* shift, void* h_out, int B, int L,
                   int dim, int64_t mod_sb, int64_t mod_sl, const int32_t* mod_index, float eps, int ln_round_bf16,
                   void* stream) {
  if (x_in == nullptr) return fail(UVB_ERR_INVALID, "x_in is null");
  if (B <= 0 || L <= 0 || dim <= 0) return fail(UVB_ERR_INVALID, "bad shape B=%d L=%d dim=%d", B, L, dim);
  if (y == nullptr && h_out == nullptr) return fail(UVB_ERR_INVALID, "nothing to do (y and h_out are both null)");
  if (y != nullptr && x_out == nullptr) return fail(UVB_ERR_INVALID, "x_out is required with y");
  if (y == nullptr && gate != nullptr) return fail(UVB_ERR_INVALID, "gate without y");
  if ((scale == nullptr) != (shift == nullptr)) return fail(UVB_ERR_INVALID, "scale and shift go together");
  if (mod_sb % 4 != 0 || mod_sl % 4 != 0) return fail(UVB_ERR_INVALID, "modulation strides must be multiples of 4");
  for (const void* q : {static_cast<const void*>(x_in), y, static_cast<const void*>(gate),
                        static_cast<const void*>(x_out), static_cast<const void*>(ln_w),
                        static_cast<const void*>(ln_b), static_cast<const void*>(scale),
                        static_cast<const void*>(shift), static_cast<const void*>(h_out)}) {
    if ((reinterpret_cast<uintptr_t>(q) & 15) != 0) return fail(UVB_ERR_INVALID, "pointers must be 16-byte aligned");
  }
  int rc = check_device();
  if (rc != UVB_OK) return rc;
  uvb::BlockGlueParams p;
  memset(&p, 0, sizeof(p));
  p.x_in = x_in;
  p.y = static_cast<const __nv_bfloat16*>(y);
  p.gate = gate;
  p.x_out = x_out;
  p.ln_w = ln_w;
  p.ln_b = ln_b;
  p.scale = scale;
  p.shift = shift;
  p.h_out = static_cast<__nv_bfloat16*>(h_out);
  p.rows = static_cast<long long>(B) * L;
  p.L = L;
  p.dim = dim;
  p.mod_sb = mod_sb;
  p.mod_sl = mod_sl;
  p.mod_index = mod_index;
  p.eps = eps;
  p.ln_round_bf16 = ln_round_bf16 != 0 ? 1 : 0;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const dim3 block(uvb::kGlueWarps * 32);
  auto blocks = [&](int wpr) {
    const long long per_cta = uvb::kGlueWarps / wpr;
    return static_cast<unsigned>((p.rows + per_cta - 1) / per_cta);
  };
  switch (dim) {   // 256 * VPL * WPR
    case 256: uvb::block_glue_kernel<1, 1><<<blocks(1), block, 0, st>>>(p); break;
    case 512: uvb::block_glue_kernel<2, 1><<<blocks(1), block, 0, st>>>(p); break;
    case 1024: uvb::block_glue_kernel<4, 1><<<blocks(1), block, 0, st>>>(p); break;
    case 1536: uvb::block_glue_kernel<3, 2><<<blocks(2), block, 0, st>>>(p); break;
    case 2048: uvb::block_glue_kernel<4, 2><<<blocks(2), block, 0, st>>>(p); break;
    case 3072: uvb::block_glue_kernel<3, 4><<<blocks(4), block, 0, st>>>(p); break;
    case 4096: uvb::block_glue_kernel<4, 4><<<blocks(4), block, 0, st>>>(p); break;
    case 5120: uvb::block_glue_kernel<5, 4><<<blocks(4), block, 0, st>>>(p); break;
    default:
      return fail(UVB_ERR_UNSUPPORTED, "block glue supports dim in {256, 512, 1024, 1536, 2048, 3072, 4096, 5120}, got %d", dim);
  }
  UVB_CUDA(cudaGetLastError());
  return UVB_OK;
}

int uvb_fmha_fwd_sp_bf16(const void* q, const void* k, const void* v, void* const* o_peers, int n_peers,
                         int head_offset, int total_heads, const int32_t* k_lens, int B, int Lq, int Lk,
                         int N, const int64_t* q_strides, const int64_t* k_strides,
                         const int64_t* v_strides, float scale, void* workspace,
                         int64_t workspace_bytes, void* stream) {
  if (n_peers <= 0) return fail(UVB_ERR_INVALID, "n_peers must be positive");
  return launch_fmha<false>(q, k, v, nullptr, k_lens, nullptr, nullptr, nullptr, B, Lq, Lk, N, q_strides,
                            k_strides, v_strides, nullptr, scale, workspace, workspace_bytes, stream,
                            o_peers, n_peers, head_offset, total_heads);
}

// ------------------------------- exchange buffers, IPC, hand-off flags -------------------------------
int uvb_sp_buffer_alloc(int64_t bytes, void** dev_ptr) {
  if (bytes <= 0 || dev_ptr == nullptr) return fail(UVB_ERR_INVALID, "bad arguments");
  UVB_CUDA(cudaMalloc(dev_ptr, static_cast<size_t>(bytes)));
  UVB_CUDA(cudaMemset(*dev_ptr, 0, static_cast<size_t>(bytes)));
  UVB_CUDA(cudaDeviceSynchronize());
  return UVB_OK;
}

int uvb_sp_buffer_free(void* dev_ptr) {
  UVB_CUDA(cudaFree(dev_ptr));
  return UVB_OK;
}

int uvb_sp_ipc_export(void* dev_ptr, void* handle64) {
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  if (dev_ptr == nullptr || handle64 == nullptr) return fail(UVB_ERR_INVALID, "null pointer");
  cudaIpcMemHandle_t h;
  UVB_CUDA(cudaIpcGetMemHandle(&h, dev_ptr));
  memcpy(handle64, &h, sizeof(h));
  return UVB_OK;
}

int uvb_sp_ipc_import(const void* handle64, void** peer_ptr) {
  if (handle64 == nullptr || peer_ptr == nullptr) return fail(UVB_ERR_INVALID, "null pointer");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, sizeof(h));
  UVB_CUDA(cudaIpcOpenMemHandle(peer_ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return UVB_OK;
}

int uvb_sp_ipc_close(void* peer_ptr) {
  UVB_CUDA(cudaIpcCloseMemHandle(peer_ptr));
  return UVB_OK;
}

int uvb_sp_signal(void* const* flag_ptrs, int n, uint32_t value, void* stream) {
  if (flag_ptrs == nullptr || n <= 0 || n > uvb::kMaxPeers) return fail(UVB_ERR_INVALID, "bad flag list");
  uvb::SpSignalParams p;
  memset(&p, 0, sizeof(p));
  for (int i = 0; i < n; ++i) {
    if (flag_ptrs[i] == nullptr) return fail(UVB_ERR_INVALID, "null flag pointer %d", i);
    p.flag[i] = static_cast<uint32_t*>(flag_ptrs[i]);
  }
  p.n = n;
  p.value = value;
  uvb::sp_signal_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(p);
  UVB_CUDA(cudaGetLastError());
  return UVB_OK;
}

int uvb_sp_signal_wait(void* const* flag_ptrs, int n, uint32_t value, const void* wait_flags, void* stream) {
  if (flag_ptrs == nullptr || wait_flags == nullptr || n <= 0 || n > uvb::kMaxPeers)
    return fail(UVB_ERR_INVALID, "bad flag list");
  uvb::SpSignalParams p;
  memset(&p, 0, sizeof(p));
  for (int i = 0; i < n; ++i) {
    if (flag_ptrs[i] == nullptr) return fail(UVB_ERR_INVALID, "null flag pointer %d", i);
    p.flag[i] = static_cast<uint32_t*>(flag_ptrs[i]);
  }
  p.n = n;
  p.value = value;
  const int secs = g_knobs[UVB_KNOB_SP_WAIT_TIMEOUT_S];
  const unsigned long long timeout_ns = secs <= 0 ? 0ull : static_cast<unsigned long long>(secs) * 1000000000ull;
  uvb::sp_signal_wait_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(
      p, static_cast<const uint32_t*>(wait_flags), timeout_ns);
  UVB_CUDA(cudaGetLastError());
  return UVB_OK;
}

int uvb_sp_wait(const void* flags, int n, uint32_t value, void* stream) {
  if (flags == nullptr || n <= 0 || n > 32) return fail(UVB_ERR_INVALID, "bad flag array");
  const int secs = g_knobs[UVB_KNOB_SP_WAIT_TIMEOUT_S];
  const unsigned long long timeout_ns = secs <= 0 ? 0ull : static_cast<unsigned long long>(secs) * 1000000000ull;
  uvb::sp_wait_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const uint32_t*>(flags), n, value,
                                                                       timeout_ns);
  UVB_CUDA(cudaGetLastError());
  return UVB_OK;
}

int uvb_xattn_fwd_bf16(const void* q, const void* k, const void* v, void* o, const int32_t* k_lens,
                       const float* key_logit_scale, const float* key_pv_weight,
                       const float* out_bias, int B, int Lq, int Lk, int N,
                       const int64_t* q_strides, const int64_t* k_strides,
                       const int64_t* v_strides, const int64_t* o_strides, float scale,
                       void* workspace, int64_t workspace_bytes, void* stream) {
  return launch_fmha<true>(q, k, v, o, k_lens, key_logit_scale, key_pv_weight, out_bias, B, Lq, Lk,
                           N, q_strides, k_strides, v_strides, o_strides, scale, workspace,
                           workspace_bytes, stream);
}

int uvb_linear_bf16(const void* x, const void* w, const float* bias, void* y, int M, int N, int K, int64_t ldx,
                    int64_t ldw, int64_t ldy, int act, void* stream) {
  if (y == nullptr) return fail(UVB_ERR_INVALID, "null matrix pointer");
  return linear_impl(x, w, bias, y, M, N, K, ldx, ldw, ldy, act, stream, GemmPeers());
}

int uvb_linear_bf16_sp(const void* x, const void* w, const float* bias, void* const* y_peers, int n_peers, int M, int N,
                       int K, int64_t ldx, int64_t ldw, int64_t ld_peer, int act, void* stream) {
  if (y_peers == nullptr || n_peers <= 0 || n_peers > uvb::kMaxPeers) return fail(UVB_ERR_INVALID, "bad peer list");
  if (N % n_peers != 0 || (N / n_peers) % 64 != 0)
    return fail(UVB_ERR_INVALID, "N=%d must split into %d column groups of a multiple of 64 columns", N, n_peers);
  for (int j = 0; j < n_peers; ++j) {
    if (y_peers[j] == nullptr) return fail(UVB_ERR_INVALID, "null peer pointer %d", j);
  }
  GemmPeers peers;
  peers.ptrs = y_peers;
  peers.n = n_peers;
  peers.ld = ld_peer;
  return linear_impl(x, w, bias, nullptr, M, N, K, ldx, ldw, 0, act, stream, peers);
}

int uvb_unipc_step(const float* cond, const float* uncond, const float* x, const float* last, const float* m0,
                   const float* m1, float* m_out, float* xc_out, float* x_next, int64_t n,
                   const uvb_unipc_coef* coef, void* stream) {
  if (cond == nullptr || x == nullptr || m_out == nullptr || xc_out == nullptr || x_next == nullptr || coef == nullptr)
    return fail(UVB_ERR_INVALID, "null pointer");
  if (n <= 0) return fail(UVB_ERR_INVALID, "bad element count %lld", static_cast<long long>(n));
  if (coef->corrector_order < 0 || coef->corrector_order > 2 || coef->predictor_order < 1 || coef->predictor_order > 2)
    return fail(UVB_ERR_UNSUPPORTED, "UniPC orders (corrector %d, predictor %d) beyond solver_order 2",
                coef->corrector_order, coef->predictor_order);
  if (coef->corrector_order > 0 && (last == nullptr || m0 == nullptr)) return fail(UVB_ERR_INVALID, "corrector needs last and m0");
  if (coef->corrector_order == 2 && m1 == nullptr) return fail(UVB_ERR_INVALID, "second-order corrector needs m1");
  if (coef->predictor_order == 2 && m0 == nullptr) return fail(UVB_ERR_INVALID, "second-order predictor needs m0");
  for (const void* q : {static_cast<const void*>(cond), static_cast<const void*>(uncond), static_cast<const void*>(x),
                        static_cast<const void*>(last), static_cast<const void*>(m0), static_cast<const void*>(m1),
                        static_cast<const void*>(m_out), static_cast<const void*>(xc_out), static_cast<const void*>(x_next)}) {
    if ((reinterpret_cast<uintptr_t>(q) & 15) != 0) return fail(UVB_ERR_INVALID, "pointers must be 16-byte aligned");
  }
  int rc = check_device();
  if (rc != UVB_OK) return rc;
  int sms = 0;
  if ((rc = sm_count(&sms)) != UVB_OK) return rc;
  uvb::SamplerStepParams p;
  memset(&p, 0, sizeof(p));
  p.cond = cond;
  p.uncond = uncond;
  p.x = x;
  p.last = last;
  p.m0 = m0;
  p.m1 = m1;
  p.m_out = m_out;
  p.xc_out = xc_out;
  p.x_next = x_next;
  p.n = n;
  p.guide = coef->guide_scale;
  p.sigma = coef->sigma;
  p.corr_order = coef->corrector_order;
  p.c_a = coef->c_a;
  p.c_b = coef->c_b;
  p.c_ab = coef->c_ab;
  p.c_rk = coef->c_rk;
  p.c_rho0 = coef->c_rho0;
  p.c_rho_last = coef->c_rho_last;
  p.pred_order = coef->predictor_order;
  p.p_a = coef->p_a;
  p.p_b = coef->p_b;
  p.p_ab = coef->p_ab;
  p.p_rk = coef->p_rk;
  p.p_rho0 = coef->p_rho0;
  p.history_bf16 = coef->history_bf16 != 0 ? 1 : 0;
  long long blocks = (n / 4 + 255) / 256;
  if (blocks < 1) blocks = 1;
  if (blocks > 8LL * sms) blocks = 8LL * sms;       // grid-stride: a whole number of CTAs per SM
  uvb::sampler_step_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
  UVB_CUDA(cudaGetLastError());
  return UVB_OK;
}

}  // extern "C"
