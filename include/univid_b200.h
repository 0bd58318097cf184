/* univid_b200 -- C ABI of the B200-native Wan DiT attention hot path.
 *
 * The reference (AIGeeksGroup/UniVid) is pure Python: its "FFI" for this path is the set of
 * third-party kernel calls made from models/wan/utils/modules/{attention,model}.py and
 * models/wan/distributed/{ulysses,util,sequence_parallel}.py.  Each entry point below replaces
 * one of those call sites (cited per function).  Conventions:
 *   - every pointer named q/k/v/o/..._in/_out is a DEVICE pointer; `grid_fhw` and the stride
 *     arrays are HOST pointers read during the call;
 *   - `stream` is a cudaStream_t passed as void* (torch.cuda.current_stream().cuda_stream);
 *   - functions return 0 on success, a negative uvb_status otherwise; the message is available
 *     from uvb_last_error() (thread-local); nothing throws, allocates device memory or
 *     synchronises the device;
 *   - there is no CPU fallback: without an sm_100 device the compute calls fail with
 *     UVB_ERR_CUDA / UVB_ERR_UNSUPPORTED.
 */
#ifndef UNIVID_B200_H_
#define UNIVID_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum uvb_status {
  UVB_OK = 0,
  UVB_ERR_INVALID = -1,     /* bad argument (null pointer, shape, alignment) */
  UVB_ERR_UNSUPPORTED = -2, /* head_dim != 128, device is not sm_100, ... */
  UVB_ERR_CUDA = -3         /* CUDA runtime / driver error; see uvb_last_error() */
};

enum uvb_dtype { UVB_BF16 = 0, UVB_F32 = 1 };

/* Library ABI version (major*100 + minor). */
int uvb_version(void);

/* Message of the last failing call on this thread ("" if none). */
const char* uvb_last_error(void);

/* Explicit, process-wide tuning knobs.  The library never reads environment variables: kernel selection
 * changes only through this call (benchmarks and A/B scripts use it; a product never has to).  The
 * defaults are the shipped configuration.  uvb_set_knob returns 0 or UVB_ERR_INVALID; uvb_get_knob
 * returns the value (>= 0) or UVB_ERR_INVALID.  (No reference counterpart: the reference selects its
 * attention backend with module-level flags, attention.py:4-17.) */
enum uvb_knob {
  UVB_KNOB_FMHA_PAIR = 0,     /* 1: CTA-pair (cta_group::2) attention kernel for Lk > 2048 (default); 0: single CTAs
                                 (2: pairs with early S release -- lab builds only, -DUVB_LAB_VARIANTS) */
  UVB_KNOB_FMHA_SPLIT = 1,    /* 1: stream-K split of the remainder query blocks (default); 0: never split */
  UVB_KNOB_GEMM_CTAS = 2,     /* 2: CTA-pair GEMM tiles (default); 1: single-CTA tiles */
  UVB_KNOB_GEMM_BN = 3,       /* 0: tile width chosen per problem (default); 192 | 256: pinned */
  UVB_KNOB_GEMM_SMALL = 4,    /* 1: one-wave 128x64 tiles for small problems (default); 0: off */
  UVB_KNOB_PROLOGUE_PAIR = 5, /* q AND k given: 2 = streaming kernel, persistent CTAs + bulk-copy ring (default, widths
                                 1536 / 3072 / 5120); 1 = token-pair kernel; 0 = one row per warp group */
  UVB_KNOB_FMHA_POLY = 6,     /* lab builds only: one exp2 pair in every n (2, 3, 4) on the FMA pipe; 0 (default, shipped) = MUFU only */
  UVB_KNOB_SP_WAIT_TIMEOUT_S = 7, /* seconds uvb_sp_wait spins for a peer's hand-off flag before it traps (default 600,
                                     like an NCCL collective under torch.distributed); 0 = wait for ever */
  UVB_KNOB_XATTN_PAIR = 8,    /* 1: CTA-pair variant of the short-key (Lk <= 2048, cross-attention) kernel (default); 0: single CTAs */
  UVB_KNOB_COUNT = 9
};
int uvb_set_knob(int knob, int value);
int uvb_get_knob(int knob);

/* Fused WanRMSNorm(q), WanRMSNorm(k) over the full width dim = N*128, followed by the 3-D RoPE.
 * Replaces: WanRMSNorm.forward (model.py:77-85) x2 + rope_apply (model.py:38-66) x2 of
 * WanSelfAttention.forward (model.py:137-147); with cos_sin == NULL it is the norm-only prologue of
 * WanCrossAttention.forward (model.py:170-171); tok_offset > 0 is the rank slice of the
 * sequence-parallel rope_apply (distributed/sequence_parallel.py:23-61).
 *
 *   q_in, k_in   [B, L, N*128] in `in_dtype` (either may be NULL to skip that tensor)
 *   wq, wk       [N*128] fp32 RMSNorm weights; NULL = no normalisation for that tensor (qk_norm=False,
 *                nn.Identity at model.py:123-124), rotation only
 *   cos_sin      [1024, 64, 2] fp32 (cos, sin) of the reference `freqs` table (model.py:398-405),
 *                or NULL for no rotation
 *   row_scale    [L] fp32 or NULL; pre_bias [N*128] fp32 or NULL: when pre_bias != NULL the k row is
 *                first mapped to row_scale[l]*k + pre_bias (text-weighted context, model_pipeline.py:1789-1797
 *                folded through the k projection)
 *   q_out, k_out bf16; element (b, l, n, d) is written at
 *                b*out_sb + l*out_sl + (n / hpg)*out_sg + (n % hpg)*128 + d     (element strides)
 *                hpg = N, out_sg = 0 gives plain [B, L, N, 128]; hpg = N/p gives the Ulysses send
 *                layout [p][B][L][N/p][128] (util.py:27 chunk(scatter_dim=2) fused into the store)
 *   grid_fhw     HOST int32 [B, 3] token grid (f, h, w) per sample (B <= 8), NULL iff cos_sin NULL
 *   tok_offset   global index of local token 0; tokens >= f*h*w are normalised but not rotated
 */
int uvb_qk_norm_rope(const void* q_in, const void* k_in, int in_dtype, const float* wq,
                     const float* wk, const float* cos_sin, const float* row_scale,
                     const float* pre_bias, void* q_out, void* k_out, int B, int L, int N,
                     const int32_t* grid_fhw, int tok_offset, float eps, int hpg, int64_t out_sb,
                     int64_t out_sl, int64_t out_sg, void* stream);

/* Pure copy of v [B, L, N, 128] bf16 into the same grouped layout as above.
 * Replaces: the chunk(...).contiguous() pack of all_to_all (distributed/util.py:27). */
int uvb_head_scatter_bf16(const void* v_in, void* v_out, int B, int L, int N, int hpg,
                          int64_t out_sb, int64_t out_sl, int64_t out_sg, void* stream);

/* softmax(q k^T * scale) v, non-causal, head_dim 128, bf16, fp32 accumulation.
 * Replaces: flash_attn_varlen_func / flash_attn_interface / F.scaled_dot_product_attention at
 * attention.py:96,113,175 as reached from flash_attention() (attention.py:24) and attention()
 * (attention.py:133).
 *
 *   q [B, Lq, N, 128], k/v [B, Lk, N, 128], o [B, Lq, N, 128]; *_strides are HOST int64[3]
 *   = (batch, token, head) strides in elements, NULL = contiguous.  All strides must be multiples
 *   of 8 elements and every base pointer 16-byte aligned.
 *   k_lens  DEVICE int32 [B] or NULL: keys >= k_lens[b] are masked (attention.py:72-80).
 *   scale   softmax scale (reference default: 128^-0.5).
 *   workspace, workspace_bytes   see uvb_fmha_workspace_bytes(); may be NULL / 0.
 */
int uvb_fmha_fwd_bf16(const void* q, const void* k, const void* v, void* o, const int32_t* k_lens,
                      int B, int Lq, int Lk, int N, const int64_t* q_strides,
                      const int64_t* k_strides, const int64_t* v_strides, const int64_t* o_strides,
                      float scale, void* workspace, int64_t workspace_bytes, void* stream);

/* Size in bytes of the scratch buffer the attention kernels use to split query blocks over the key
 * axis when the number of 256-row query blocks is not a multiple of the SM count (the kernel is
 * persistent, one CTA per SM; see csrc/fmha_fwd_sm100.cuh).  The caller owns the buffer: DEVICE memory,
 * 256-byte aligned, ZERO-FILLED ONCE after allocation (the kernels leave it zero-filled where it
 * matters), used by one stream at a time.  `workspace == NULL` is allowed everywhere and selects the
 * schedule that never splits a query block (a partial last wave).  Returns -1 without a CUDA device. */
int64_t uvb_fmha_workspace_bytes(void);

/* ---------------------------------------------------------------------------------------------------------
 * Fused Ulysses exchange over NVLink peer memory (replaces: all_to_all() x4 in distributed_attention,
 * models/wan/distributed/ulysses.py:32-46 / util.py:21-31, and the grouped p2p it issues through NCCL).
 * One process per GPU.  Every rank owns an exchange buffer (uvb_sp_buffer_alloc), exports it
 * (uvb_sp_ipc_export), and maps every peer's (uvb_sp_ipc_import).  The producers then store q/k/v head
 * groups straight into the destination rank's buffer (uvb_qk_norm_rope_sp, uvb_head_scatter_sp), the
 * attention kernel TMA-stores each output tile into the buffer of the rank that owns its tokens
 * (uvb_fmha_fwd_sp_bf16), and two flag kernels order producers and consumers across GPUs: uvb_sp_signal
 * after a producer, uvb_sp_wait before the consumer, with one monotonically increasing value per exchange.
 * No NCCL call and no pack/unpack pass remains on the data path.
 *
 * uvb_sp_buffer_alloc is the only call in this library that allocates device memory (cudaMalloc: IPC needs
 * a whole allocation) and synchronises; it is a set-up call, not on the hot path.  Handles are 64 bytes.
 */
int uvb_sp_buffer_alloc(int64_t bytes, void** dev_ptr);     /* zero-filled */
int uvb_sp_buffer_free(void* dev_ptr);
int uvb_sp_ipc_export(void* dev_ptr, void* handle64);
int uvb_sp_ipc_import(const void* handle64, void** peer_ptr);
int uvb_sp_ipc_close(void* peer_ptr);
/* flag_ptrs: HOST array of n DEVICE pointers (this rank's flag word inside each peer's buffer): after all
 * work queued on `stream` so far, *flag_ptrs[i] = value (system-scope release). */
int uvb_sp_signal(void* const* flag_ptrs, int n, uint32_t value, void* stream);
/* flags: DEVICE uint32[n] in this rank's buffer: work queued on `stream` afterwards starts only once every
 * flags[i] - value >= 0 (wrap-safe).  Traps after ~10 s instead of hanging the GPU. */
int uvb_sp_wait(const void* flags, int n, uint32_t value, void* stream);
/* uvb_sp_signal followed by uvb_sp_wait(wait_flags, n, value) in one launch. */
int uvb_sp_signal_wait(void* const* flag_ptrs, int n, uint32_t value, const void* wait_flags, void* stream);

/* uvb_qk_norm_rope with peer stores: with n_peers = p > 0 (N == hpg * p) head group j of element (b, l) goes
 * to {q,k}_peers[j] + b*out_sb + l*out_sl + (n % hpg)*128 + d; {q,k}_peers are HOST arrays of p DEVICE
 * pointers, q_out/k_out/out_sg are ignored.  n_peers == 0 is uvb_qk_norm_rope. */
int uvb_qk_norm_rope_sp(const void* q_in, const void* k_in, int in_dtype, const float* wq,
                        const float* wk, const float* cos_sin, const float* row_scale,
                        const float* pre_bias, void* q_out, void* k_out, void* const* q_peers,
                        void* const* k_peers, int n_peers, int B, int L, int N,
                        const int32_t* grid_fhw, int tok_offset, float eps, int hpg, int64_t out_sb,
                        int64_t out_sl, int64_t out_sg, void* stream);
int uvb_head_scatter_sp(const void* v_in, void* v_out, void* const* peers, int n_peers, int B, int L,
                        int N, int hpg, int64_t out_sb, int64_t out_sl, int64_t out_sg, void* stream);
/* uvb_fmha_fwd_bf16 on a head shard [B, Lq, N, 128] whose output rows [j*Lq/p, (j+1)*Lq/p) are stored into
 * rank j's buffer o_peers[j] = [B, Lq/p, total_heads, 128] at heads [head_offset, head_offset + N). */
int uvb_fmha_fwd_sp_bf16(const void* q, const void* k, const void* v, void* const* o_peers, int n_peers,
                         int head_offset, int total_heads, const int32_t* k_lens, int B, int Lq, int Lk,
                         int N, const int64_t* q_strides, const int64_t* k_strides,
                         const int64_t* v_strides, float scale, void* workspace,
                         int64_t workspace_bytes, void* stream);

/* Fused elementwise glue of a WanAttentionBlock around the attention calls.
 * Replaces, per call, the eager chain of WanAttentionBlock.forward (model.py:244-257): the gated residual
 * `x = x + y * e[k]` (:247 / :252 / :257), WanLayerNorm (:88-98: norm1 / norm2 without affine, norm3 with),
 * the adaLN modulation `norm(x).float() * (1 + e[k]) + e[k']` (:244 / :255) and the autocast cast of the
 * result to bf16 at the consuming nn.Linear.
 *
 *   x' = x_in + y * gate        if y != NULL (y bf16 [B, L, dim]; gate NULL = 1); x' is stored to x_out
 *                               (fp32, may alias x_in)
 *   h  = LayerNorm(x') * ln_w + ln_b  (ln_w / ln_b NULL = no affine), eps inside the rsqrt
 *   h  = h * (1 + scale) + shift      (both NULL = no modulation)   -> h_out bf16 [B, L, dim]; h_out NULL
 *                               skips the LayerNorm part (residual update only)
 *   gate / scale / shift are fp32 views into the modulation tensor (modulation + e).chunk(6): element
 *   (b, l, c) at ptr + b*mod_sb + l*mod_sl + c (element strides; mod_sl = 0 broadcasts one row per sample).
 *   mod_index: DEVICE int32 [B*L] or NULL.  When given, token (b, l) reads modulation row mod_index[b*L + l] instead
 *   of row l: the reference expands the timestep to one value per token (textimage2video.py:372-377, model.py:460-468)
 *   and materialises [B, L, 6, dim]; with few distinct timesteps one row per distinct value is enough.
 *   ln_round_bf16 != 0: the LayerNorm result (after the affine) is rounded to bf16 before the modulation.  This is
 *   WanLayerNorm's `.type_as(x)` (model.py:98) for a bf16 x -- the first block of the DiT receives the bf16 output of
 *   the patch embedding (model.py:447, under autocast); the caller passes that x widened to fp32 (exact).
 *   dim in {256, 512, 1024, 1536, 2048, 3072, 4096, 5120}; all pointers 16-byte aligned.
 */
int uvb_block_glue(const float* x_in, const void* y, const float* gate, float* x_out, const float* ln_w,
                   const float* ln_b, const float* scale, const float* shift, void* h_out, int B, int L,
                   int dim, int64_t mod_sb, int64_t mod_sl, const int32_t* mod_index, float eps, int ln_round_bf16,
                   void* stream);

/* Diagnostics: when set to a DEVICE buffer of (number of SMs) x 32 uint64, every attention launch records
 * per CTA {smid, start ns, end ns of each piece of work (up to 30)} (%globaltimer).  NULL (default) = off. */
void uvb_debug_fmha_timeline(void* device_buffer);

/* Cross-attention with per-key modifiers fused (UniVid "Temperature Modality Alignment").
 * Replaces: WanCrossAttention.forward's flash_attention call (model.py:175) as entered through
 * Wan22ContextWrapper.hooked_forward (model_pipeline.py:1756-1803).
 *   key_logit_scale  DEVICE fp32 [ceil128(Lk)] or NULL: logits[:, j] *= key_logit_scale[j]
 *                    (a per-key temperature on the text keys)
 *   key_pv_weight    DEVICE fp32 [ceil128(Lk)] or NULL: probabilities are multiplied by w[j] AFTER
 *                    normalisation (out = sum_j p_j w_j v_j / sum_j p_j); with v = bias-free value
 *                    projection this equals scaling context rows by w_j (SURVEY.md A7')
 *   out_bias         DEVICE fp32 [N*128] or NULL, added to the normalised output (the value bias)
 */
int uvb_xattn_fwd_bf16(const void* q, const void* k, const void* v, void* o, const int32_t* k_lens,
                       const float* key_logit_scale, const float* key_pv_weight,
                       const float* out_bias, int B, int Lq, int Lk, int N,
                       const int64_t* q_strides, const int64_t* k_strides,
                       const int64_t* v_strides, const int64_t* o_strides, float scale,
                       void* workspace, int64_t workspace_bytes, void* stream);

/* y = act(x . w^T + bias): the nn.Linear layers of the DiT block around the attention core, under bf16 autocast
 * (SURVEY.md sec. 8f, rank 2).  Replaces: F.linear as called by WanSelfAttention / WanCrossAttention q, k, v, o
 * (model.py:119-122, :137-139, :152-154, :170-173, :178-179) and by the block's ffn (model.py:212-214, :256),
 * where act = UVB_ACT_GELU_TANH also replaces nn.GELU(approximate='tanh') (model.py:213) on the first ffn layer.
 *
 *   x [M, K] bf16 (leading dimension ldx elements), w [N, K] bf16 (nn.Linear.weight layout, ldw),
 *   bias [N] fp32 or NULL (autocast rounds the bias to bf16 before the add: pass the rounded values),
 *   y [M, N] bf16 (ldy).  N, K and the leading dimensions are multiples of 8; pointers 16-byte aligned.
 *   Accumulation in fp32; the biased sum is rounded to bf16 (the Linear's output) and, with GELU, the activation
 *   is evaluated in fp32 on that rounded value and rounded again -- the reference's rounding points.
 */
enum uvb_act { UVB_ACT_NONE = 0, UVB_ACT_GELU_TANH = 1 };
int uvb_linear_bf16(const void* x, const void* w, const float* bias, void* y, int M, int N, int K, int64_t ldx,
                    int64_t ldw, int64_t ldy, int act, void* stream);

/* The same GEMM with the output scattered by column groups: columns [j * N/n_peers, (j + 1) * N/n_peers) of the result
 * are TMA-stored as a [M, N/n_peers] matrix (leading dimension ld_peer) at y_peers[j].  Replaces: the v projection
 * (model.py:139) followed by the head-chunk pack of all_to_all (distributed/util.py:27) in the sequence-parallel
 * self-attention (distributed/sequence_parallel.py:147-176): y_peers[j] is this rank's slot inside rank j's exchange
 * buffer (mapped over NVLink), so the projection's epilogue IS the first half of the Ulysses exchange of v.
 * N/n_peers must be a multiple of 64; y_peers is a HOST array of n_peers (<= 8) device pointers. */
int uvb_linear_bf16_sp(const void* x, const void* w, const float* bias, void* const* y_peers, int n_peers, int M, int N,
                       int K, int64_t ldx, int64_t ldw, int64_t ld_peer, int act, void* stream);

/* Sampler update between two DiT forwards, one fused elementwise pass.
 * Replaces: the classifier-free-guidance combine `uncond + g * (cond - uncond)` (models/wan/textimage2video.py:385-386)
 * and FlowUniPCMultistepScheduler.step (models/wan/utils/fm_solvers_unipc.py:657-741: convert_model_output :320-323,
 * UniC corrector :549-628, UniP predictor :395-486) for solver_order <= 2, predict_x0, flow_prediction.
 * The host computes the scalar coefficients in the reference's fp32 arithmetic (uvb_unipc_coef); the kernel applies
 *   v    = uncond ? uncond + guide * (cond - uncond) : cond
 *   m_t  = x - sigma * v
 *   x_c  = corrector_order ? c_a * last - c_b * m0 - c_ab * ([c_rho0 * (m1 - m0) / c_rk +] c_rho_last * (m_t - m0)) : x
 *   next = p_a * x_c - p_b * m_t [- p_ab * (p_rho0 * (m0 - m_t) / p_rk)]          (bracket: order 2)
 * with every operation individually rounded (bit-identical to the fp32 reference chain).  history_bf16 != 0
 * reproduces what the same code computes inside torch.amp.autocast('cuda', bfloat16), where the product runs it
 * (textimage2video.py:330-331): torch.einsum over the history terms (fm_solvers_unipc.py:471, :614) then runs in bf16 --
 * rho and D1 are rounded to bf16, so is their product, and in the predictor the product with the fp32 scalar alpha_t * B_h
 * is rounded to bf16 once more.
 *   cond, uncond (or NULL), x, last, m0, m1: DEVICE fp32 [n], 16-byte aligned; last/m0/m1 may be NULL when the orders
 *   do not need them.  m_out, xc_out, x_next: DEVICE fp32 [n] outputs (must not alias the inputs of later steps the
 *   caller still needs).  coef: HOST pointer.
 */
typedef struct uvb_unipc_coef {
  float guide_scale;
  float sigma;
  int32_t corrector_order;   /* 0 = no corrector, 1 or 2 */
  float c_a, c_b, c_ab, c_rk, c_rho0, c_rho_last;
  int32_t predictor_order;   /* 1 or 2 */
  float p_a, p_b, p_ab, p_rk, p_rho0;
  int32_t history_bf16;      /* 0: fp32 chain; 1: the bf16 roundings of the history einsum under bf16 autocast */
} uvb_unipc_coef;
int uvb_unipc_step(const float* cond, const float* uncond, const float* x, const float* last, const float* m0,
                   const float* m1, float* m_out, float* xc_out, float* x_next, int64_t n,
                   const uvb_unipc_coef* coef, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* UNIVID_B200_H_ */
