"""CPU oracle for the Wan DiT attention hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A restatement, in plain PyTorch CPU ops, of the algorithm the reference (AIGeeksGroup/UniVid,
mounted at /root/reference while this repo is built) executes for the path named in BASELINE.json.
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import this module; the product package ``univid_b200`` never does.

Pinning: the reference ships no tests, golden vectors or fixtures for this path (SURVEY.md sec. 4,
8c), so the oracle is pinned against outputs of the reference's own source executed in the build
container: ``tests/golden/make_golden.py`` loads /root/reference/models/wan/utils/modules/
{model,attention}.py, models/wan/distributed/*.py and the ``Wan22ContextWrapper`` class of
models/model_pipeline.py (by path, unmodified), runs them on seeded inputs and freezes
inputs+outputs under ``tests/golden/``; ``tests/test_oracle_golden.py`` checks every function here
against those files (bit-exact where the reference is deterministic fp32/fp64 arithmetic).

Every function cites the reference file:line it follows (paths relative to /root/reference).
Rounding points are reproduced explicitly instead of through ``torch.autocast``: "bf16 mode" below is
what the reference computes under ``torch.autocast(dtype=torch.bfloat16)`` (textimage2video.py:330),
"fp32 mode" is the same math with fp32 parameters and no autocast (the gold of the north-star
tolerance).
"""
import math

import torch
import torch.nn.functional as F

HEAD_DIM = 128


# ------------------------------------------------------------------------------------------------
# RoPE tables -- model.py:27-35 (rope_params) and model.py:398-405 (three bands concatenated)
# ------------------------------------------------------------------------------------------------
def rope_angles(max_seq_len, dim, theta=10000.0):
    """Angles pos * theta^(-2j/dim), float64, shape [max_seq_len, dim // 2] (model.py:30-33)."""
    assert dim % 2 == 0
    inv = 1.0 / torch.pow(torch.tensor(float(theta), dtype=torch.float64),
                          torch.arange(0, dim, 2, dtype=torch.float64) / dim)
    return torch.outer(torch.arange(max_seq_len, dtype=torch.float64), inv)


def rope_params(max_seq_len, dim, theta=10000.0):
    """Complex128 table exp(i * angle) (model.py:34).  torch.polar, like the reference: its sincos
    differs from torch.cos/torch.sin by one ulp on some entries."""
    ang = rope_angles(max_seq_len, dim, theta)
    return torch.polar(torch.ones_like(ang), ang)


def band_sizes(head_dim=HEAD_DIM):
    """Complex frequencies per (temporal, height, width) band: c-2(c//3), c//3, c//3 (model.py:43)."""
    c = head_dim // 2
    return c - 2 * (c // 3), c // 3, c // 3


def make_freqs(head_dim=HEAD_DIM, max_pos=1024):
    """The [1024, head_dim/2] table WanModel.__init__ builds (model.py:398-405)."""
    d = head_dim
    return torch.cat([
        rope_params(max_pos, d - 4 * (d // 6)),
        rope_params(max_pos, 2 * (d // 6)),
        rope_params(max_pos, 2 * (d // 6)),
    ], dim=1)


# ------------------------------------------------------------------------------------------------
# WanRMSNorm -- model.py:69-85
# ------------------------------------------------------------------------------------------------
def rms_norm(x, weight, eps):
    """(x.float() * rsqrt(mean(x.float()^2, -1) + eps)).type_as(x) * weight (model.py:82-85).

    The reduction runs over the whole model width, the normalised value is rounded to x's dtype
    and the product with the fp32 weight promotes back to fp32.
    """
    xf = x.float()
    inv = torch.rsqrt(xf.pow(2).mean(dim=-1, keepdim=True) + eps)
    return (xf * inv).to(x.dtype) * weight


# ------------------------------------------------------------------------------------------------
# rope_apply -- model.py:38-66 ; sequence-parallel variant -- sequence_parallel.py:23-61
# ------------------------------------------------------------------------------------------------
def _token_rotations(f, h, w, freqs):
    """Per-token complex multipliers [f*h*w, head_dim/2]: token t -> (t//(h*w), (t//w)%h, t%w), the
    first band indexed by the frame, the second by the row, the third by the column (model.py:53-58)."""
    nf, nh, nw = band_sizes(freqs.shape[1] * 2)
    ff, fh, fw = freqs.split([nf, nh, nw], dim=1)
    t = torch.arange(f * h * w)
    return torch.cat([ff[t // (h * w)], fh[(t // w) % h], fw[t % w]], dim=1)


def _rotate(x, rot):
    """x [S, N, D] any float dtype, rot [S, D/2] complex128 -> float64 [S, N, D]; interleaved pairs
    (x[2j], x[2j+1]) are multiplied by rot[:, j] in float64 (model.py:51-61)."""
    xd = x.to(torch.float64)
    xe, xo = xd[..., 0::2], xd[..., 1::2]
    c, s = rot.real.unsqueeze(1), rot.imag.unsqueeze(1)
    out = torch.empty_like(xd)
    out[..., 0::2] = xe * c - xo * s
    out[..., 1::2] = xe * s + xo * c
    return out


def rope_apply(x, grid_sizes, freqs):
    """x [B, L, N, D]; tokens >= f*h*w are passed through (model.py:62); returns fp32 (model.py:66)."""
    out = []
    for i, (f, h, w) in enumerate(grid_sizes.tolist()):
        n_tok = f * h * w
        rot = _token_rotations(f, h, w, freqs)
        xi = torch.cat([_rotate(x[i, :n_tok], rot), x[i, n_tok:].to(torch.float64)])
        out.append(xi)
    return torch.stack(out).float()


def sp_rope_apply(x, grid_sizes, freqs, rank, world):
    """Rank slice of the rotation table, padded with 1+0j up to s*world (sequence_parallel.py:46-55).
    x [B, s, N, D] is the local shard; tokens beyond s are passed through (:56)."""
    s = x.size(1)
    out = []
    for i, (f, h, w) in enumerate(grid_sizes.tolist()):
        rot = _token_rotations(f, h, w, freqs)
        pad = s * world - rot.shape[0]
        if pad > 0:
            rot = torch.cat([rot, torch.ones(pad, rot.shape[1], dtype=rot.dtype)])
        rot = rot[rank * s:(rank + 1) * s]
        xi = torch.cat([_rotate(x[i, :s], rot), x[i, s:].to(torch.float64)])
        out.append(xi)
    return torch.stack(out).float()


# ------------------------------------------------------------------------------------------------
# attention core -- attention.py:24-130 (flash route) and attention.py:133-179 (SDPA route)
# ------------------------------------------------------------------------------------------------
def attention_sdpa(q, k, v, dtype=torch.bfloat16):
    """The reference's torch-SDPA route (attention.py:164-179): [B, L, N, D] -> transpose -> cast to
    `dtype` -> F.scaled_dot_product_attention (no mask, default scale D^-0.5) -> transpose back.
    Returns `dtype`.  k_lens / q_lens are ignored on this route (attention.py:165-168)."""
    qt, kt, vt = (u.transpose(1, 2).to(dtype) for u in (q, k, v))
    out = F.scaled_dot_product_attention(qt, kt, vt, attn_mask=None, is_causal=False, dropout_p=0.0)
    return out.transpose(1, 2).contiguous()


def attention_varlen(q, k, v, k_lens=None, softmax_scale=None, compute_dtype=torch.bfloat16,
                     key_logit_scale=None, key_pv_weight=None, out_bias=None):
    """Semantics of the flash route (attention.py:56-130): q, k, v are cast to the half dtype, keys
    at positions >= k_lens[b] are excluded (attention.py:72-80), softmax(q k^T * scale) v with fp32
    accumulation, result cast back to q's input dtype (attention.py:130).  Evaluated here in fp32 on
    the rounded operands.  The optional per-key modifiers describe the fused cross-attention
    variant (include/univid_b200.h, uvb_xattn_fwd_bf16); they are not part of the reference call."""
    out_dtype = q.dtype
    b, lq, n, d = q.shape
    lk = k.shape[1]
    scale = d ** -0.5 if softmax_scale is None else softmax_scale
    qf, kf, vf = (u.to(compute_dtype).float().transpose(1, 2) for u in (q, k, v))
    logits = torch.matmul(qf, kf.transpose(-1, -2)) * scale            # [B, N, Lq, Lk]
    if key_logit_scale is not None:
        logits = logits * key_logit_scale.float().view(1, 1, 1, lk)
    if k_lens is not None:
        pos = torch.arange(lk).view(1, 1, 1, lk)
        logits = logits.masked_fill(pos >= k_lens.view(b, 1, 1, 1), float("-inf"))
    p = torch.softmax(logits, dim=-1)
    p = torch.nan_to_num(p, nan=0.0)                                   # k_len == 0 -> zeros
    if key_pv_weight is not None:
        p = p * key_pv_weight.float().view(1, 1, 1, lk)
    out = torch.matmul(p, vf).transpose(1, 2)                          # [B, Lq, N, D]
    if out_bias is not None:
        # the bias stands for sum_j p_j b = b * sum_j p_j: a row without any key (k_len == 0) gets none
        has_keys = (p.sum(dim=-1) > 0).transpose(1, 2).unsqueeze(-1).float() if key_pv_weight is None else \
            (torch.nan_to_num(torch.softmax(logits, dim=-1), nan=0.0).sum(dim=-1) > 0).transpose(1, 2).unsqueeze(-1).float()
        out = out + out_bias.float().view(1, 1, n, d) * has_keys
    return out.to(compute_dtype).to(out_dtype)


# ------------------------------------------------------------------------------------------------
# modules -- model.py:101-180
# ------------------------------------------------------------------------------------------------
def _linear(x, w, b, bf16):
    """nn.Linear under autocast: operands cast to bf16, bf16 result (fp32 accumulate inside)."""
    if bf16:
        return F.linear(x.to(torch.bfloat16), w.to(torch.bfloat16),
                        None if b is None else b.to(torch.bfloat16))
    return F.linear(x.float(), w.float(), None if b is None else b.float())


def self_attention_qkv(x, prm, grid_sizes, freqs, num_heads, eps=1e-6, bf16=True,
                       sp_rank=None, sp_world=None):
    """q, k, v as they enter the attention core (model.py:137-147): q/k = rope(norm(linear(x)))
    in fp32, v = linear(x).  With sp_rank/sp_world the sequence-parallel rotation is used
    (sequence_parallel.py:155-163)."""
    b, s = x.shape[:2]
    d = x.shape[2] // num_heads
    q = rms_norm(_linear(x, prm["q.weight"], prm["q.bias"], bf16), prm["norm_q.weight"], eps)
    k = rms_norm(_linear(x, prm["k.weight"], prm["k.bias"], bf16), prm["norm_k.weight"], eps)
    v = _linear(x, prm["v.weight"], prm["v.bias"], bf16)
    q, k, v = (u.view(b, s, num_heads, d) for u in (q, k, v))
    if sp_rank is None:
        return rope_apply(q, grid_sizes, freqs), rope_apply(k, grid_sizes, freqs), v
    return (sp_rope_apply(q, grid_sizes, freqs, sp_rank, sp_world),
            sp_rope_apply(k, grid_sizes, freqs, sp_rank, sp_world), v)


def self_attention(x, prm, seq_lens, grid_sizes, freqs, num_heads, eps=1e-6, bf16=True,
                   route="sdpa"):
    """WanSelfAttention.forward (model.py:126-155).  route="sdpa" is the reference's torch-SDPA path
    (k_lens ignored, bf16 result, attention.py:164-179); route="varlen" honours k_lens=seq_lens
    like the flash path (attention.py:72-80, result cast to q's fp32, :130)."""
    q, k, v = self_attention_qkv(x, prm, grid_sizes, freqs, num_heads, eps, bf16)
    cd = torch.bfloat16 if bf16 else torch.float32
    if route == "sdpa":
        a = attention_sdpa(q, k, v, dtype=cd)
    else:
        a = attention_varlen(q, k, v, k_lens=seq_lens, compute_dtype=cd)
    return _linear(a.flatten(2), prm["o.weight"], prm["o.bias"], bf16)


def cross_attention(x, context, prm, num_heads, context_lens=None, eps=1e-6, bf16=True,
                    route="sdpa"):
    """WanCrossAttention.forward (model.py:160-180): no RoPE, K/V from the context."""
    b = x.size(0)
    d = x.shape[2] // num_heads
    q = rms_norm(_linear(x, prm["q.weight"], prm["q.bias"], bf16), prm["norm_q.weight"], eps)
    k = rms_norm(_linear(context, prm["k.weight"], prm["k.bias"], bf16), prm["norm_k.weight"], eps)
    v = _linear(context, prm["v.weight"], prm["v.bias"], bf16)
    q, k, v = q.view(b, -1, num_heads, d), k.view(b, -1, num_heads, d), v.view(b, -1, num_heads, d)
    cd = torch.bfloat16 if bf16 else torch.float32
    if route == "sdpa":
        a = attention_sdpa(q, k, v, dtype=cd)
    else:
        a = attention_varlen(q, k, v, k_lens=context_lens, compute_dtype=cd)
    return _linear(a.flatten(2), prm["o.weight"], prm["o.bias"], bf16)


def animate_cross_attention(x, context, prm, num_heads, context_lens=None, eps=1e-6, bf16=True,
                            route="sdpa", use_img_emb=True, img_tokens=257):
    """WanAnimateCrossAttention.forward (models/wan/utils/modules/animate/model_animate.py:111-146): the first 257
    context rows are CLIP image tokens with their own k_img / v_img / norm_k_img (:103-106, :117-119, :129-132); the
    same queries attend to the image keys (all of them, k_lens=None) and to the text keys (k_lens=context_lens),
    the two results are added in the attention output dtype (:141-144) and projected by `o`.  Under the SDPA route
    both results are bf16 and so is their sum; under the flash route they are fp32 (attention.py:130)."""
    b = x.size(0)
    d = x.shape[2] // num_heads
    cd = torch.bfloat16 if bf16 else torch.float32

    def attend(q, k, v, lens):
        if route == "sdpa":
            return attention_sdpa(q, k, v, dtype=cd)
        return attention_varlen(q, k, v, k_lens=lens, compute_dtype=cd)

    if use_img_emb:
        context_img, context = context[:, :img_tokens], context[:, img_tokens:]
    q = rms_norm(_linear(x, prm["q.weight"], prm["q.bias"], bf16), prm["norm_q.weight"], eps).view(b, -1, num_heads, d)
    k = rms_norm(_linear(context, prm["k.weight"], prm["k.bias"], bf16), prm["norm_k.weight"], eps).view(b, -1, num_heads, d)
    v = _linear(context, prm["v.weight"], prm["v.bias"], bf16).view(b, -1, num_heads, d)
    a = attend(q, k, v, context_lens).flatten(2)
    if use_img_emb:
        k_img = rms_norm(_linear(context_img, prm["k_img.weight"], prm["k_img.bias"], bf16),
                         prm["norm_k_img.weight"], eps).view(b, -1, num_heads, d)
        v_img = _linear(context_img, prm["v_img.weight"], prm["v_img.bias"], bf16).view(b, -1, num_heads, d)
        a = a + attend(q, k_img, v_img, None).flatten(2)
    return _linear(a, prm["o.weight"], prm["o.bias"], bf16)


# ------------------------------------------------------------------------------------------------
# WanLayerNorm -- model.py:88-98 ; WanAttentionBlock -- model.py:183-259
# ------------------------------------------------------------------------------------------------
def layer_norm(x, eps, weight=None, bias=None):
    """nn.LayerNorm.forward(x.float()).type_as(x) (model.py:92-98): statistics and affine in fp32, result
    rounded back to x's dtype."""
    return F.layer_norm(x.float(), (x.shape[-1],), None if weight is None else weight.float(),
                        None if bias is None else bias.float(), eps).to(x.dtype)


def attention_block(x, e, prm, seq_lens, grid_sizes, freqs, context, context_lens, num_heads, eps=1e-6,
                    bf16=True, route="sdpa", cross_attn_norm=True):
    """WanAttentionBlock.forward (model.py:219-259).  prm holds the block's state dict ('modulation',
    'self_attn.*', 'norm3.*', 'cross_attn.*', 'ffn.0.*', 'ffn.2.*').  e: fp32 [B, L1, 6, C].
    Dtype flow under bf16 autocast: the modulation arithmetic and the residual updates run in fp32
    (:239-240, :246-247, :256-257), every nn.Linear rounds its input to bf16 and returns bf16, GELU(tanh)
    acts on the bf16 output of ffn.0."""
    assert e.dtype == torch.float32
    sub = lambda pre: {k[len(pre):]: v for k, v in prm.items() if k.startswith(pre)}
    m = (prm["modulation"].float().unsqueeze(0) + e).chunk(6, dim=2)                     # :239
    m = [u.squeeze(2) for u in m]
    h = layer_norm(x, eps).float() * (1 + m[1]) + m[0]                                    # :244
    y = self_attention(h, sub("self_attn."), seq_lens, grid_sizes, freqs, num_heads, eps, bf16, route)
    x = x + y * m[2]                                                                       # :247
    n3 = layer_norm(x, eps, prm["norm3.weight"], prm["norm3.bias"]) if cross_attn_norm else x
    x = x + cross_attention(n3, context, sub("cross_attn."), num_heads, context_lens, eps, bf16, route)   # :252
    h = layer_norm(x, eps).float() * (1 + m[4]) + m[3]                                    # :254-255
    f = _linear(h, prm["ffn.0.weight"], prm["ffn.0.bias"], bf16)
    f = F.gelu(f, approximate="tanh")
    y = _linear(f, prm["ffn.2.weight"], prm["ffn.2.bias"], bf16)
    return x + y * m[5]                                                                    # :257


# ------------------------------------------------------------------------------------------------
# WanModel.forward -- model.py:410-497 (patchify + pad, time / text embeddings, blocks, head, unpatchify)
# ------------------------------------------------------------------------------------------------
def sinusoidal_embedding_1d(dim, position):
    """cos | sin of position * 10000^(-i / half), float64 (model.py:13-24)."""
    half = dim // 2
    pos = position.to(torch.float64)
    ang = torch.outer(pos, torch.pow(10000, -torch.arange(half).to(pos).div(half)))
    return torch.cat([torch.cos(ang), torch.sin(ang)], dim=1)


def dit_forward(lat, t, ctx, prm, seq_len, num_heads, num_layers, dim, freq_dim, text_len, out_dim,
                patch_size=(1, 2, 2), eps=1e-6, bf16=False, route="sdpa"):
    """WanModel.forward (model.py:410-497) for model_type 't2v' / 'ti2v' (no y).  prm = the model's state dict.
    lat: list of [C_in, F, H, W]; t: [B] (B == 1 only, like the reference's expand at :461) or [B, seq_len];
    ctx: list of [L, text_dim].  fp32 evaluation by default (bf16=False: the 'fp32 reference' of the tolerance)."""
    freqs = make_freqs(dim // num_heads)
    xs = [F.conv3d(u.unsqueeze(0), prm["patch_embedding.weight"], prm["patch_embedding.bias"], stride=patch_size)
          for u in lat]                                                                              # :445
    grid_sizes = torch.stack([torch.tensor(u.shape[2:], dtype=torch.long) for u in xs])               # :446-447
    xs = [u.flatten(2).transpose(1, 2) for u in xs]
    seq_lens = torch.tensor([u.size(1) for u in xs], dtype=torch.long)
    assert int(seq_lens.max()) <= seq_len
    x = torch.cat([torch.cat([u, u.new_zeros(1, seq_len - u.size(1), u.size(2))], dim=1) for u in xs])  # :451-454
    if t.dim() == 1:
        t = t.expand(t.size(0), seq_len)                                                              # :460-461
    bt = t.size(0)
    emb = sinusoidal_embedding_1d(freq_dim, t.flatten()).unflatten(0, (bt, seq_len)).float()          # :465-467
    e = F.linear(F.silu(F.linear(emb, prm["time_embedding.0.weight"], prm["time_embedding.0.bias"])),
                 prm["time_embedding.2.weight"], prm["time_embedding.2.bias"])
    e0 = F.linear(F.silu(e), prm["time_projection.1.weight"], prm["time_projection.1.bias"]).unflatten(2, (6, dim))
    c = torch.stack([torch.cat([u, u.new_zeros(text_len - u.size(0), u.size(1))]) for u in ctx])      # :473-478
    c = _linear(F.gelu(_linear(c, prm["text_embedding.0.weight"], prm["text_embedding.0.bias"], bf16),
                       approximate="tanh"), prm["text_embedding.2.weight"], prm["text_embedding.2.bias"], bf16)
    for i in range(num_layers):
        pre = f"blocks.{i}."
        blk = {k[len(pre):]: v for k, v in prm.items() if k.startswith(pre)}
        x = attention_block(x, e0, blk, seq_lens, grid_sizes, freqs, c, None, num_heads, eps, bf16, route,
                            cross_attn_norm="norm3.weight" in blk)
    m = (prm["head.modulation"].unsqueeze(0) + e.unsqueeze(2)).chunk(2, dim=2)                         # Head.forward :286
    h = layer_norm(x, eps) * (1 + m[1].squeeze(2)) + m[0].squeeze(2)
    y = F.linear(h.float(), prm["head.head.weight"], prm["head.head.bias"])                            # fp32 autocast region
    out = []
    for u, v in zip(y, grid_sizes.tolist()):                                                           # unpatchify :516-521
        u = u[:math.prod(v)].view(*v, *patch_size, out_dim)
        u = torch.einsum("fhwpqrc->cfphqwr", u)
        out.append(u.reshape(out_dim, *[i * j for i, j in zip(v, patch_size)]).float())
    return out


def init_block_params(dim, ffn_dim, generator, realistic_bias=True):
    """State dict of one WanAttentionBlock (cross_attn_norm=True) with the init rules of SURVEY.md sec. 8d."""
    prm = {"modulation": torch.randn(1, 6, dim, generator=generator) / dim ** 0.5}
    for pre in ("self_attn.", "cross_attn."):
        for k, v in init_attention_params(dim, generator, realistic_bias).items():
            prm[pre + k] = v
    prm["norm3.weight"] = 1 + 0.1 * torch.randn(dim, generator=generator)
    prm["norm3.bias"] = 0.05 * torch.randn(dim, generator=generator)
    for name, (o, i) in (("ffn.0", (ffn_dim, dim)), ("ffn.2", (dim, ffn_dim))):
        bound = math.sqrt(6.0 / (o + i))
        prm[f"{name}.weight"] = (torch.rand(o, i, generator=generator) * 2 - 1) * bound
        prm[f"{name}.bias"] = torch.randn(o, generator=generator) * 0.02
    return prm


# ------------------------------------------------------------------------------------------------
# Temperature Modality Alignment -- model_pipeline.py:1699-1735 (schedule), :1756-1803 (hook)
# ------------------------------------------------------------------------------------------------
def text_weight(call_index, total_sampling_steps=50, transition_ratio=0.4, w_max=1.3, w_min=1.0,
                schedule="cosine", use_dynamic_text_weight=True):
    """_calculate_text_weight (model_pipeline.py:1699-1735).  `call_index` counts DiT forwards
    (hooked_dit_forward, :1856-1866), not sampler steps."""
    if not use_dynamic_text_weight:
        return 1.0
    transition = int(total_sampling_steps * transition_ratio)
    if call_index >= transition:
        return w_min
    progress = call_index / max(transition, 1)
    if schedule == "linear":
        return w_max - (w_max - w_min) * progress
    if schedule == "cosine":
        return w_min + (w_max - w_min) * (1 + math.cos(math.pi * progress)) / 2
    if schedule == "exponential":
        return w_min + (w_max - w_min) * math.exp(-5 * progress)
    return 1.0


def text_len_for(context, bagel_sequence_length=128):
    """text_len = min(bagel_sequence_length, seq_len // 2) (model_pipeline.py:1789)."""
    seq_len = context.shape[1] if context.dim() > 1 else context.shape[0]
    return min(bagel_sequence_length, seq_len // 2)


def weight_context(context, w, bagel_sequence_length=128):
    """The hook's context[:, :text_len] *= w through a ones mask (model_pipeline.py:1789-1797)."""
    tl = text_len_for(context, bagel_sequence_length)
    mask = torch.ones_like(context)
    if context.dim() == 3:
        mask[:, :tl, :] *= w
    else:
        mask[:tl, :] *= w
    return context * mask


def cross_attention_text_weighted(x, context, prm, num_heads, w, bagel_sequence_length=128,
                                  eps=1e-6, bf16=True, route="sdpa"):
    """WanCrossAttention.forward as entered through hooked_forward when use_bagel_context is armed
    and w != 1 (model_pipeline.py:1756-1803)."""
    ctx = weight_context(context, w, bagel_sequence_length) if w != 1.0 else context
    return cross_attention(x, ctx, prm, num_heads, None, eps, bf16, route)


# ------------------------------------------------------------------------------------------------
# Ulysses -- distributed/util.py:21-31, distributed/ulysses.py:9-47 (single-process emulation)
# ------------------------------------------------------------------------------------------------
def all_to_all_emulated(shards, scatter_dim, gather_dim):
    """What util.all_to_all leaves on every rank: rank r receives chunk r (along scatter_dim) of each
    rank's tensor and concatenates them in rank order along gather_dim (util.py:27-30)."""
    world = len(shards)
    chunks = [list(u.chunk(world, dim=scatter_dim)) for u in shards]
    return [torch.cat([chunks[src][dst] for src in range(world)], dim=gather_dim).contiguous()
            for dst in range(world)]


def ulysses_attention_emulated(q_shards, k_shards, v_shards, seq_lens, compute_dtype=torch.bfloat16):
    """distributed_attention (ulysses.py:30-47) for all ranks at once: heads<->sequence exchange,
    attention on [B, L, N/p, D] with k_lens=seq_lens, inverse exchange."""
    q = all_to_all_emulated(q_shards, 2, 1)
    k = all_to_all_emulated(k_shards, 2, 1)
    v = all_to_all_emulated(v_shards, 2, 1)
    x = [attention_varlen(a, b, c, k_lens=seq_lens, compute_dtype=compute_dtype)
         for a, b, c in zip(q, k, v)]
    return all_to_all_emulated(x, 1, 2)


def sp_self_attention_emulated(x, prm, seq_lens, grid_sizes, freqs, num_heads, world, eps=1e-6,
                               bf16=True):
    """sp_attn_forward on every rank (sequence_parallel.py:147-176): x [B, L, dim] with L % world == 0
    is chunked over tokens (sp_dit_forward :119), q/k/v are cast to bf16 before the exchange
    (:165-168).  Returns the list of per-rank outputs [B, L/world, dim]."""
    xs = x.chunk(world, dim=1)
    qs, ks, vs = [], [], []
    for r in range(world):
        q, k, v = self_attention_qkv(xs[r], prm, grid_sizes, freqs, num_heads, eps, bf16, r, world)
        half = torch.bfloat16
        qs.append(q.to(half)), ks.append(k.to(half)), vs.append(v.to(half))
    outs = ulysses_attention_emulated(qs, ks, vs, seq_lens)
    return [_linear(o.flatten(2), prm["o.weight"], prm["o.bias"], bf16) for o in outs]


# ------------------------------------------------------------------------------------------------
# synthetic parameters
# ------------------------------------------------------------------------------------------------
def init_attention_params(dim, generator, realistic_bias=False):
    """Weights of one WanSelfAttention/WanCrossAttention: xavier-uniform linears with zero bias and
    unit norm weights (WanModel.init_weights, model.py:530-534, :75); with realistic_bias the second
    weight set of SURVEY.md sec. 8d (bias ~ N(0, 0.02), norm weight ~ 1 + N(0, 0.1))."""
    prm = {}
    bound = math.sqrt(6.0 / (dim + dim))
    for name in ("q", "k", "v", "o"):
        prm[f"{name}.weight"] = (torch.rand(dim, dim, generator=generator) * 2 - 1) * bound
        prm[f"{name}.bias"] = (torch.randn(dim, generator=generator) * 0.02 if realistic_bias
                               else torch.zeros(dim))
    for name in ("norm_q", "norm_k"):
        prm[f"{name}.weight"] = (1 + 0.1 * torch.randn(dim, generator=generator) if realistic_bias
                                 else torch.ones(dim))
    return prm
