"""Drop-in for the attention part of models/wan/utils/modules/model.py (logical path
wan/modules/model.py) backed by the sm_100a kernels of libunivid_b200.so.

Hot path (BASELINE.json north_star), reference lines in brackets:
  WanRMSNorm            [model.py:69-85]   q/k norm over the full model width
  rope_params/rope_apply[model.py:27-66]   3-D rotary embedding, fp64 in the reference
  WanSelfAttention      [model.py:101-155] fused norm+rope prologue kernel -> tcgen05 flash attention
  WanCrossAttention     [model.py:158-180] norm-only prologue -> flash attention over 512 context keys;
                                           opt-in fused text weighting (model_pipeline.py:1756-1803)
Everything else in this file (WanLayerNorm, WanAttentionBlock, Head, WanModel) is the plain-PyTorch
harness around the hot path that the denoise-step measurement needs; it
keeps the reference's parameter names so reference checkpoints load, and its module/attribute names
so the product's monkey patches (Wan22ContextWrapper, sp_attn_forward, LoRA targets) keep working.
"""
import math
import os
import weakref

import torch
import torch.nn as nn

from ... import _ext
from .attention import _k_lens_arg, flash_attention

__all__ = ['WanModel']


def sinusoidal_embedding_1d(dim, position):
    """[len(position), dim] float64: cos(p * 10000^(-i/half)) | sin(...)  (model.py:14-24)."""
    assert dim % 2 == 0
    half = dim // 2
    pos = position.to(torch.float64)
    inv = torch.pow(10000, -torch.arange(half, dtype=torch.float64, device=pos.device) / half)
    ang = pos[:, None] * inv[None, :]
    return torch.cat([ang.cos(), ang.sin()], dim=1)


@torch.amp.autocast('cuda', enabled=False)
def rope_params(max_seq_len, dim, theta=10000):
    """complex128 [max_seq_len, dim/2] table exp(i * p * theta^(-2j/dim))  (model.py:27-35)."""
    assert dim % 2 == 0
    expo = torch.arange(0, dim, 2, dtype=torch.float64) / dim
    ang = torch.outer(torch.arange(max_seq_len, dtype=torch.float64), 1.0 / torch.pow(theta, expo))
    return torch.polar(torch.ones_like(ang), ang)


_COS_SIN_CACHE = {}


def _cos_sin_table(freqs, device):
    """fp32 [1024, 64, 2] (cos, sin) device table for the prologue kernel, cached per freqs tensor.  The key holds a
    weak reference to the tensor (checked on every hit), so an address recycled for another table never aliases."""
    key = (id(freqs), freqs.data_ptr(), tuple(freqs.shape), str(device))
    hit = _COS_SIN_CACHE.get(key)
    tab = hit[1] if hit is not None and hit[0]() is freqs and hit[2] == _version_of(freqs) else None
    if tab is None:
        if freqs.dim() != 2 or freqs.shape[1] != 64 or not freqs.is_complex():
            raise NotImplementedError('the RoPE kernel expects the [M, 64] complex table of head_dim 128')
        f = freqs
        if f.shape[0] < 1024:
            f = torch.cat([f, torch.ones(1024 - f.shape[0], 64, dtype=f.dtype, device=f.device)])
        tab = torch.stack([f.real[:1024], f.imag[:1024]], dim=-1).to(device=device, dtype=torch.float32).contiguous()
        if len(_COS_SIN_CACHE) > 16:
            _COS_SIN_CACHE.clear()
        _COS_SIN_CACHE[key] = (weakref.ref(freqs), tab, _version_of(freqs))
    return tab


@torch.amp.autocast('cuda', enabled=False)
def rope_apply(x, grid_sizes, freqs):
    """Stand-alone 3-D RoPE with the reference's semantics (model.py:38-66): x [B, L, N, D] any float
    dtype, rotation evaluated in float64, tokens >= f*h*w untouched, fp32 result.  API-compatibility
    entry point (model_animate.py and sequence_parallel.py import it); WanSelfAttention below does
    not call it -- its rotation is fused into the prologue kernel."""
    half = x.size(3) // 2
    widths = [half - 2 * (half // 3), half // 3, half // 3]
    f_t, f_h, f_w = freqs.split(widths, dim=1)
    out = x.to(torch.float64).clone()
    for i, (f, h, w) in enumerate(grid_sizes.tolist()):
        n_tok = f * h * w
        t = torch.arange(n_tok, device=x.device)
        rot = torch.cat([f_t[t // (h * w)], f_h[(t // w) % h], f_w[t % w]], dim=1).to(x.device)
        c, s = rot.real[:, None, :], rot.imag[:, None, :]
        xe, xo = out[i, :n_tok, :, 0::2].clone(), out[i, :n_tok, :, 1::2].clone()
        out[i, :n_tok, :, 0::2] = xe * c - xo * s
        out[i, :n_tok, :, 1::2] = xe * s + xo * c
    return out.float()


class WanRMSNorm(nn.Module):
    """y = (x.float() * rsqrt(mean(x^2) + eps)).type_as(x) * weight  (model.py:69-85).  Used stand-alone
    this module runs as PyTorch ops; inside WanSelfAttention / WanCrossAttention only `.weight` and
    `.eps` are read and the normalisation happens in the fused prologue kernel."""

    def __init__(self, dim, eps=1e-5):
        super().__init__()
        self.dim = dim
        self.eps = eps
        self.weight = nn.Parameter(torch.ones(dim))

    def forward(self, x):
        xf = x.float()
        return (xf * torch.rsqrt(xf.square().mean(dim=-1, keepdim=True) + self.eps)).type_as(x) * self.weight


class WanLayerNorm(nn.LayerNorm):
    """fp32 LayerNorm returning the input dtype (model.py:88-98)."""

    def __init__(self, dim, eps=1e-6, elementwise_affine=False):
        super().__init__(dim, elementwise_affine=elementwise_affine, eps=eps)

    def forward(self, x):
        return super().forward(x.float()).type_as(x)


def _norm_weight(norm):
    """(weight, eps, pre) for a q/k norm module: a WanRMSNorm is folded into the kernel; nn.Identity
    (qk_norm=False, model.py:123-124) means rotation only; anything else is applied as a module first."""
    if isinstance(norm, WanRMSNorm):
        return norm.weight, norm.eps, None
    if isinstance(norm, nn.Identity):
        return None, 0.0, None
    return None, 0.0, norm


def _proj_for_kernel(t):
    if t.dtype not in (torch.bfloat16, torch.float32):
        raise NotImplementedError(f'univid_b200 attention runs on bf16/fp32 projections, got {t.dtype}')
    return t.contiguous()


# ----------------------------------------------------------------------------------------------------------
# nn.Linear through the tcgen05 GEMM.  The attention modules keep their nn.Linear
# children (checkpoint keys, LoRA wrappers, sp_attn_forward all rely on them); `_lin(mod, x)` computes
# mod(x) with uvb_linear_bf16 when `mod` is a plain nn.Linear running the product's inference configuration
# (CUDA, bf16 autocast or bf16 parameters, no autograd) and calls the module otherwise -- a wrapped or
# trainable projection keeps the semantics its owner gave it.  UVB_LINEAR=0 routes everything through the
# modules (A/B against cuBLAS).
# ----------------------------------------------------------------------------------------------------------
_USE_GEMM = os.environ.get('UVB_LINEAR', '1') != '0'
_LINEAR_ATTR = '_uvb_linear_operands'


def _version_of(t):
    try:
        return t._version
    except RuntimeError:           # inference tensors do not track a version: never cache-hit on them
        return object()


def _linear_operands(mod):
    """(weight bf16 [N, K], bias fp32 [N] holding bf16-rounded values | None) of an nn.Linear, cached ON THE MODULE
    until the parameters change: autocast casts weight AND bias to bf16 on every call (the reference pays that cast
    per autocast region); here the copy is made once per parameter version.  A bf16 parameter is used as it is (no
    copy).  Invalidation: parameter identity, `_version` (optimizer steps, load_state_dict, copy_), storage address
    and device (`.to()`, `.cpu()`); the drop-in modules also drop the copies in `_apply` (so `model.cpu()` releases
    them at once) and when the module is garbage-collected.  NOT detected: in-place edits through `.data`
    (`w.data += delta`, PEFT merge_and_unload) of an fp32 parameter -- call clear_linear_cache(model) after those."""
    w, b = mod.weight, mod.bias
    ver = (id(w), _version_of(w), w.data_ptr(), w.device, None if b is None else (id(b), _version_of(b), b.data_ptr()))
    hit = mod.__dict__.get(_LINEAR_ATTR)
    if hit is not None and hit[0] == ver:
        return hit[1], hit[2]
    with torch.no_grad():
        wb = w.detach() if w.dtype == torch.bfloat16 else w.detach().to(torch.bfloat16)
        bb = None if b is None else b.detach().to(torch.bfloat16).float()
    mod.__dict__[_LINEAR_ATTR] = (ver, wb, bb)
    return wb, bb


def clear_linear_cache(module=None):
    """Drop the cached bf16 GEMM operands of every nn.Linear under `module` (required after editing weights through
    `.data`, which no version counter sees).  Returns the number of entries dropped."""
    if module is None:
        raise TypeError('clear_linear_cache(module): pass the model (or sub-module) whose weights were edited')
    n = 0
    for m in module.modules():
        if m.__dict__.pop(_LINEAR_ATTR, None) is not None:
            n += 1
    return n


class _DropOperandsOnApply:
    """Mixin: `.to()` / `.cpu()` / `.cuda()` / `.half()` go through nn.Module._apply -- drop the cached bf16 operand
    copies first, so moving a model off the GPU (the reference's offload_model=True) frees them immediately."""

    def _apply(self, fn, *args, **kwargs):
        clear_linear_cache(self)
        return super()._apply(fn, *args, **kwargs)


def _gemm_ok(mod, x):
    if not (_USE_GEMM and type(mod) is nn.Linear and x.is_cuda and mod.weight.is_cuda):
        return False
    if torch.is_grad_enabled() and (x.requires_grad or mod.weight.requires_grad
                                    or (mod.bias is not None and mod.bias.requires_grad)):
        return False
    if mod.in_features % 8 != 0 or mod.out_features % 8 != 0 or x.dtype not in (torch.bfloat16, torch.float32):
        return False
    if torch.is_autocast_enabled('cuda'):
        return torch.get_autocast_dtype('cuda') == torch.bfloat16
    return mod.weight.dtype == torch.bfloat16 and x.dtype == torch.bfloat16


def _lin_to_peers(mod, x, peers):
    """mod(x) with the result scattered by column groups into the peers' buffers (uvb_linear_bf16_sp); False when the
    projection cannot take the GEMM route (the caller then projects and scatters separately)."""
    if not _gemm_ok(mod, x) or x.dim() != 3 or x.size(0) != 1 or (mod.out_features // peers[1]) % 64 != 0:
        return False
    w, b = _linear_operands(mod)
    _ext.linear(x if x.dtype == torch.bfloat16 else x.to(torch.bfloat16), w, b, peers=peers)
    return True


def _lin(mod, x, act=_ext.ACT_NONE):
    """mod(x) for a projection module (optionally followed by tanh-GELU when act is set: the caller has
    checked that the activation module is nn.GELU(approximate='tanh'))."""
    if _gemm_ok(mod, x):
        w, b = _linear_operands(mod)
        return _ext.linear(x if x.dtype == torch.bfloat16 else x.to(torch.bfloat16), w, b, act=act)
    y = mod(x)
    return y if act == _ext.ACT_NONE else nn.functional.gelu(y, approximate='tanh')


def _ffn_forward(ffn, h):
    """block.ffn(h): Linear -> GELU(tanh) -> Linear with the activation fused into the first GEMM's epilogue when
    the container has exactly the reference's structure (model.py:212-214)."""
    if (isinstance(ffn, nn.Sequential) and len(ffn) == 3 and type(ffn[1]) is nn.GELU and ffn[1].approximate == 'tanh'
            and _gemm_ok(ffn[0], h) and _gemm_ok(ffn[2], h)):
        return _lin(ffn[2], _lin(ffn[0], h, act=_ext.ACT_GELU_TANH))
    return ffn(h)


class WanSelfAttention(_DropOperandsOnApply, nn.Module):

    def __init__(self,
                 dim,
                 num_heads,
                 window_size=(-1, -1),
                 qk_norm=True,
                 eps=1e-6):
        assert dim % num_heads == 0
        super().__init__()
        self.dim = dim
        self.num_heads = num_heads
        self.head_dim = dim // num_heads
        self.window_size = window_size
        self.qk_norm = qk_norm
        self.eps = eps

        self.q = nn.Linear(dim, dim)
        self.k = nn.Linear(dim, dim)
        self.v = nn.Linear(dim, dim)
        self.o = nn.Linear(dim, dim)
        self.norm_q = WanRMSNorm(dim, eps=eps) if qk_norm else nn.Identity()
        self.norm_k = WanRMSNorm(dim, eps=eps) if qk_norm else nn.Identity()

    def _prologue(self, q_lin, k_lin, cos_sin, grid_sizes, tok_offset=0, groups=1, peers=None):
        """norm_q/norm_k (+RoPE) -> bf16 [B, L, N, 128] (or the Ulysses send layout when groups > 1; with
        `peers` the head groups are stored straight into the peer ranks' exchange buffers)."""
        wq, eps_q, pre_q = _norm_weight(self.norm_q)
        wk, eps_k, pre_k = _norm_weight(self.norm_k)
        if pre_q is not None:
            q_lin = pre_q(q_lin)
        if pre_k is not None:
            k_lin = pre_k(k_lin)
        q_lin = None if q_lin is None else _proj_for_kernel(q_lin)
        k_lin = None if k_lin is None else _proj_for_kernel(k_lin)
        eps = eps_q if wq is not None else eps_k
        return _ext.qk_norm_rope(q_lin, k_lin, wq, wk, eps, self.num_heads, cos_sin=cos_sin,
                                 grid_sizes=grid_sizes, tok_offset=tok_offset, groups=groups, peers=peers)

    def _out_proj(self, x):
        """o(x.flatten(2)); outside autocast the bf16 attention result is cast to the weight dtype
        (the reference hands fp32 to `o`, attention.py:130 -- the values are identical)."""
        x = x.flatten(2)
        if not torch.is_autocast_enabled():
            w = getattr(self.o, 'weight', None)
            if w is not None and w.dtype != x.dtype:
                x = x.to(w.dtype)
        return _lin(self.o, x)

    def forward(self, x, seq_lens, grid_sizes, freqs):
        r"""
        Args:
            x(Tensor): Shape [B, L, C]
            seq_lens(Tensor): Shape [B], valid tokens per sample (keys beyond are masked)
            grid_sizes(Tensor): Shape [B, 3], (F, H, W) token grid per sample
            freqs(Tensor): complex RoPE table [1024, C / num_heads / 2]
        """
        if tuple(self.window_size) != (-1, -1):
            raise NotImplementedError('univid_b200: sliding-window self-attention is not implemented')
        b, s, n, d = *x.shape[:2], self.num_heads, self.head_dim
        q, k = self._prologue(_lin(self.q, x), _lin(self.k, x), _cos_sin_table(freqs, x.device), grid_sizes)
        v = _lin(self.v, x).view(b, s, n, d)
        if v.dtype != torch.bfloat16:
            v = v.to(torch.bfloat16)
        x = _ext.fmha_fwd(q, k, v, k_lens=_k_lens_arg(seq_lens, b, s, x.device))
        return self._out_proj(x)


def _tensor_key(t):
    return (id(t), t.data_ptr(), _version_of(t), tuple(t.shape), t.dtype, t.device)


def _weights_key(*mods):
    out = []
    for m in mods:
        for prm in (getattr(m, 'weight', None), getattr(m, 'bias', None)):
            out.append(None if prm is None else (id(prm), prm.data_ptr(), _version_of(prm)))
    return tuple(out)


class WanCrossAttention(WanSelfAttention):
    # The context (text embedding of the prompt) is the SAME tensor in every one of the ~100 DiT calls of a sampling
    # run (two of them with classifier-free guidance), and so are the weights: its k / v projections -- and for the
    # fused text weighting the bias-free projections and the biases -- are computed once per (context tensor, weights)
    # and reused.  Keyed on tensor identity + version, so a context that is rebuilt per call
    # (e.g. pre-scaled by the reference's hook) simply misses.  At most `context_cache_size` contexts per module.
    context_cache_size = 2

    def _context_side(self, context, fused):
        """Cached context-side operands: ('plain': k normed [B, L2, N, 128] bf16, v [B, L2, N, 128] bf16) or
        ('fused': bias-free k projection bf16 [B, L2, C], bias-free v projection fp32 [B, L2, C], b_k, b_v fp32)."""
        cacheable = (self.context_cache_size > 0 and not torch.is_grad_enabled()
                     and type(self.k) is nn.Linear and type(self.v) is nn.Linear)
        key = None
        if cacheable:
            key = (fused, _tensor_key(context), _weights_key(self.k, self.v, self.norm_k),
                   torch.is_autocast_enabled('cuda'))
            cache = self.__dict__.setdefault('_uvb_ctx_cache', {})
            hit = cache.get(key)
            if hit is not None and hit[0]() is context:
                return hit[1]
        b, n, d = context.size(0), self.num_heads, self.head_dim
        if not fused:
            _, k = self._prologue(None, _lin(self.k, context), None, None)
            v = _lin(self.v, context).view(b, -1, n, d)
            if v.dtype != torch.bfloat16:
                v = v.to(torch.bfloat16)
            val = (k, v)
        else:
            # k(w*c) = w*(k(c) - b_k) + b_k and v(w*c) = w*(v(c) - b_v) + b_v: run the projections on the
            # unscaled context (as modules, so LoRA wrappers stay in the loop) and recover the biases as
            # the image of zero.
            zero = context.new_zeros(1, 1, context.size(-1))
            b_k, b_v = _lin(self.k, zero).flatten().float(), _lin(self.v, zero).flatten().float()
            k_lin = (_lin(self.k, context).float() - b_k).to(torch.bfloat16).contiguous()
            v_lin = _lin(self.v, context).float() - b_v
            val = (k_lin, v_lin, b_k, b_v)
        if cacheable:
            for old in [k for k, v in cache.items() if v[0]() is None]:       # contexts that no longer exist
                del cache[old]
            if len(cache) >= 2 * self.context_cache_size:                      # plain / fused entries of each context
                for old in list(cache.keys())[:len(cache) - 2 * self.context_cache_size + 1]:
                    del cache[old]
            cache[key] = (weakref.ref(context), val)
        return val

    def _apply(self, fn, *args, **kwargs):
        self.__dict__.pop('_uvb_ctx_cache', None)
        return super()._apply(fn, *args, **kwargs)

    def forward(self, x, context, context_lens, text_weight=1.0, text_len=0):
        r"""
        Args:
            x(Tensor): Shape [B, L1, C]
            context(Tensor): Shape [B, L2, C]
            context_lens(Tensor): Shape [B] or None
            text_weight, text_len: opt-in FUSED form of UniVid's dynamic text weighting: equivalent to
                calling with context[:, :text_len] * text_weight (what Wan22ContextWrapper's hook does,
                model_pipeline.py:1789-1797) but the scaled context is never materialised: the weight is
                folded into the k-norm prologue (exactly: k' = RMSNorm(w u + b_k)) and into the 512
                bias-free value rows, with the value bias added to the normalised output by the kernel.  With the defaults this is the reference forward (model.py:160-180); a context
                pre-scaled by the reference hook goes through that default path unchanged.
        """
        b, n, d = x.size(0), self.num_heads, self.head_dim
        fused = text_weight != 1.0 and text_len > 0
        q, _ = self._prologue(_lin(self.q, x), None, None, None)
        if not fused:
            k, v = self._context_side(context, False)
            lk = k.size(1)
            x = _ext.fmha_fwd(q, k, v, k_lens=_k_lens_arg(context_lens, b, lk, x.device))
            return self._out_proj(x)

        wk, eps_k, pre_k = _norm_weight(self.norm_k)
        if pre_k is not None:
            raise NotImplementedError('fused text weighting needs a WanRMSNorm or Identity norm_k')
        lk = context.size(1)
        k_lin, v_lin, b_k, b_v = self._context_side(context, True)
        w_vec = _weight_vector(lk, int(text_len), float(text_weight), x.device)
        # out = sum_j p_j (w_j l_j + b_v) = sum_j p_j (w_j l_j) + b_v: the weight rides on the 512 bias-free
        # value rows (one tiny elementwise op) instead of on every probability inside the attention kernel
        v_w = (v_lin * w_vec.view(1, lk, 1)).to(torch.bfloat16).view(b, lk, n, d)
        _, k = _ext.qk_norm_rope(None, k_lin, None, wk, eps_k, n, row_scale=w_vec, pre_bias=b_k)
        x = _ext.fmha_fwd(q, k, v_w, k_lens=_k_lens_arg(context_lens, b, lk, x.device), out_bias=b_v)
        return self._out_proj(x)


_WEIGHT_VECTORS = {}


def _weight_vector(lk, text_len, weight, device):
    """fp32 [lk]: `weight` on the first text_len rows, 1 elsewhere; cached per value (a sampling run revisits the same
    ~40 weights in every layer)."""
    key = (lk, text_len, weight, str(device))
    v = _WEIGHT_VECTORS.get(key)
    if v is None:
        v = torch.ones(lk, dtype=torch.float32, device=device)
        v[:text_len] = weight
        if len(_WEIGHT_VECTORS) > 256:
            _WEIGHT_VECTORS.clear()
        _WEIGHT_VECTORS[key] = v
    return v


class WanAttentionBlock(_DropOperandsOnApply, nn.Module):
    """DiT block: adaLN-modulated self-attention, cross-attention, FFN (model.py:183-259)."""

    def __init__(self,
                 dim,
                 ffn_dim,
                 num_heads,
                 window_size=(-1, -1),
                 qk_norm=True,
                 cross_attn_norm=False,
                 eps=1e-6):
        super().__init__()
        self.dim = dim
        self.ffn_dim = ffn_dim
        self.num_heads = num_heads
        self.window_size = window_size
        self.qk_norm = qk_norm
        self.cross_attn_norm = cross_attn_norm
        self.eps = eps

        self.norm1 = WanLayerNorm(dim, eps)
        self.self_attn = WanSelfAttention(dim, num_heads, window_size, qk_norm, eps)
        self.norm3 = WanLayerNorm(dim, eps, elementwise_affine=True) if cross_attn_norm else nn.Identity()
        self.cross_attn = WanCrossAttention(dim, num_heads, (-1, -1), qk_norm, eps)
        self.norm2 = WanLayerNorm(dim, eps)
        self.ffn = nn.Sequential(nn.Linear(dim, ffn_dim), nn.GELU(approximate='tanh'), nn.Linear(ffn_dim, dim))
        self.modulation = nn.Parameter(torch.randn(1, 6, dim) / dim**0.5)

    def forward(self, x, e, seq_lens, grid_sizes, freqs, context, context_lens, e_index=None):
        r"""
        x [B, L, C]; e [B, L or 1, 6, C] fp32 time modulation (a singleton token axis broadcasts, which is
        what a scalar timestep expanded over the sequence amounts to, model.py:460-468).
        e_index: int32 [B, L] or None.  When given, e is [1, U, 6, C] -- one row per DISTINCT timestep -- and token
        (b, l) uses row e_index[b, l] (WanModel.embed builds this form for per-token timesteps with few distinct
        values; same values as the reference's [B, L, 6, C] expansion, never materialised).
        """
        assert e.dtype == torch.float32
        with torch.amp.autocast('cuda', dtype=torch.float32):
            mod = self.modulation.unsqueeze(0) + e                       # fp32 [B, L or 1, 6, C] ([1, U, 6, C] indexed)
        if e_index is not None and not self._fused_glue_ok(x, mod):
            mod = mod[0][e_index.long()]                                 # eager path: materialise [B, L, 6, C]
            e_index = None
        shift_a, scale_a, gate_a, shift_f, scale_f, gate_f = (mod[:, :, k] for k in range(6))

        if self._fused_glue_ok(x, mod):
            return self._forward_fused(x, (shift_a, scale_a, gate_a, shift_f, scale_f, gate_f), seq_lens, grid_sizes,
                                       freqs, context, context_lens, e_index)
        return self._forward_eager(x, (shift_a, scale_a, gate_a, shift_f, scale_f, gate_f), seq_lens, grid_sizes, freqs,
                                   context, context_lens)

    def _mods(self, x, e, e_index):
        """The six modulation chunks of this block (model.py:239-240) and whether the fused glue path applies."""
        assert e.dtype == torch.float32
        with torch.amp.autocast('cuda', dtype=torch.float32):
            mod = self.modulation.unsqueeze(0) + e
        return tuple(mod[:, :, k] for k in range(6)), self._fused_glue_ok(x, mod)

    def _forward_eager(self, x, mods, seq_lens, grid_sizes, freqs, context, context_lens):
        shift_a, scale_a, gate_a, shift_f, scale_f, gate_f = mods

        y = self.self_attn(torch.addcmul(shift_a, self.norm1(x).float(), 1 + scale_a), seq_lens, grid_sizes, freqs)
        with torch.amp.autocast('cuda', dtype=torch.float32):
            x = x + y * gate_a
        x = x + self.cross_attn(self.norm3(x), context, context_lens)
        y = self.ffn(torch.addcmul(shift_f, self.norm2(x).float(), 1 + scale_f))
        with torch.amp.autocast('cuda', dtype=torch.float32):
            x = x + y * gate_f
        return x

    def _fused_glue_ok(self, x, mod):
        """The fused glue kernel covers the inference configuration of the product: fp32 residual stream on the
        GPU (or the bf16 output of the patch embedding entering the first block), bf16 autocast (so the branch
        outputs are bf16), no autograd, plain WanLayerNorm / Identity norms."""
        return (x.is_cuda and x.dtype in (torch.float32, torch.bfloat16) and x.dim() == 3 and x.size(-1) in _ext.GLUE_DIMS
                and torch.is_autocast_enabled('cuda') and torch.get_autocast_dtype('cuda') == torch.bfloat16
                and not (torch.is_grad_enabled() and (x.requires_grad or mod.requires_grad))
                and type(self.norm1) is WanLayerNorm and type(self.norm2) is WanLayerNorm
                and type(self.norm3) in (WanLayerNorm, nn.Identity)
                and not self.norm1.elementwise_affine and not self.norm2.elementwise_affine)

    def _forward_fused(self, x, mods, seq_lens, grid_sizes, freqs, context, context_lens, e_index=None, pending=None,
                       defer=False):
        """Same dataflow as above with the elementwise glue in uvb_block_glue: one pass per residual update,
        producing the next branch's bf16 input in the same pass.
        Cross-block fusion (used by WanModel._run_blocks, which owns x): with defer=True the block's LAST residual
        update `x + ffn(..) * gate_f` is not applied but returned as pending = (y, gate_f); the next block passes it
        in and its first glue call performs that update together with its own norm1 + modulation -- one launch and
        one full read + write of the fp32 residual stream less per block (12 of the 40 bytes per element the glue
        moves), same arithmetic in the same order."""
        shift_a, scale_a, gate_a, shift_f, scale_f, gate_f = mods
        ix = e_index
        if pending is not None:
            # x is the previous block's fp32 stream, owned by the caller of _run_blocks: updated in place
            y_prev, gate_prev = pending
            x, h = _ext.block_glue(x, y=y_prev, gate=gate_prev, scale=scale_a, shift=shift_a, eps=self.norm1.eps,
                                   inplace=True, index=ix)
        else:
            # a bf16 x (first block: the patch embedding ran under autocast) is widened to fp32 -- exact -- and norm1's
            # result is rounded to bf16 like WanLayerNorm's .type_as(x) does; the residual `x + y * gate` is fp32 either way
            first_bf16 = x.dtype == torch.bfloat16
            x = x.float().contiguous()
            _, h = _ext.block_glue(x, scale=scale_a, shift=shift_a, eps=self.norm1.eps, index=ix, ln_round_bf16=first_bf16)
        y = self.self_attn(h, seq_lens, grid_sizes, freqs)
        if isinstance(self.norm3, WanLayerNorm):
            ln3 = (self.norm3.weight, self.norm3.bias) if self.norm3.elementwise_affine else (None, None)
            x, h = _ext.block_glue(x, y=_bf16c(y), gate=gate_a, ln=ln3, eps=self.norm3.eps, index=ix)   # new x: the caller's is kept
        else:
            x, _ = _ext.block_glue(x, y=_bf16c(y), gate=gate_a, want_h=False, index=ix)
            h = x
        c = self.cross_attn(h, context, context_lens)
        x, h = _ext.block_glue(x, y=_bf16c(c), gate=None, scale=scale_f, shift=shift_f, eps=self.norm2.eps, inplace=True,
                               index=ix)
        y = _ffn_forward(self.ffn, h)
        if defer:
            return x, (_bf16c(y), gate_f)
        x, _ = _ext.block_glue(x, y=_bf16c(y), gate=gate_f, want_h=False, inplace=True, index=ix)
        return x


def _bf16c(t):
    return (t if t.dtype == torch.bfloat16 else t.to(torch.bfloat16)).contiguous()


class Head(nn.Module):
    """Final modulated LayerNorm + projection to patch pixels (model.py:262-290)."""

    def __init__(self, dim, out_dim, patch_size, eps=1e-6):
        super().__init__()
        self.dim = dim
        self.out_dim = out_dim
        self.patch_size = patch_size
        self.eps = eps
        self.norm = WanLayerNorm(dim, eps)
        self.head = nn.Linear(dim, math.prod(patch_size) * out_dim)
        self.modulation = nn.Parameter(torch.randn(1, 2, dim) / dim**0.5)

    def forward(self, x, e):
        assert e.dtype == torch.float32
        with torch.amp.autocast('cuda', dtype=torch.float32):
            shift, scale = (u.squeeze(2) for u in (self.modulation.unsqueeze(0) + e.unsqueeze(2)).chunk(2, dim=2))
            return self.head(self.norm(x) * (1 + scale) + shift)


class WanModel(nn.Module):
    """Wan DiT backbone (model.py:293-546) as a plain nn.Module harness around the attention hot path.
    Same constructor arguments, parameter names and forward signature as the reference; the reference's
    diffusers mixins (config registration, from_pretrained) are not reproduced."""

    def __init__(self,
                 model_type='t2v',
                 patch_size=(1, 2, 2),
                 text_len=512,
                 in_dim=16,
                 dim=2048,
                 ffn_dim=8192,
                 freq_dim=256,
                 text_dim=4096,
                 out_dim=16,
                 num_heads=16,
                 num_layers=32,
                 window_size=(-1, -1),
                 qk_norm=True,
                 cross_attn_norm=True,
                 eps=1e-6):
        super().__init__()
        assert model_type in ['t2v', 'i2v', 'ti2v', 's2v']
        self.model_type = model_type
        self.patch_size = patch_size
        self.text_len = text_len
        self.in_dim = in_dim
        self.dim = dim
        self.ffn_dim = ffn_dim
        self.freq_dim = freq_dim
        self.text_dim = text_dim
        self.out_dim = out_dim
        self.num_heads = num_heads
        self.num_layers = num_layers
        self.window_size = window_size
        self.qk_norm = qk_norm
        self.cross_attn_norm = cross_attn_norm
        self.eps = eps
        # per-token timesteps with at most this many distinct values are embedded once per value (0 = always expand)
        self.max_distinct_timesteps = 8

        self.patch_embedding = nn.Conv3d(in_dim, dim, kernel_size=patch_size, stride=patch_size)
        self.text_embedding = nn.Sequential(nn.Linear(text_dim, dim), nn.GELU(approximate='tanh'), nn.Linear(dim, dim))
        self.time_embedding = nn.Sequential(nn.Linear(freq_dim, dim), nn.SiLU(), nn.Linear(dim, dim))
        self.time_projection = nn.Sequential(nn.SiLU(), nn.Linear(dim, dim * 6))
        self.blocks = nn.ModuleList([
            WanAttentionBlock(dim, ffn_dim, num_heads, window_size, qk_norm, cross_attn_norm, eps)
            for _ in range(num_layers)
        ])
        self.head = Head(dim, out_dim, patch_size, eps)

        # plain attribute, not a buffer, so .to(dtype) leaves the complex128 table alone (model.py:397)
        assert (dim % num_heads) == 0 and (dim // num_heads) % 2 == 0
        d = dim // num_heads
        self.freqs = torch.cat([
            rope_params(1024, d - 4 * (d // 6)),
            rope_params(1024, 2 * (d // 6)),
            rope_params(1024, 2 * (d // 6)),
        ], dim=1)
        self.init_weights()

    def embed(self, x, t, context, seq_len, y=None):
        """Everything before the blocks (model.py:436-487): patchify + pad, time / text embeddings.
        Returns (x [B, seq_len, C], e, kwargs for the blocks)."""
        if self.model_type == 'i2v':
            assert y is not None
        device = self.patch_embedding.weight.device
        if self.freqs.device != device:
            self.freqs = self.freqs.to(device)
        if y is not None:
            x = [torch.cat([u, v], dim=0) for u, v in zip(x, y)]

        x = [self.patch_embedding(u.unsqueeze(0)) for u in x]
        grid_sizes = torch.stack([torch.tensor(u.shape[2:], dtype=torch.long) for u in x])
        x = [u.flatten(2).transpose(1, 2) for u in x]
        seq_lens = torch.tensor([u.size(1) for u in x], dtype=torch.long)
        assert seq_lens.max() <= seq_len
        x = torch.cat([torch.cat([u, u.new_zeros(1, seq_len - u.size(1), u.size(2))], dim=1) for u in x])

        # a [B] timestep is one value per sample: keep a singleton token axis and let it broadcast instead
        # of materialising seq_len identical rows (the reference expands, model.py:460-461; same values)
        if t.dim() == 1:
            t = t.unsqueeze(1)
        # Per-token timesteps [B, seq_len] (what the sampling loop passes, textimage2video.py:372-377) take only a
        # few distinct values (t2v: one; ti2v: 0 for the given frame and t for the rest): embed each distinct value
        # once and hand the blocks a [B, L] row index instead of a [B, L, 6, C] fp32 tensor (1.2 GB at 32 760
        # tokens, 9.3 GB at 75 600).  torch.unique costs one host sync per DiT forward.
        e_index = None
        if t.size(1) > 1 and self.max_distinct_timesteps > 0:
            uniq, inverse = torch.unique(t, return_inverse=True)
            if uniq.numel() <= self.max_distinct_timesteps:
                t, e_index = uniq.unsqueeze(0), inverse.to(torch.int32).contiguous()
        with torch.amp.autocast('cuda', dtype=torch.float32):
            bt, lt = t.shape
            e = self.time_embedding(sinusoidal_embedding_1d(self.freq_dim, t.flatten()).unflatten(0, (bt, lt)).float())
            e0 = self.time_projection(e).unflatten(2, (6, self.dim))
            assert e.dtype == torch.float32 and e0.dtype == torch.float32

        context = self._embed_context(context)
        kwargs = dict(e=e0, seq_lens=seq_lens, grid_sizes=grid_sizes, freqs=self.freqs, context=context,
                      context_lens=None)
        if e_index is not None:
            kwargs['e_index'] = e_index
        return x, e, kwargs

    def _embed_context(self, context):
        """text_embedding of the padded prompt embeddings (model.py:473-478).  A sampling run calls the DiT ~100 times
        with the same one or two prompt tensors: the embedded context is computed once per (input tensors, weights) and
        the SAME tensor object is handed to the blocks every time, which is what lets WanCrossAttention reuse its
        context-side projections.  Inference only; at most 4 prompts are remembered."""
        cacheable = not torch.is_grad_enabled() and all(torch.is_tensor(u) for u in context)
        key = None
        if cacheable:
            key = (tuple(_tensor_key(u) for u in context), _weights_key(self.text_embedding[0], self.text_embedding[2]),
                   torch.is_autocast_enabled('cuda'), torch.get_autocast_dtype('cuda'))
            cache = self.__dict__.setdefault('_uvb_text_cache', {})
            hit = cache.get(key)
            if hit is not None and all(r() is u for r, u in zip(hit[0], context)):
                return hit[1]
        out = self.text_embedding(
            torch.stack([torch.cat([u, u.new_zeros(self.text_len - u.size(0), u.size(1))]) for u in context]))
        if cacheable:
            if len(cache) >= 4:
                del cache[next(iter(cache))]
            cache[key] = ([weakref.ref(u) for u in context], out)
        return out

    def _apply(self, fn, *args, **kwargs):
        self.__dict__.pop('_uvb_text_cache', None)
        return super()._apply(fn, *args, **kwargs)

    def forward(self, x, t, context, seq_len, y=None):
        r"""
        x: list of [C_in, F, H, W] latents; t: [B] (or [B, seq_len]) timesteps; context: list of [L, C]
        text embeddings; seq_len: padded token count.  Returns a list of [C_out, F, H, W] tensors.
        """
        x, e, kwargs = self.embed(x, t, context, seq_len, y)
        x = self._run_blocks(x, kwargs)
        x = self.head(x, self.token_embedding(e, kwargs.get('e_index')))
        x = self.unpatchify(x, kwargs['grid_sizes'])
        return [u.float() for u in x]

    def _run_blocks(self, x, kwargs):
        """`for block in self.blocks: x = block(x, **kwargs)` (model.py:489-490).  When every block is a plain
        WanAttentionBlock on the fused-glue path, the last residual update of block i is deferred into the first glue
        call of block i + 1 (WanAttentionBlock._forward_fused: one launch and 30 % of the glue's HBM traffic less per
        block, identical arithmetic); anything else -- wrapped or patched blocks, training, other dtypes -- runs the
        reference loop."""
        blocks = list(self.blocks)
        plain = all(type(b) is WanAttentionBlock and 'forward' not in b.__dict__ for b in blocks)
        if not plain or not blocks:
            for block in blocks:
                x = block(x, **kwargs)
            return x
        e, e_index = kwargs['e'], kwargs.get('e_index')
        args = (kwargs['seq_lens'], kwargs['grid_sizes'], kwargs['freqs'], kwargs['context'], kwargs['context_lens'])
        pending = None
        for i, block in enumerate(blocks):
            mods, fused = block._mods(x, e, e_index)
            if not fused:
                if pending is not None:      # cannot happen for a homogeneous model; stay correct anyway
                    x, _ = _ext.block_glue(x, y=pending[0], gate=pending[1], want_h=False, inplace=True, index=e_index)
                    pending = None
                x = block(x, **kwargs)
                continue
            x, pending = block._forward_fused(x, mods, *args, e_index=e_index, pending=pending, defer=True)
        if pending is not None:
            x, _ = _ext.block_glue(x, y=pending[0], gate=pending[1], want_h=False, inplace=True, index=e_index)
        return x

    @staticmethod
    def token_embedding(e, e_index):
        """The head's per-token time embedding [B, L, C] from the de-duplicated form (e [1, U, C], e_index [B, L]);
        e is returned unchanged when there is no index."""
        return e if e_index is None else e[0][e_index.long()]

    def unpatchify(self, x, grid_sizes):
        """[L, C_out * prod(patch)] token rows back to [C_out, F*pf, H*ph, W*pw] (model.py:499-522)."""
        c = self.out_dim
        out = []
        for u, v in zip(x, grid_sizes.tolist()):
            u = u[:math.prod(v)].view(*v, *self.patch_size, c)
            u = torch.einsum('fhwpqrc->cfphqwr', u)
            out.append(u.reshape(c, *[i * j for i, j in zip(v, self.patch_size)]))
        return out

    def init_weights(self):
        """Xavier-uniform linears with zero bias; N(0, .02) text/time embeddings; zero output head
        (model.py:524-546)."""
        for m in self.modules():
            if isinstance(m, nn.Linear):
                nn.init.xavier_uniform_(m.weight)
                if m.bias is not None:
                    nn.init.zeros_(m.bias)
        nn.init.xavier_uniform_(self.patch_embedding.weight.flatten(1))
        for seq in (self.text_embedding, self.time_embedding):
            for m in seq.modules():
                if isinstance(m, nn.Linear):
                    nn.init.normal_(m.weight, std=.02)
        nn.init.zeros_(self.head.head.weight)
