"""Drop-in for the attention classes of models/wan/utils/modules/animate/model_animate.py (logical path
wan/modules/animate/model_animate.py): the Wan-Animate variants of the hot path.

  WanAnimateSelfAttention   [model_animate.py:54-83]   same computation as WanSelfAttention.forward: inherited
  WanAnimateCrossAttention  [model_animate.py:86-146]  cross-attention with a second key/value branch over the 257
                                                       CLIP image tokens at the head of the context

Only the attention classes live here; the rest of the Animate model (face / motion encoders, adapters, the block and
model classes) is outside the hot path and not reproduced.  Same kernels as wan/modules/model.py: q is normalised
once, the two attentions run back to back on the same q (257 and L2 keys, both ragged tiles of 128), their sum is
formed in fp32 exactly once like the reference's `x + img_x`, and `o` projects it.
"""
import torch
import torch.nn as nn

from .... import _ext
from ..attention import _k_lens_arg
from ..model import WanRMSNorm, WanSelfAttention, _lin

__all__ = ['WanAnimateSelfAttention', 'WanAnimateCrossAttention']

IMG_TOKENS = 257     # CLIP ViT-H/14 tokens prepended to the context (model_animate.py:118)


class WanAnimateSelfAttention(WanSelfAttention):
    """forward(x, seq_lens, grid_sizes, freqs): identical to WanSelfAttention.forward (model_animate.py:56-83
    repeats model.py:126-155 line by line), so the fused prologue + attention path is inherited."""


class WanAnimateCrossAttention(WanSelfAttention):

    def __init__(self, dim, num_heads, window_size=(-1, -1), qk_norm=True, eps=1e-6, use_img_emb=True):
        super().__init__(dim, num_heads, window_size, qk_norm, eps)
        self.use_img_emb = use_img_emb
        if use_img_emb:
            self.k_img = nn.Linear(dim, dim)
            self.v_img = nn.Linear(dim, dim)
            self.norm_k_img = WanRMSNorm(dim, eps=eps) if qk_norm else nn.Identity()

    def _kv(self, k_mod, v_mod, norm, ctx):
        """norm(k_mod(ctx)), v_mod(ctx) as bf16 [B, L, N, 128]."""
        b, n, d = ctx.size(0), self.num_heads, self.head_dim
        if isinstance(norm, WanRMSNorm):
            _, k = _ext.qk_norm_rope(None, _kernel_input(_lin(k_mod, ctx)), None, norm.weight, norm.eps, n)
        else:
            k = norm(_lin(k_mod, ctx)).view(b, -1, n, d).to(torch.bfloat16)
        v = _lin(v_mod, ctx).view(b, -1, n, d)
        return k, v if v.dtype == torch.bfloat16 else v.to(torch.bfloat16)

    def forward(self, x, context, context_lens):
        r"""
        Args:
            x(Tensor): Shape [B, L1, C]
            context(Tensor): Shape [B, 257 + L2, C] (image tokens first) when use_img_emb, else [B, L2, C]
            context_lens(Tensor): Shape [B] or None: valid TEXT keys per sample
        """
        b, n, d = x.size(0), self.num_heads, self.head_dim
        if self.use_img_emb:
            context_img, context = context[:, :IMG_TOKENS], context[:, IMG_TOKENS:]
        q, _ = self._prologue(_lin(self.q, x), None, None, None)
        k, v = self._kv(self.k, self.v, self.norm_k, context)
        out = _ext.fmha_fwd(q, k, v, k_lens=_k_lens_arg(context_lens, b, k.size(1), x.device))
        if self.use_img_emb:
            k_img, v_img = self._kv(self.k_img, self.v_img, self.norm_k_img, context_img)
            img = _ext.fmha_fwd(q, k_img, v_img)
            out = (out.float() + img.float()).to(torch.bfloat16)     # one fp32 sum, rounded where `o` would round it
        return self._out_proj(out)


def _kernel_input(t):
    if t.dtype not in (torch.bfloat16, torch.float32):
        raise NotImplementedError(f'univid_b200 attention runs on bf16/fp32 projections, got {t.dtype}')
    return t.contiguous()
