"""Drop-in for models/wan/utils/modules/attention.py of the reference (logical path
wan/modules/attention.py): same two entry points, same signatures, backed by the sm_100a
tcgen05 flash-attention kernel instead of flash_attn / flash_attn_interface / torch SDPA.

Behavioural contract kept from the reference (attention.py:24-130):
  * q [B, Lq, N, C], k/v [B, Lk, N, C]; q, k, v are cast to `dtype` if they are not already half
    (attention.py:56-83); the result is returned in q's ORIGINAL dtype (attention.py:130);
  * keys at positions >= k_lens[b] are ignored (attention.py:72-80);
  * softmax_scale None means C ** -0.5; q_scale multiplies q (folded into the softmax scale here).
What the reference forwards to flash-attn but this path does not implement raises
NotImplementedError instead of being silently ignored: causal, dropout, sliding window, q_lens,
fp16, head_dim != 128, grouped-query head counts.
"""
import warnings

import torch

from ... import _ext

__all__ = [
    'flash_attention',
    'attention',
]

# The reference probes for flash_attn wheels; here the native kernel is always the backend.
FLASH_ATTN_3_AVAILABLE = False
FLASH_ATTN_2_AVAILABLE = False
UVB_NATIVE_AVAILABLE = True


def _k_lens_arg(k_lens, b, lk, device):
    """int32 device tensor, or None when no key is masked (avoids a host->device copy per call)."""
    if k_lens is None:
        return None
    if not torch.is_tensor(k_lens):
        k_lens = torch.tensor(list(k_lens), dtype=torch.int32)
    if k_lens.numel() != b:
        raise ValueError(f'k_lens must have {b} entries')
    if not k_lens.is_cuda:
        if bool((k_lens >= lk).all()):
            return None
        return k_lens.to(device=device, dtype=torch.int32, non_blocking=True)
    return k_lens.to(torch.int32)


def flash_attention(
    q,
    k,
    v,
    q_lens=None,
    k_lens=None,
    dropout_p=0.,
    softmax_scale=None,
    q_scale=None,
    causal=False,
    window_size=(-1, -1),
    deterministic=False,
    dtype=torch.bfloat16,
    version=None,
):
    """
    q:              [B, Lq, Nq, C1].
    k:              [B, Lk, Nk, C1].
    v:              [B, Lk, Nk, C2].
    k_lens:         [B] valid key count per sample (tensor on any device, or a sequence).
    softmax_scale:  float, scaling of QK^T before the softmax (default C1 ** -0.5).
    dtype:          compute dtype applied when q/k/v are not half already; bfloat16 only.
    `version` and `deterministic` are accepted for signature compatibility (the kernel is
    deterministic; there is one backend).
    """
    half_dtypes = (torch.float16, torch.bfloat16)
    assert dtype in half_dtypes
    assert q.device.type == 'cuda' and q.size(-1) <= 256
    if causal:
        raise NotImplementedError('univid_b200.flash_attention: causal masking is not implemented')
    if dropout_p:
        raise NotImplementedError('univid_b200.flash_attention: dropout is not implemented')
    if tuple(window_size) != (-1, -1):
        raise NotImplementedError('univid_b200.flash_attention: sliding-window attention is not implemented')
    if q_lens is not None:
        raise NotImplementedError('univid_b200.flash_attention: q_lens is not implemented (all query rows are computed)')
    if dtype != torch.bfloat16 or any(u.dtype == torch.float16 for u in (q, k, v)):
        raise NotImplementedError('univid_b200.flash_attention computes in bfloat16 only')

    b, lq, lk, out_dtype = q.size(0), q.size(1), k.size(1), q.dtype

    def half(x):
        return x if x.dtype == torch.bfloat16 else x.to(dtype)

    q, k, v = half(q), half(k), half(v)
    scale = q.size(-1) ** -0.5 if softmax_scale is None else softmax_scale
    if q_scale is not None:
        scale = scale * float(q_scale)
    x = _ext.fmha_fwd(q, k, v, k_lens=_k_lens_arg(k_lens, b, lk, q.device), softmax_scale=scale)
    return x.type(out_dtype)


def attention(
    q,
    k,
    v,
    q_lens=None,
    k_lens=None,
    dropout_p=0.,
    softmax_scale=None,
    q_scale=None,
    causal=False,
    window_size=(-1, -1),
    deterministic=False,
    dtype=torch.bfloat16,
    fa_version=None,
):
    """Dispatcher of the reference (attention.py:133-179).  There the torch-SDPA branch is taken
    when no flash_attn wheel is importable and silently drops k_lens, softmax_scale and q_scale;
    here the native kernel is always available, so this forwards to flash_attention()."""
    if fa_version is not None and fa_version not in (2, 3):
        warnings.warn(f'fa_version={fa_version} ignored: univid_b200 has a single sm_100a backend')
    return flash_attention(
        q=q,
        k=k,
        v=v,
        q_lens=q_lens,
        k_lens=k_lens,
        dropout_p=dropout_p,
        softmax_scale=softmax_scale,
        q_scale=q_scale,
        causal=causal,
        window_size=window_size,
        deterministic=deterministic,
        dtype=dtype,
        version=fa_version,
    )
