"""Drop-in for models/wan/distributed/ulysses.py (DeepSpeed-Ulysses attention, arXiv 2309.14509).

`distributed_attention` keeps the reference signature and semantics (ulysses.py:9-47): q/k/v arrive
sharded over tokens [B, L/p, N, D], an all-to-all turns them into head shards over the full sequence
[B, L, N/p, D], attention runs per head, and the inverse exchange restores token shards.  The
exchange is expressed on destination-major buffers (see `pack_heads` / `attend_exchanged`) so that
the fused caller `sp_attn_forward` can have its producer kernels write those buffers directly.
"""
import torch
import torch.distributed as dist

from ... import _ext
from ..modules.attention import _k_lens_arg
from .util import exchange, get_world_size


def pack_heads(x, world):
    """[B, s, N, D] bf16 -> send buffer [p, B, s, N/p, D] (slot j = head group j, for rank j)."""
    if x.dtype != torch.bfloat16:
        x = x.to(torch.bfloat16)
    return _ext.head_scatter(x.contiguous(), world)


def _gathered_sequence(recv):
    """recv [p, B, s, n, D] (slot i = tokens of rank i) -> [B, p*s, n, D]; a free view when B == 1."""
    p, b, s, n, d = recv.shape
    if b == 1:
        return recv.view(1, p * s, n, d)
    return recv.permute(1, 0, 2, 3, 4).reshape(b, p * s, n, d)


def attend_exchanged(q_send, k_send, v_send, seq_lens, group=None):
    """q/k/v send buffers [p, B, s, N/p, D] -> local attention output [B, s, N, D] (bf16)."""
    p, b, s, n, d = q_send.shape
    q = _gathered_sequence(exchange(q_send, group))
    k = _gathered_sequence(exchange(k_send, group))
    v = _gathered_sequence(exchange(v_send, group))
    o = _ext.fmha_fwd(q, k, v, k_lens=_k_lens_arg(seq_lens, b, p * s, q.device))     # [B, L, n, D]
    # destination-major view of o: slot j = token chunk j
    o_send = o.view(b, p, s, n, d).transpose(0, 1)
    o_send = o_send.reshape(p, b, s, n, d) if b == 1 else o_send.contiguous()
    o_recv = exchange(o_send, group)                                              # slot i = head group i
    return o_recv.permute(1, 2, 0, 3, 4).reshape(b, s, p * n, d)


def distributed_attention(
        q,
        k,
        v,
        seq_lens,
        window_size=(-1, -1),
):
    """
    Args:
        q:           [B, Lq // p, Nq, C1].
        k:           [B, Lk // p, Nk, C1].
        v:           [B, Lk // p, Nk, C2].
        seq_lens:    [B], length of each (unsharded) sequence in the batch
        window_size: only (-1, -1) (global attention) is implemented.
    Returns [B, Lq // p, Nq, C2] in q's dtype.
    """
    if not dist.is_initialized():
        raise ValueError("distributed group should be initialized.")
    if tuple(window_size) != (-1, -1):
        raise NotImplementedError('univid_b200: sliding-window attention is not implemented')
    world = get_world_size()
    out_dtype = q.dtype
    if world == 1:
        from ..modules.attention import flash_attention
        return flash_attention(q, k, v, k_lens=seq_lens)
    if q.size(2) % world != 0:
        raise ValueError(f'{q.size(2)} heads cannot be split over {world} ranks')
    from . import p2p
    b, s, n, d = q.shape
    ctx = p2p.context(b, s, n, q.device) if (d == 128 and k.shape == q.shape and v.shape == q.shape) else None
    if ctx is not None:
        ctx.next_epoch()
        for t, ptrs in ((q, ctx.q_peers), (k, ctx.k_peers), (v, ctx.v_peers)):
            t = t if t.dtype == torch.bfloat16 else t.to(torch.bfloat16)
            _ext.head_scatter(t.contiguous(), world, peers=(ptrs, ctx.send_sb, ctx.send_sl))
        x = ctx.attend(_k_lens_arg(seq_lens, b, world * s, q.device))
        return x.to(out_dtype, copy=True)      # o_recv is reused by the next exchange
    x = attend_exchanged(pack_heads(q, world), pack_heads(k, world), pack_heads(v, world), seq_lens)
    return x.type(out_dtype)
