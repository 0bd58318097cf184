"""Drop-in for models/wan/distributed/sequence_parallel.py: the functions the reference binds onto
WanSelfAttention / WanModel instances with types.MethodType when use_sp=True
(textimage2video.py:143-147).

sp_attn_forward is the multi-GPU hot path.  Default (NVLink peer memory, wan/distributed/p2p.py): the fused
prologue kernel normalises, rotates (with the rank's token offset) and stores every head group of q and k
straight into the exchange buffer of the rank that owns it, v is scattered the same way by a copy kernel,
the attention kernel reads [B, L, N/p, 128] in place and TMA-stores each output tile into the rank that
owns its tokens; flag kernels order the GPUs.  Without CUDA IPC (or with UVB_SP_P2P=0) the same producers
write local send buffers and the exchange is three NCCL all-to-alls in and one out.
"""
import torch

from ... import _ext
from ..modules.model import _cos_sin_table, _lin, _lin_to_peers, sinusoidal_embedding_1d  # noqa: F401
from . import p2p
from ..modules.attention import _k_lens_arg
from .ulysses import attend_exchanged, distributed_attention  # noqa: F401
from .util import gather_forward, get_rank, get_world_size


def pad_freqs(original_tensor, target_len):
    """Pad a [S, 1, C] rotation table with ones (identity rotation) up to target_len (sequence_parallel.py:10-20)."""
    seq_len, s1, s2 = original_tensor.shape
    pad = original_tensor.new_ones(target_len - seq_len, s1, s2)
    return torch.cat([original_tensor, pad], dim=0)


@torch.amp.autocast('cuda', enabled=False)
def rope_apply(x, grid_sizes, freqs):
    """Rank-local RoPE with the reference semantics (sequence_parallel.py:23-61): x [B, s, N, D] holds
    tokens [rank*s, (rank+1)*s) of the sequence; positions beyond f*h*w rotate by 1+0j.  Stand-alone,
    PyTorch ops in float64; sp_attn_forward uses the fused kernel instead."""
    s, half = x.size(1), x.size(3) // 2
    widths = [half - 2 * (half // 3), half // 3, half // 3]
    f_t, f_h, f_w = freqs.split(widths, dim=1)
    rank, world = get_rank(), get_world_size()
    out = x.to(torch.float64).clone()
    for i, (f, h, w) in enumerate(grid_sizes.tolist()):
        t = torch.arange(f * h * w, device=freqs.device)
        rot = torch.cat([f_t[t // (h * w)], f_h[(t // w) % h], f_w[t % w]], dim=1).unsqueeze(1)
        if rot.size(0) < s * world:
            rot = pad_freqs(rot, s * world)
        rot = rot[rank * s:(rank + 1) * s].to(x.device)
        c, sn = rot.real, rot.imag
        xe, xo = out[i, :s, :, 0::2].clone(), out[i, :s, :, 1::2].clone()
        out[i, :s, :, 0::2] = xe * c - xo * sn
        out[i, :s, :, 1::2] = xe * sn + xo * c
    return out.float()


def _plain_blocks(model, x, kwargs):
    for block in model.blocks:
        x = block(x, **kwargs)
    return x


def sp_dit_forward(
    self,
    x,
    t,
    context,
    seq_len,
    y=None,
):
    """WanModel.forward with the token dimension sharded over the ranks (sequence_parallel.py:64-144):
    embeddings are computed on every rank, each rank keeps its chunk of the tokens through the blocks
    and the head, and the outputs are all-gathered before unpatchify."""
    # Re-align the ranks on the host once per DiT forward (one tiny NCCL collective, ~20 us against a forward of tens
    # of milliseconds): whatever skew the ranks accumulated OUTSIDE the DiT (rank-0-only VAE decode or save, offload
    # reloads, lazy initialisation) is absorbed here by a collective with torch.distributed's own watchdog, so the
    # GPU-side flag waits of the fused exchange (uvb_sp_wait) only ever cover intra-forward skew.
    import torch.distributed as dist
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()
    x, e, kwargs = self.embed(x, t, context, seq_len, y)
    world, rank = get_world_size(), get_rank()
    x = torch.chunk(x, world, dim=1)[rank]
    if kwargs.get('e_index') is not None:   # de-duplicated per-token timesteps: shard the row index with the tokens
        kwargs['e_index'] = torch.chunk(kwargs['e_index'], world, dim=1)[rank].contiguous()
    elif e.size(1) > 1:   # per-token timesteps: shard the modulation with the tokens
        e = torch.chunk(e, world, dim=1)[rank]
        kwargs['e'] = torch.chunk(kwargs['e'], world, dim=1)[rank]
    x = self._run_blocks(x, kwargs) if hasattr(self, '_run_blocks') else _plain_blocks(self, x, kwargs)
    x = self.head(x, self.token_embedding(e, kwargs.get('e_index')))
    x = gather_forward(x, dim=1)
    x = self.unpatchify(x, kwargs['grid_sizes'])
    return [u.float() for u in x]


def sp_attn_forward(self, x, seq_lens, grid_sizes, freqs, dtype=torch.bfloat16):
    """WanSelfAttention.forward on a token shard x [B, L/p, C] (sequence_parallel.py:147-176)."""
    if dtype != torch.bfloat16:
        raise NotImplementedError('univid_b200 sequence-parallel attention computes in bfloat16 only')
    b, s, n, d = *x.shape[:2], self.num_heads, self.head_dim
    world, rank = get_world_size(), get_rank()
    if world == 1:
        return type(self).forward(self, x, seq_lens, grid_sizes, freqs)
    if n % world != 0:
        raise ValueError(f'{n} heads cannot be split over {world} ranks')
    ctx = p2p.context(b, s, n, x.device)
    if ctx is not None:
        ctx.next_epoch()
        # v: the projection's epilogue TMA-stores every head group into its rank's exchange buffer (B == 1); q and k
        # go through the fused norm + RoPE prologue, which stores the same way
        if not _lin_to_peers(self.v, x, (ctx.v_peers, world, ctx.send_sl)):
            v = _lin(self.v, x).view(b, s, n, d)
            v = v if v.dtype == torch.bfloat16 else v.to(torch.bfloat16)
            _ext.head_scatter(v.contiguous(), world, peers=(ctx.v_peers, ctx.send_sb, ctx.send_sl))
        self._prologue(_lin(self.q, x), _lin(self.k, x), _cos_sin_table(freqs, x.device), grid_sizes, tok_offset=rank * s,
                       groups=world, peers=(ctx.q_peers, ctx.k_peers, ctx.send_sb, ctx.send_sl))
        return self._out_proj(ctx.attend(_k_lens_arg(seq_lens, b, world * s, x.device)))
    v = _lin(self.v, x).view(b, s, n, d)
    if v.dtype != torch.bfloat16:
        v = v.to(torch.bfloat16)
    q_send, k_send = self._prologue(_lin(self.q, x), _lin(self.k, x), _cos_sin_table(freqs, x.device), grid_sizes,
                                    tok_offset=rank * s, groups=world)
    v_send = _ext.head_scatter(v.contiguous(), world)
    x = attend_exchanged(q_send, k_send, v_send, seq_lens)
    return self._out_proj(x)
