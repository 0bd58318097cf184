"""Drop-in for models/wan/distributed/util.py: process-group helpers and the Ulysses exchange.

The reference packs with p `.contiguous()` chunk copies, calls list-form dist.all_to_all and
concatenates (util.py:21-31).  Here the exchange is one `all_to_all_single` over a [p, ...] buffer;
on the hot path (sequence_parallel.sp_attn_forward) the producers already write that buffer in
destination-major order, so no pack kernels remain and this generic function is only the
API-compatible entry point.
"""
import torch
import torch.distributed as dist


def init_distributed_group():
    """Initialise the sequence-parallel group: NCCL, one process per GPU (util.py:6-10)."""
    if not dist.is_initialized():
        dist.init_process_group(backend='nccl')


def get_rank():
    return dist.get_rank()


def get_world_size():
    return dist.get_world_size()


def exchange(send, group=None):
    """send [p, ...] (slot j goes to rank j) -> recv [p, ...] (slot i came from rank i)."""
    recv = torch.empty_like(send)
    dist.all_to_all_single(recv, send, group=group)
    return recv


def all_to_all(x, scatter_dim, gather_dim, group=None, **kwargs):
    """`scatter` x along one dimension and `gather` along another (util.py:21-31): rank r ends up
    with chunk r (along scatter_dim) of every rank's x, concatenated in rank order along gather_dim."""
    world_size = dist.get_world_size(group) if group is not None else get_world_size()
    if world_size > 1:
        if x.size(scatter_dim) % world_size != 0:
            raise ValueError(f'dimension {scatter_dim} ({x.size(scatter_dim)}) is not divisible by {world_size} ranks')
        send = torch.stack(x.chunk(world_size, dim=scatter_dim))
        recv = exchange(send, group=group)
        x = torch.cat(recv.unbind(0), dim=gather_dim).contiguous()
    return x


def all_gather(tensor):
    world_size = dist.get_world_size()
    if world_size == 1:
        return [tensor]
    out = [torch.empty_like(tensor) for _ in range(world_size)]
    dist.all_gather(out, tensor)
    return out


def gather_forward(input, dim):
    """Concatenate every rank's tensor along `dim` (util.py:42-51)."""
    if dist.get_world_size() == 1:
        return input
    return torch.cat(all_gather(input), dim=dim).contiguous()
