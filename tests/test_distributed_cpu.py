"""N>1 host logic on CPU: world_size-2 `gloo` processes run the drop-in util.all_to_all and are checked
against the oracle's single-process emulation of the reference exchange (util.py:21-31)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import wan_attention_oracle as orc


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, results):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from univid_b200.wan.distributed import util
        assert util.get_rank() == rank and util.get_world_size() == world
        g = torch.Generator().manual_seed(100 + rank)
        x = torch.randn(1, 6, 4, 8, generator=g)                      # [B, L/p, N, D]
        fwd = util.all_to_all(x, scatter_dim=2, gather_dim=1)         # -> [B, L, N/p, D]
        back = util.all_to_all(fwd, scatter_dim=1, gather_dim=2)      # -> [B, L/p, N, D]
        gathered = util.gather_forward(x, dim=1)
        results[rank] = (x, fwd, back, gathered)
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_all_to_all_matches_reference_semantics_world2():
    world = 2
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), results), nprocs=world, join=True)
    shards = [results[r][0] for r in range(world)]
    want = orc.all_to_all_emulated(shards, scatter_dim=2, gather_dim=1)
    for r in range(world):
        x, fwd, back, gathered = results[r]
        assert fwd.shape == (1, 12, 2, 8)
        assert torch.equal(fwd, want[r])
        assert torch.equal(back, x)                                    # the inverse exchange restores the shard
        assert torch.equal(gathered, torch.cat(shards, dim=1))


def test_distributed_attention_requires_initialised_group():
    from univid_b200.wan.distributed import ulysses
    q = torch.zeros(1, 4, 2, 128)
    with pytest.raises(ValueError, match="initialized"):             # ulysses.py:27-28
        ulysses.distributed_attention(q, q, q, torch.tensor([4]))


def test_sp_rope_standalone_matches_reference(golden, monkeypatch):
    from univid_b200.wan.distributed import sequence_parallel as sp
    g = torch.Generator().manual_seed(7)
    xs = torch.randn(2, 28, 2, 128, generator=g)
    gs = torch.tensor([[2, 3, 4], [1, 4, 5]])
    f = orc.make_freqs(128)
    for world in (2, 4):
        outs = []
        for r in range(world):
            monkeypatch.setattr(sp, "get_rank", lambda r=r: r)
            monkeypatch.setattr(sp, "get_world_size", lambda world=world: world)
            outs.append(sp.rope_apply(xs.chunk(world, dim=1)[r], gs, f))
        assert torch.equal(torch.cat(outs, dim=1), golden[f"sp_rope_world{world}"])
