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


def _sp_forward_worker(rank, world, port, results):
    """sp_dit_forward on a tiny WanModel whose blocks are replaced by a token-local stand-in (the real blocks need the
    CUDA kernels): checks the host logic of the sequence-parallel forward -- token chunking, the per-token timestep
    row index following the tokens, the head on the local chunk and the all-gather before unpatchify."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import importlib
        import types
        mdl = importlib.import_module("univid_b200.wan.modules.model")
        sp = importlib.import_module("univid_b200.wan.distributed.sequence_parallel")
        torch.manual_seed(0)                                          # identical weights on every rank
        model = mdl.WanModel(model_type="ti2v", dim=256, ffn_dim=256, num_heads=2, num_layers=2, text_len=8, text_dim=32,
                             freq_dim=32, in_dim=4, out_dim=4).eval()
        torch.nn.init.normal_(model.head.head.weight, std=0.05)

        class TokenLocalBlock(torch.nn.Module):
            """x + f(modulation row of the token): any per-token function commutes with the token sharding."""

            def forward(self, x, e, seq_lens, grid_sizes, freqs, context, context_lens, e_index=None):
                rows = e[0][e_index.long()] if e_index is not None else e.expand(x.size(0), x.size(1), -1, -1)
                return x * (1 + rows[:, :, 1]) + rows[:, :, 0] + context.float().mean()

        model.blocks = torch.nn.ModuleList([TokenLocalBlock() for _ in range(2)])
        g = torch.Generator().manual_seed(5)
        lat = [torch.randn(4, 2, 4, 8, generator=g)]                  # 2 x 2 x 4 = 16 tokens
        ctx = [torch.randn(5, 32, generator=g)]
        t = torch.full((1, 16), 700.0)
        t[0, :8] = 0.0                                                # first frame given: two distinct timesteps
        with torch.no_grad():
            full = model(lat, t, ctx, 16)[0]
            sharded = types.MethodType(sp.sp_dit_forward, model)(lat, t, ctx, 16)[0]
            model.max_distinct_timesteps = 0                          # the reference's materialised expansion
            sharded_expanded = types.MethodType(sp.sp_dit_forward, model)(lat, t, ctx, 16)[0]
        results[rank] = (full, sharded, sharded_expanded)
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_sp_dit_forward_host_logic_world2():
    world = 2
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_sp_forward_worker, args=(world, _free_port(), results), nprocs=world, join=True)
    for r in range(world):
        full, sharded, sharded_expanded = results[r]
        assert sharded.shape == full.shape == (4, 2, 4, 8)
        assert torch.allclose(sharded, full, atol=1e-5), (sharded - full).abs().max()
        assert torch.allclose(sharded_expanded, full, atol=1e-5)
    assert torch.equal(results[0][1], results[1][1])                 # every rank ends with the whole output
