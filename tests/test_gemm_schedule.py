"""CPU model of the GEMM's static tile schedule (csrc/gemm_bf16_sm100.cuh::gemm_tile_coords and the host rules of
csrc/c_api.cu::launch_gemm / pick_tile_n / uvb_linear_bf16): every output tile is visited exactly once, the tiles a
wave of persistent workers holds share few row tiles (operand reuse in L2), the group width keeps the B slab under
the L2 budget, and the tile width / small-problem rules pick what DESIGN.md says they pick.  The arithmetic below
restates the C++ line by line; if one side changes, this test is where the other has to follow."""
import math

import pytest

BM = 128
SMS, PAIRS = 148, 74
L2_SLAB_BUDGET = 80e6


def tile_coords(t, n_m, n_n, group_n):
    per_group = n_m * group_n
    g = t // per_group
    r = t - g * per_group
    w = min(group_n, n_n - g * group_n)
    tm = r // w
    return tm, g * group_n + (r - tm * w)


def group_width(n_n, bn, K):
    fit = max(1, int(L2_SLAB_BUDGET / (bn * K * 2.0)))
    groups = (n_n + fit - 1) // fit
    return (n_n + groups - 1) // groups


def pick_tile_n(M, N, ctas, workers):
    n_m = (M + BM * ctas - 1) // (BM * ctas)

    def cost(bn, penalty):
        tiles = n_m * ((N + bn - 1) // bn)
        return math.ceil(tiles / workers) * bn * penalty
    return 192 if cost(192, 1.03) < cost(256, 1.0) else 256


def is_small(M, N):
    return ((M + BM - 1) // BM) * ((N + 63) // 64) <= SMS


SHAPES = [(32760, 1536, 1536), (32760, 8960, 1536), (32760, 1536, 8960), (75600, 5120, 5120), (75600, 13824, 5120),
          (75600, 5120, 13824), (9450, 5120, 5120), (27280, 3072, 3072), (27280, 14336, 3072), (1950, 1536, 1536),
          (512, 1536, 1536), (1, 1536, 1536), (1000, 2296, 200)]


@pytest.mark.parametrize("M,N,K", SHAPES)
@pytest.mark.parametrize("ctas", [1, 2])
def test_every_tile_once_and_waves_are_local(M, N, K, ctas):
    workers = PAIRS if ctas == 2 else SMS
    bn = pick_tile_n(M, N, ctas, workers)
    n_m = (M + BM * ctas - 1) // (BM * ctas)
    n_n = (N + bn - 1) // bn
    gn = group_width(n_n, bn, K)
    assert 1 <= gn <= n_n and gn * bn * K * 2.0 <= max(L2_SLAB_BUDGET, bn * K * 2.0)
    seen = set()
    coords = [tile_coords(t, n_m, n_n, gn) for t in range(n_m * n_n)]
    for tm, tn in coords:
        assert 0 <= tm < n_m and 0 <= tn < n_n
        seen.add((tm, tn))
    assert len(seen) == n_m * n_n                                        # a bijection: each tile exactly once
    # tiles that run at the same time (one wave = `workers` consecutive indices) touch few distinct row tiles: their
    # A panels are shared through L2 and the group's B slab is the only other operand
    for start in range(0, len(coords), workers):
        wave = coords[start:start + workers]
        rows = {tm for tm, _ in wave}
        assert len(rows) <= math.ceil(len(wave) / gn) + 2
        assert len({tn // gn for _, tn in wave}) <= 2                     # at most a group boundary inside a wave


def test_tile_width_and_small_problem_rules():
    # N = 1536 on 74 pairs: 128 row tiles x 6 column tiles of 256 = 10.4 waves -> 11; x 8 of 192 = 13.8 -> 14 (cheaper)
    assert pick_tile_n(32760, 1536, 2, PAIRS) == 192
    assert pick_tile_n(32760, 8960, 2, PAIRS) == 256                     # 4480 tiles: 60.5 waves, nothing to gain
    assert pick_tile_n(75600, 5120, 2, PAIRS) == 256                     # 296 x 20 = exactly 80 waves
    assert pick_tile_n(9450, 5120, 2, PAIRS) == 256                      # 8-GPU Ulysses shard: exactly 10 waves
    assert is_small(512, 1536) and is_small(1, 1536) and is_small(26, 256)
    assert not is_small(1950, 1536) and not is_small(512, 5120)
    # rasterisation groups: 14B ffn[2] (K = 13 824) splits its 20 column tiles in 2 x 10 (2 passes over A), the 1.3B
    # shapes keep all of B in one slab
    assert group_width(20, 256, 13824) == 10
    assert group_width(54, 256, 5120) == 27
    assert group_width(35, 256, 1536) == 35 and group_width(8, 192, 8960) == 8
