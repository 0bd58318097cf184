"""Host-side model of the persistent attention kernel's static schedule (csrc/fmha_fwd_sm100.cuh, FmhaSched):
every (query block, key tile) must be covered exactly once, every split block has exactly one owner (the CTA
with its last key tile), the owner finds exactly the CTAs that wrote partials for it, each CTA writes at most
one partial, and a CTA never processes an owner piece before its own partial (no wait chains)."""
import itertools

import pytest


def schedule(U, G, n_kv, split=True):
    """Mirror of FmhaSched::init / seg for all CTAs: returns per-CTA list of (unit, a, b, owner)."""
    W, R = divmod(U, G)
    rem_total = R * n_kv if split else 0
    rem_lo = lambda c: rem_total * c // G
    out = []
    for g in range(G):
        segs = [(w * G + g, 0, n_kv, True) for w in range(W)]
        if not split:
            if g < R:
                segs.append((W * G + g, 0, n_kv, True))
        else:
            lo, hi = rem_lo(g), rem_lo(g + 1)
            if hi > lo:
                u0, a0 = divmod(lo, n_kv)
                ln = hi - lo
                if a0 + ln <= n_kv:
                    segs.append((W * G + u0, a0, a0 + ln, a0 + ln == n_kv))
                else:
                    segs.append((W * G + u0 + 1, 0, a0 + ln - n_kv, False))
                    segs.append((W * G + u0, a0, n_kv, True))
        out.append(segs)
    return out, rem_lo, W


def owner_parts(g, seg, rem_lo, W, G, n_kv):
    """Mirror of the owner epilogue's search for the CTAs holding key tiles [0, a) of its unit."""
    unit, a, b, owner = seg
    assert owner and a > 0
    unit_lo = (unit - W * G) * n_kv
    g_lo = g
    while g_lo > 0 and rem_lo(g_lo) > unit_lo:
        g_lo -= 1
    return [gp for gp in range(g_lo, g) if rem_lo(gp + 1) != rem_lo(gp)]


CASES = [(1536, 148, 256), (768, 148, 256), (384, 148, 256), (1480, 148, 591), (11840, 148, 591),
         (96, 148, 16), (1, 148, 1), (1, 1, 1), (1, 4, 4), (3, 12, 4), (5, 148, 1), (149, 148, 4),
         (295, 148, 4), (1536, 148, 4), (7, 3, 5), (2, 148, 591), (147, 148, 3), (200, 148, 2)]


@pytest.mark.parametrize("U,sms,n_kv", CASES)
@pytest.mark.parametrize("split", [True, False])
def test_schedule_covers_every_tile_once(U, sms, n_kv, split):
    # grid size as chosen by launch_fmha (c_api.cu)
    G = min(sms, U * n_kv) if split else min(sms, U)
    segs, rem_lo, W = schedule(U, G, n_kv, split)
    cover = {}
    for g, lst in enumerate(segs):
        partials = 0
        seen_owner_with_parts = False
        for (unit, a, b, owner) in lst:
            assert 0 <= unit < U and 0 <= a < b <= n_kv
            assert owner == (b == n_kv)
            for t in range(a, b):
                assert (unit, t) not in cover
                cover[(unit, t)] = g
            if not owner:
                partials += 1
                assert not seen_owner_with_parts, "partial after an owner piece that waits -> chain"
            elif a > 0:
                seen_owner_with_parts = True
        assert partials <= 1
    assert len(cover) == U * n_kv
    if not split:
        assert all(o for lst in segs for (_, _, _, o) in lst)
        return
    # owners find exactly the writers of their unit's partials
    writers = {}
    for g, lst in enumerate(segs):
        for (unit, a, b, owner) in lst:
            if not owner:
                writers.setdefault(unit, []).append(g)
    for g, lst in enumerate(segs):
        for seg in lst:
            unit, a, b, owner = seg
            if owner and a > 0:
                assert owner_parts(g, seg, rem_lo, W, G, n_kv) == sorted(writers.pop(unit))
                assert all(gp < g for gp in owner_parts(g, seg, rem_lo, W, G, n_kv))
    assert not writers, "a partial without an owner"
    # balance: no CTA does more than one key tile above the mean
    work = [sum(b - a for (_, a, b, _) in lst) for lst in segs]
    assert max(work) - min(work) <= 1 + (n_kv if False else 0)


def test_schedule_random_sweep():
    for U, G, n_kv in itertools.product([1, 2, 3, 5, 17, 147, 148, 149, 300], [1, 2, 7, 148], [1, 2, 3, 16]):
        test_schedule_covers_every_tile_once(U, G, n_kv, True)
        test_schedule_covers_every_tile_once(U, G, n_kv, False)
