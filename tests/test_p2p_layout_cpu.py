"""Host-side layout arithmetic of the fused Ulysses exchange (wan/distributed/p2p.py::exchange_layout) on the CPU:
every rank "stores" its head groups into numpy models of the peers' exchange buffers with exactly the offsets and
strides the kernels are given, and the result must equal the oracle's emulation of the reference all_to_all
(util.py:21-31) in both directions."""
import numpy as np
import pytest
import torch

from oracle import wan_attention_oracle as orc
from univid_b200.wan.distributed import p2p


@pytest.mark.parametrize("B,s,N,p", [(1, 6, 4, 2), (2, 5, 8, 4), (1, 3, 8, 8), (3, 4, 6, 2)])
def test_exchange_layout_reproduces_all_to_all(B, s, N, p):
    n, L, D = N // p, p * s, 128
    g = torch.Generator().manual_seed(B * 100 + s * 10 + p)
    shards = [torch.randn(B, s, N, D, generator=g) for _ in range(p)]          # rank i's local q (token shard)
    lay = [p2p.exchange_layout(B, s, N, p, r) for r in range(p)]
    bufs = [np.zeros(lay[0]["nbytes"] // 2, dtype=np.float32) for _ in range(p)]   # one "bf16 element" per cell
    # --- scatter heads / gather sequence: rank i stores head group j into rank j's q_recv
    for i in range(p):
        li = lay[i]
        assert li["off_q"] % 256 == 0 and li["off_k"] % 256 == 0 and li["off_o"] % 256 == 0
        x = shards[i].numpy()
        for j in range(p):
            base = (li["off_q"] + li["slot_bytes"]) // 2
            for b in range(B):
                for l in range(s):
                    for h in range(n):
                        o = base + b * li["send_sb"] + l * li["send_sl"] + h * D
                        bufs[j][o:o + D] = x[b, l, j * n + h]
    want = orc.all_to_all_emulated(shards, scatter_dim=2, gather_dim=1)        # rank j: [B, L, n, D]
    for j in range(p):
        got = bufs[j][lay[j]["off_q"] // 2: lay[j]["off_q"] // 2 + B * L * n * D].reshape(B, L, n, D)
        assert np.array_equal(got, want[j].numpy()), f"q_recv of rank {j}"
    # --- back: rank j stores rows [i*s, (i+1)*s) of its [B, L, n, D] result into rank i's o_recv at head j*n
    for j in range(p):
        y = want[j].numpy() * 2 + 1
        for i in range(p):
            ob = lay[i]["off_o"] // 2
            for b in range(B):
                for l in range(s):
                    o = ob + ((b * s + l) * N + lay[j]["o_head_offset"]) * D
                    bufs[i][o:o + n * D] = y[b, i * s + l].reshape(-1)
    back = orc.all_to_all_emulated([w * 2 + 1 for w in want], scatter_dim=1, gather_dim=2)   # rank i: [B, s, N, D]
    for i in range(p):
        got = bufs[i][lay[i]["off_o"] // 2: lay[i]["off_o"] // 2 + B * s * N * D].reshape(B, s, N, D)
        assert np.array_equal(got, back[i].numpy()), f"o_recv of rank {i}"
        assert np.array_equal(got, shards[i].numpy() * 2 + 1)                   # and it is the identity round trip


def test_flag_words_do_not_collide():
    for p_ in (2, 4, 8):
        lays = [p2p.exchange_layout(1, 4, 8, p_, r) for r in range(p_)]
        words = [l["qkv_flag_bytes"] for l in lays] + [l["o_flag_bytes"] for l in lays]
        assert len(set(words)) == 2 * p_ and max(words) + 4 <= lays[0]["off_q"]
