"""The oracle's restatement of WanAttentionBlock.forward (oracle.attention_block, model.py:219-259) against the
outputs of the unmodified reference block frozen in tests/golden/wan_block_golden.pt (make_block_golden.py)."""
import os

import pytest
import torch

from oracle import wan_attention_oracle as orc
from tests.golden.make_block_golden import DIM, EPS, FFN, HEADS, block_case, checksums  # noqa: F401

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def gold():
    return torch.load(os.path.join(ROOT, "tests", "golden", "wan_block_golden.pt"), map_location="cpu", weights_only=False)


@pytest.mark.parametrize("tag,per_tok", [("bcast", False), ("pertoken", True)])
def test_block_oracle_matches_reference(gold, tag, per_tok):
    case = block_case(3, per_tok)
    for k, v in gold[f"{tag}_checksums"].items():
        assert abs(checksums(case)[k] - v) <= 1e-9 * max(1.0, abs(v)), f"RNG drift in {k}"
    freqs = orc.make_freqs(128)
    args = (case["x"], case["e"], case["prm"], case["seq_lens"], case["grid_sizes"], freqs, case["context"], None, HEADS)
    got = orc.attention_block(*args, eps=EPS, bf16=True, route="sdpa")
    assert got.dtype == torch.float32
    assert torch.equal(got, gold[f"{tag}_bf16"]), (got - gold[f"{tag}_bf16"]).abs().max()
    got32 = orc.attention_block(*args, eps=EPS, bf16=False, route="sdpa")
    assert (got32 - gold[f"{tag}_fp32"]).abs().max() <= 5e-6


def test_layer_norm_kat():
    x = torch.arange(1, 9, dtype=torch.float32).view(1, 1, 8)
    want = (x - 4.5) / torch.sqrt(torch.tensor(5.25) + 1e-6)
    assert torch.allclose(orc.layer_norm(x, 1e-6), want, atol=1e-6)
    xb = x.to(torch.bfloat16)
    assert orc.layer_norm(xb, 1e-6).dtype == torch.bfloat16      # .type_as(x), model.py:98
