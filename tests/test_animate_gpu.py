"""Wan-Animate attention classes on the GPU: the drop-in WanAnimateCrossAttention /
WanAnimateSelfAttention against the oracle (pinned to the reference classes by tests/test_animate_oracle_golden.py)
and the frozen reference outputs.  Tolerance (north_star): max-abs <= 2e-2, cosine >= 0.9999 vs the fp32 evaluation.
-m gpu."""
import importlib
import os

import pytest
import torch

from oracle import wan_attention_oracle as orc
from tests.golden.make_animate_golden import DIM, EPS, HEADS, IMG, animate_case, animate_params

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _cos(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a @ b) / (a.norm() * b.norm() + 1e-30))


def _check(got, want32):
    got = got.float().cpu()
    assert torch.isfinite(got).all()
    err = (got - want32).abs().max().item()
    assert err <= 2e-2 and _cos(got, want32) >= 0.9999, (err, _cos(got, want32))


def test_animate_cross_attention_matches_reference_golden():
    mod = importlib.import_module("univid_b200.wan.modules.animate.model_animate")
    gold = torch.load(os.path.join(ROOT, "tests", "golden", "wan_animate_golden.pt"), map_location="cpu", weights_only=False)
    case = animate_case()
    m = mod.WanAnimateCrossAttention(DIM, HEADS, eps=EPS)
    m.load_state_dict(case["prm"])
    m = m.cuda().eval()
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        got = m(case["x"].cuda(), case["context"].cuda(), None)
    assert got.dtype == torch.bfloat16
    _check(got, gold["cross_fp32"])
    assert (got.float().cpu() - gold["cross_bf16"].float()).abs().max() <= 2e-2      # the reference's own bf16 route
    sa = mod.WanAnimateSelfAttention(DIM, HEADS, eps=EPS)
    sa.load_state_dict({k: v for k, v in case["prm"].items() if "_img" not in k})
    sa = sa.cuda().eval()
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        got = sa(case["x"].cuda(), case["seq_lens"], case["grid_sizes"], orc.make_freqs(128).cuda())
    _check(got, gold["self_fp32"])


@pytest.mark.parametrize("dim,heads,lq,lt,lens", [
    (1536, 12, 1950, 512, None),            # 1.3B width, BASELINE configs[0] token count
    (3072, 24, 880, 512, [512, 300]),       # ti2v-5B width (24 heads), ragged text lengths (flash-route masking)
    (5120, 40, 300, 77, None),              # 14B width, short prompt
])
def test_animate_cross_attention_matches_oracle(dim, heads, lq, lt, lens):
    mod = importlib.import_module("univid_b200.wan.modules.animate.model_animate")
    g = torch.Generator().manual_seed(dim + lq)
    prm = animate_params(dim, g)
    b = 2 if lens else 1
    x = torch.randn(b, lq, dim, generator=g)
    context = torch.randn(b, IMG + lt, dim, generator=g)
    lens_t = None if lens is None else torch.tensor(lens)
    want = orc.animate_cross_attention(x, context, prm, heads, context_lens=lens_t, eps=EPS, bf16=False, route="varlen")
    m = mod.WanAnimateCrossAttention(dim, heads, eps=EPS)
    m.load_state_dict(prm)
    m = m.cuda().eval()
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        got = m(x.cuda(), context.cuda(), lens_t)
    _check(got, want)
    # without the image branch the class is plain cross-attention
    m2 = mod.WanAnimateCrossAttention(dim, heads, eps=EPS, use_img_emb=False)
    m2.load_state_dict({k: v for k, v in prm.items() if "_img" not in k})
    m2 = m2.cuda().eval()
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        got2 = m2(x.cuda(), context[:, IMG:].cuda(), lens_t)
    _check(got2, orc.cross_attention(x, context[:, IMG:], prm, heads, context_lens=lens_t, eps=EPS, bf16=False,
                                     route="varlen"))
