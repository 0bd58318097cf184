"""The oracle's restatement of the Wan-Animate attention classes (oracle.animate_cross_attention,
models/wan/utils/modules/animate/model_animate.py:54-146) against the outputs of the
unmodified reference classes frozen in tests/golden/wan_animate_golden.pt (make_animate_golden.py), and live against
the reference source while it is mounted."""
import os

import pytest
import torch

from oracle import ref_loader
from oracle import wan_attention_oracle as orc
from tests.golden.make_animate_golden import DIM, EPS, HEADS, IMG, animate_case, checksums

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def gold():
    return torch.load(os.path.join(ROOT, "tests", "golden", "wan_animate_golden.pt"), map_location="cpu",
                      weights_only=False)


def test_animate_cross_attention_oracle_matches_reference(gold):
    case = animate_case()
    for k, v in gold["checksums"].items():
        assert abs(checksums(case)[k] - v) <= 1e-9 * max(1.0, abs(v)), f"RNG drift in {k}"
    got = orc.animate_cross_attention(case["x"], case["context"], case["prm"], HEADS, eps=EPS, bf16=True)
    assert got.dtype == torch.bfloat16 and torch.equal(got, gold["cross_bf16"])
    got32 = orc.animate_cross_attention(case["x"], case["context"], case["prm"], HEADS, eps=EPS, bf16=False)
    assert (got32 - gold["cross_fp32"]).abs().max() <= 5e-6


def test_animate_self_attention_is_the_base_self_attention(gold):
    """WanAnimateSelfAttention.forward repeats WanSelfAttention.forward: the base oracle reproduces its output."""
    case = animate_case()
    base = {k: v for k, v in case["prm"].items() if "_img" not in k}
    freqs = orc.make_freqs(128)
    got = orc.self_attention(case["x"], base, case["seq_lens"], case["grid_sizes"], freqs, HEADS, eps=EPS, bf16=True)
    assert torch.equal(got, gold["self_bf16"])
    got32 = orc.self_attention(case["x"], base, case["seq_lens"], case["grid_sizes"], freqs, HEADS, eps=EPS, bf16=False)
    assert (got32 - gold["self_fp32"]).abs().max() <= 5e-6


def test_image_branch_is_additive():
    """out = o(attn_text + attn_img): with the image value projection zeroed the image branch contributes exactly its
    value bias rows' attention, and without image embedding the class reduces to plain cross-attention."""
    case = animate_case()
    prm = case["prm"]
    no_img = orc.animate_cross_attention(case["x"], case["context"][:, IMG:], prm, HEADS, eps=EPS, bf16=False,
                                         use_img_emb=False)
    plain = orc.cross_attention(case["x"], case["context"][:, IMG:], prm, HEADS, eps=EPS, bf16=False)
    assert torch.equal(no_img, plain)
    zero = dict(prm)
    zero["v_img.weight"] = torch.zeros_like(prm["v_img.weight"])
    zero["v_img.bias"] = torch.zeros_like(prm["v_img.bias"])
    with_zero_img = orc.animate_cross_attention(case["x"], case["context"], zero, HEADS, eps=EPS, bf16=False)
    assert (with_zero_img - plain).abs().max() <= 1e-5


@pytest.mark.skipif(not ref_loader.available(), reason="reference tree not mounted")
def test_live_reference_class_matches_golden(gold):
    _, cross_cls, _ = ref_loader.load_animate_attention()
    case = animate_case()
    m = cross_cls(DIM, HEADS, eps=EPS, use_img_emb=True)
    m.load_state_dict({k: v.clone() for k, v in case["prm"].items()})
    m = m.float().eval()
    with torch.no_grad(), torch.autocast("cpu", dtype=torch.bfloat16):
        got = m(case["x"], case["context"], None)
    assert torch.equal(got, gold["cross_bf16"])


def test_dropin_class_has_the_reference_interface():
    import importlib
    import inspect
    mod = importlib.import_module("univid_b200.wan.modules.animate.model_animate")
    cross = mod.WanAnimateCrossAttention(DIM, HEADS, eps=EPS)
    assert type(cross).__name__ == "WanAnimateCrossAttention"
    assert list(inspect.signature(cross.forward).parameters) == ["x", "context", "context_lens"]
    assert list(inspect.signature(mod.WanAnimateCrossAttention.__init__).parameters) == [
        "self", "dim", "num_heads", "window_size", "qk_norm", "eps", "use_img_emb"]
    keys = set(cross.state_dict())
    assert keys == set(animate_case()["prm"]), keys ^ set(animate_case()["prm"])
    assert not hasattr(mod.WanAnimateCrossAttention(DIM, HEADS, use_img_emb=False), "k_img")
    assert issubclass(mod.WanAnimateSelfAttention, importlib.import_module("univid_b200.wan.modules.model").WanSelfAttention)
    if ref_loader.available():
        _, ref_cls, _ = ref_loader.load_animate_attention()
        assert set(ref_cls(DIM, HEADS, eps=EPS).state_dict()) == keys
