"""Generate tests/golden/wan_animate_golden.pt: the UNMODIFIED reference WanAnimateCrossAttention and
WanAnimateSelfAttention (models/wan/utils/modules/animate/model_animate.py:54-146) executed on CPU in the build
container, bf16-autocast SDPA route and fp32 route, on seeded inputs.

    python tests/golden/make_animate_golden.py        (needs /root/reference)
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref_loader  # noqa: E402
from oracle import wan_attention_oracle as orc  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "wan_animate_golden.pt")
DIM, HEADS, EPS, IMG = 256, 2, 1e-6, 257


def animate_params(dim, generator):
    """State dict of a WanAnimateCrossAttention: the base q/k/v/o + norms plus k_img / v_img / norm_k_img."""
    prm = orc.init_attention_params(dim, generator, realistic_bias=True)
    bound = (6.0 / (2 * dim)) ** 0.5
    for name in ("k_img", "v_img"):
        prm[f"{name}.weight"] = (torch.rand(dim, dim, generator=generator) * 2 - 1) * bound
        prm[f"{name}.bias"] = 0.02 * torch.randn(dim, generator=generator)
    prm["norm_k_img.weight"] = 1 + 0.1 * torch.randn(dim, generator=generator)
    return prm


def animate_case(seed=11):
    """B=2, 70 queries, context = 257 image rows + 40 text rows, dim 256, 2 heads."""
    g = torch.Generator().manual_seed(seed)
    prm = animate_params(DIM, g)
    x = torch.randn(2, 70, DIM, generator=g)
    context = torch.randn(2, IMG + 40, DIM, generator=g)
    grid_sizes = torch.tensor([[2, 5, 7], [1, 7, 10]], dtype=torch.long)
    seq_lens = torch.tensor([70, 70], dtype=torch.long)
    return dict(prm=prm, x=x, context=context, grid_sizes=grid_sizes, seq_lens=seq_lens)


def checksums(case):
    cs = {k: float(case[k].double().abs().sum()) for k in ("x", "context")}
    for k, v in case["prm"].items():
        cs["prm." + k] = float(v.double().abs().sum())
    return cs


def main():
    assert ref_loader.available(), "reference tree not found"
    att, _ = ref_loader.load_modules()
    self_cls, cross_cls, ns = ref_loader.load_animate_attention()
    freqs = orc.make_freqs(128)
    case = animate_case()
    gold = {"checksums": checksums(case)}
    cross = cross_cls(DIM, HEADS, eps=EPS, use_img_emb=True)
    cross.load_state_dict({k: v.clone() for k, v in case["prm"].items()})
    cross = cross.float().eval()
    base = {k: v for k, v in case["prm"].items() if "_img" not in k}
    selfa = self_cls(DIM, HEADS, eps=EPS)
    selfa.load_state_dict({k: v.clone() for k, v in base.items()})
    selfa = selfa.float().eval()
    sdpa = ns["flash_attention"]
    with torch.no_grad():
        with torch.autocast("cpu", dtype=torch.bfloat16):
            gold["cross_bf16"] = cross(case["x"], case["context"], None).clone()
            gold["self_bf16"] = selfa(case["x"], case["seq_lens"], case["grid_sizes"], freqs).clone()
        ns["flash_attention"] = lambda q, k, v, **kw: att.attention(q, k, v, dtype=torch.float32)
        gold["cross_fp32"] = cross(case["x"], case["context"], None).clone()
        gold["self_fp32"] = selfa(case["x"], case["seq_lens"], case["grid_sizes"], freqs).clone()
        ns["flash_attention"] = sdpa
    torch.save(gold, OUT)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")
    for k, v in gold.items():
        if torch.is_tensor(v):
            print(f"  {k}: {tuple(v.shape)} {v.dtype}")


if __name__ == "__main__":
    main()
