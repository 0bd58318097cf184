"""Generate tests/golden/wan_block_golden.pt: the UNMODIFIED reference WanAttentionBlock (model.py:183-259)
executed on CPU in the build container, bf16-autocast SDPA route and fp32 route, on seeded inputs.

    python tests/golden/make_block_golden.py        (needs /root/reference)
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref_loader  # noqa: E402
from oracle import wan_attention_oracle as orc  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "wan_block_golden.pt")
DIM, FFN, HEADS, EPS = 256, 512, 2, 1e-6


def block_case(seed=3, per_token_modulation=False):
    """B=2, L=26 (24 / 20 real tokens), dim 256, 2 heads, ffn 512, 32 context rows."""
    g = torch.Generator().manual_seed(seed)
    prm = orc.init_block_params(DIM, FFN, g)
    x = torch.randn(2, 26, DIM, generator=g)
    context = torch.randn(2, 32, DIM, generator=g)
    e = 0.3 * torch.randn(2, 26 if per_token_modulation else 1, 6, DIM, generator=g)
    grid_sizes = torch.tensor([[2, 3, 4], [1, 4, 5]], dtype=torch.long)
    seq_lens = torch.tensor([26, 26], dtype=torch.long)      # no key masking: the SDPA route ignores k_lens
    return dict(prm=prm, x=x, context=context, e=e, grid_sizes=grid_sizes, seq_lens=seq_lens)


def checksums(case):
    cs = {k: float(case[k].double().abs().sum()) for k in ("x", "context", "e")}
    for k, v in case["prm"].items():
        cs["prm." + k] = float(v.double().abs().sum())
    return cs


def main():
    assert ref_loader.available(), "reference tree not found"
    att, model = ref_loader.load_modules()
    freqs = orc.make_freqs(128)
    gold = {}
    for tag, per_tok in (("bcast", False), ("pertoken", True)):
        case = block_case(3, per_tok)
        gold[f"{tag}_checksums"] = checksums(case)
        blk = model.WanAttentionBlock(DIM, FFN, HEADS, cross_attn_norm=True, eps=EPS)
        blk.load_state_dict({k: v.clone() for k, v in case["prm"].items()})
        blk = blk.float().eval()
        # the reference expands e over the sequence (model.py:460-468); a [B, 1, 6, C] e broadcasts identically
        e = case["e"].expand(2, 26, 6, DIM).contiguous()
        args = (case["x"], e, case["seq_lens"], case["grid_sizes"], freqs, case["context"], None)
        with torch.no_grad():
            with torch.autocast("cpu", dtype=torch.bfloat16):
                gold[f"{tag}_bf16"] = blk(*args).clone()
            orig = model.flash_attention
            model.flash_attention = lambda q, k, v, **kw: att.attention(q, k, v, dtype=torch.float32)
            gold[f"{tag}_fp32"] = blk(*args).clone()
            model.flash_attention = orig
    torch.save(gold, OUT)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")
    for k, v in gold.items():
        if torch.is_tensor(v):
            print(f"  {k}: {tuple(v.shape)} {v.dtype}")


if __name__ == "__main__":
    main()
