"""Generate tests/golden/wan_model_golden.pt: the UNMODIFIED reference WanModel.forward (model.py:410-497: patchify +
pad, sinusoidal time embedding, time_projection, blocks, head, unpatchify) executed on CPU in the build container in
fp32 (fp32 SDPA route) on seeded inputs and seeded parameters.

    python tests/golden/make_model_golden.py        (needs /root/reference or oracle/_ref)

Two timestep forms: the scalar-per-sample [B] form and the per-token [B, seq_len] form the sampling loop passes
(textimage2video.py:372-377; two distinct values, the ti2v pattern).  Parameters are regenerated at test time from the
seed (`seeded_state`), in sorted key order, so reference and drop-in models get the same weights without storing them.
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "wan_model_golden.pt")
KW = dict(model_type="t2v", dim=256, ffn_dim=512, num_heads=2, num_layers=2, text_dim=64, freq_dim=32, in_dim=4,
          out_dim=4, text_len=24, cross_attn_norm=True, eps=1e-6)
SEQ_LEN = 72          # 3*4*6 tokens per sample, NO padding: the reference's CPU (torch-SDPA) route ignores k_lens
                      # (attention.py:165-168) while its GPU (flash) route masks padded keys (:72-80); without padding
                      # the two coincide, so this golden pins the semantics the product runs on a GPU
PADDED_SEQ_LEN = 80   # padded variant (second sample 32 real tokens): checked against oracle.dit_forward(route="flash")


def seeded_state(model, seed=11):
    """Deterministic parameters for every key of model.state_dict() (sorted order): weights ~ N(0, 1/fan_in),
    biases ~ N(0, 0.02), norm weights ~ 1 + N(0, 0.1), modulation ~ N(0, 1/dim) -- the head is NOT zero like
    init_weights leaves it, so the output depends on everything upstream."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    for k in sorted(model.state_dict().keys()):
        v = model.state_dict()[k]
        if "norm" in k and k.endswith("weight"):
            t = 1 + 0.1 * torch.randn(v.shape, generator=g)
        elif k.endswith("bias"):
            t = 0.02 * torch.randn(v.shape, generator=g)
        elif k.endswith("modulation"):
            t = torch.randn(v.shape, generator=g) / v.shape[-1] ** 0.5
        else:
            fan_in = v[0].numel()
            t = torch.randn(v.shape, generator=g) / fan_in ** 0.5
        out[k] = t.to(v.dtype)
    return out


def model_case(seed=5, padded=False):
    g = torch.Generator().manual_seed(seed)
    second = (4, 2, 8, 8) if padded else (4, 3, 8, 12)
    lat = [torch.randn(4, 3, 8, 12, generator=g), torch.randn(*second, generator=g)]
    ctx = [torch.randn(20, 64, generator=g), torch.randn(9, 64, generator=g)]
    seq_len = PADDED_SEQ_LEN if padded else SEQ_LEN
    # scalar-per-sample form: the reference's `t.expand(t.size(0), seq_len)` (model.py:461) only accepts B == 1,
    # so that case runs the first sample alone
    t_scalar = torch.tensor([500.0])
    t_token = torch.full((2, seq_len), 700.0)
    t_token[0, :24] = 0.0          # first latent frame given (its tokens carry timestep 0)
    t_token[1, :16] = 0.0
    return dict(lat=lat, ctx=ctx, t_scalar=t_scalar, t_token=t_token, seq_len=seq_len)


def checksums(case):
    cs = {}
    for k in ("lat", "ctx"):
        for i, u in enumerate(case[k]):
            cs[f"{k}{i}"] = float(u.double().abs().sum())
    return cs


def inputs_for(case, form):
    """(latents, timesteps, contexts) of the `scalar` (B = 1) or `token` (B = 2) case."""
    if form == "scalar":
        return case["lat"][:1], case["t_scalar"], case["ctx"][:1]
    return case["lat"], case["t_token"], case["ctx"]


def run_reference_fp32(model_mod, att_mod, case, form):
    m = model_mod.WanModel(**KW)
    m.load_state_dict(seeded_state(m))
    m = m.float().eval()
    orig = model_mod.flash_attention
    model_mod.flash_attention = lambda q, k, v, **kw: att_mod.attention(q, k, v, dtype=torch.float32)
    try:
        with torch.no_grad():
            lat, t, ctx = inputs_for(case, form)
            return [u.clone() for u in m(lat, t, ctx, case["seq_len"])]
    finally:
        model_mod.flash_attention = orig


def main():
    import warnings
    from oracle import ref_loader
    assert ref_loader.available(), "reference tree not found"
    att, model = ref_loader.load_modules()
    case = model_case()
    gold = {"checksums": checksums(case), "kw": KW, "seq_len": SEQ_LEN}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        gold["scalar_fp32"] = run_reference_fp32(model, att, case, "scalar")
        gold["token_fp32"] = run_reference_fp32(model, att, case, "token")
    torch.save(gold, OUT)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")
    for k in ("scalar_fp32", "token_fp32"):
        print(" ", k, [tuple(u.shape) for u in gold[k]], [float(u.abs().max()) for u in gold[k]])


if __name__ == "__main__":
    main()
