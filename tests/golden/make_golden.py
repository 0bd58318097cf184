"""Generate tests/golden/*.pt from the UNMODIFIED reference executed on CPU in the build container.

    python tests/golden/make_golden.py          (needs /root/reference; not available on the GPU box)

The reference has no golden vectors of its own; these files freeze what its
source computes on seeded inputs so the oracle (oracle/wan_attention_oracle.py) and, through it, the
CUDA path are pinned to the reference rather than to our reading of it.  Inputs are regenerated from
seeds at test time; a checksum of every input is stored so RNG drift is detected, not silently
accepted.
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref_loader  # noqa: E402
from oracle import wan_attention_oracle as orc  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
DIM, HEADS, EPS = 256, 2, 1e-6


def checksum(t):
    return float(t.double().abs().sum())


def small_case(seed=0):
    """Shared by make_golden.py and the tests: B=2, L=26 (24 / 20 real tokens), dim 256, 2 heads."""
    g = torch.Generator().manual_seed(seed)
    prm_self = orc.init_attention_params(DIM, g, realistic_bias=True)
    prm_cross = orc.init_attention_params(DIM, g, realistic_bias=True)
    x = torch.randn(2, 26, DIM, generator=g)
    context = torch.randn(2, 32, DIM, generator=g)
    grid_sizes = torch.tensor([[2, 3, 4], [1, 4, 5]], dtype=torch.long)
    seq_lens = torch.tensor([24, 20], dtype=torch.long)
    return dict(prm_self=prm_self, prm_cross=prm_cross, x=x, context=context,
                grid_sizes=grid_sizes, seq_lens=seq_lens)


def input_checksums(case):
    cs = {"x": checksum(case["x"]), "context": checksum(case["context"])}
    for grp in ("prm_self", "prm_cross"):
        for k, v in case[grp].items():
            cs[f"{grp}.{k}"] = checksum(v)
    return cs


def build_ref_module(cls, prm, dtype=torch.float32):
    m = cls(DIM, HEADS, eps=EPS)
    m.load_state_dict({k: v.clone() for k, v in prm.items()})
    return m.to(dtype).eval()


class _Cfg:
    use_dynamic_text_weight = True
    total_sampling_steps = 50
    text_weight_transition_ratio = 0.4
    text_weight_max = 1.3
    text_weight_min = 1.0
    text_weight_schedule = "cosine"
    bagel_sequence_length = 128


def main():
    assert ref_loader.available(), "reference tree not found"
    att, model = ref_loader.load_modules()
    torch.manual_seed(0)
    gold = {}

    # ---- KATs: tables, norm, rope (deterministic fp64/fp32 arithmetic) ------------------------------
    d = orc.HEAD_DIM
    freqs = torch.cat([model.rope_params(1024, d - 4 * (d // 6)), model.rope_params(1024, 2 * (d // 6)),
                       model.rope_params(1024, 2 * (d // 6))], dim=1)
    gold["freqs_real"], gold["freqs_imag"] = freqs.real.clone(), freqs.imag.clone()
    norm = model.WanRMSNorm(8, eps=1e-6)
    gold["rmsnorm_1to8"] = norm(torch.arange(1, 9, dtype=torch.float32).view(1, 1, 8)).detach()
    xk = (torch.arange(128, dtype=torch.float32) / 128).view(1, 1, 1, 128).expand(1, 6, 1, 128).contiguous()
    gold["rope_kat_grid123"] = model.rope_apply(xk, torch.tensor([[1, 2, 3]]), freqs)

    # ---- module-level cases ------------------------------------------------------------------------
    case = small_case(0)
    gold["input_checksums"] = input_checksums(case)
    x, ctx, gs, sl = case["x"], case["context"], case["grid_sizes"], case["seq_lens"]

    sa = build_ref_module(model.WanSelfAttention, case["prm_self"])
    ca = build_ref_module(model.WanCrossAttention, case["prm_cross"])
    with torch.no_grad():
        # the reference's torch-SDPA route under bf16 autocast (textimage2video.py:330)
        with torch.autocast("cpu", dtype=torch.bfloat16):
            gold["self_bf16"] = sa(x, sl, gs, freqs).clone()
            gold["cross_bf16"] = ca(x, ctx, None).clone()
            q = sa.norm_q(sa.q(x)).view(2, 26, HEADS, d)
            gold["q_rope_bf16path"] = model.rope_apply(q, gs, freqs).clone()
            gold["q_norm_bf16path"] = q.clone()
    # fp32 gold: fp32 params, no autocast, fp32 SDPA
    orig_fa = model.flash_attention
    model.flash_attention = lambda q, k, v, **kw: att.attention(q, k, v, dtype=torch.float32)
    with torch.no_grad():
        gold["self_fp32"] = sa(x, sl, gs, freqs).clone()
        gold["cross_fp32"] = ca(x, ctx, None).clone()
    model.flash_attention = orig_fa

    # ---- text-weight hook (Wan22ContextWrapper) ----------------------------------------------------
    Wrapper = ref_loader.load_context_wrapper()
    import logging

    class Holder(torch.nn.Module):
        def __init__(self, m):
            super().__init__()
            self.blk = m

    class Pipe:
        def __init__(self, m):
            self.model = Holder(m)
            self.text_encoder = type("T", (), {"__call__": lambda self, t, d: None})()

    ca2 = build_ref_module(model.WanCrossAttention, case["prm_cross"])
    wr = Wrapper(Pipe(ca2), None, logging.getLogger("golden"), _Cfg())
    wr.use_bagel_context, wr.bagel_context = True, [ctx]
    sched = []
    for c in range(0, 25):
        wr.set_timestep(c)
        sched.append(wr.text_weight_multiplier)
    gold["schedule_cosine_0_24"] = torch.tensor(sched, dtype=torch.float64)
    for name in ("linear", "exponential"):
        cfg = _Cfg()
        cfg.text_weight_schedule = name
        wr.config = cfg
        gold[f"schedule_{name}_0_24"] = torch.tensor([wr._calculate_text_weight(c) for c in range(25)],
                                                     dtype=torch.float64)
    wr.config = _Cfg()
    wr.set_timestep(5)
    gold["hook_w5"] = float(wr.text_weight_multiplier)
    with torch.no_grad(), torch.autocast("cpu", dtype=torch.bfloat16):
        gold["cross_hook_w5_bf16"] = ca2(x, ctx, None).clone()
    model.flash_attention = lambda q, k, v, **kw: att.attention(q, k, v, dtype=torch.float32)
    with torch.no_grad():
        gold["cross_hook_w5_fp32"] = ca2(x, ctx, None).clone()
    model.flash_attention = orig_fa

    # ---- sequence-parallel RoPE (reference sequence_parallel.rope_apply with simulated ranks) ------
    util, uly, sp = ref_loader.load_distributed()
    g = torch.Generator().manual_seed(7)
    xs = torch.randn(2, 28, HEADS, d, generator=g)       # L padded to 28 (multiple of 4)
    gold["sp_rope_input_checksum"] = checksum(xs)
    for world in (2, 4):
        outs = []
        for r in range(world):
            sp.get_rank, sp.get_world_size = (lambda r=r: r), (lambda world=world: world)
            outs.append(sp.rope_apply(xs.chunk(world, dim=1)[r], gs, freqs))
        gold[f"sp_rope_world{world}"] = torch.cat(outs, dim=1)
    gold["sp_rope_full"] = model.rope_apply(xs, gs, freqs)

    torch.save(gold, os.path.join(OUT, "wan_attention_golden.pt"))
    print("wrote", os.path.join(OUT, "wan_attention_golden.pt"),
          os.path.getsize(os.path.join(OUT, "wan_attention_golden.pt")), "bytes")
    for k, v in gold.items():
        if torch.is_tensor(v):
            print(f"  {k}: {tuple(v.shape)} {v.dtype}")


if __name__ == "__main__":
    main()
