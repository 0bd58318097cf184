"""Checks that need the reference tree itself (build container only; skipped on the GPU box)."""
import logging

import pytest
import torch
import torch.nn as nn

from oracle import ref_loader
from oracle import wan_attention_oracle as orc

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="/root/reference not mounted")


def test_oracle_matches_reference_on_random_shapes():
    att, model = ref_loader.load_modules()
    g = torch.Generator().manual_seed(42)
    dim, heads = 384, 3
    prm = orc.init_attention_params(dim, g, realistic_bias=True)
    x = torch.randn(1, 61, dim, generator=g)
    gs, sl = torch.tensor([[3, 4, 5]]), torch.tensor([60])
    freqs = orc.make_freqs(128)
    ref = model.WanSelfAttention(dim, heads, eps=1e-6)
    ref.load_state_dict(prm)
    with torch.no_grad(), torch.autocast("cpu", dtype=torch.bfloat16):
        want = ref(x, sl, gs, freqs)
    got = orc.self_attention(x, prm, sl, gs, freqs, heads, 1e-6, bf16=True)
    assert torch.equal(got, want)


def test_parameter_names_match_reference_modules():
    _, model = ref_loader.load_modules()
    from univid_b200.wan.modules import model as mine
    kw = dict(dim=256, ffn_dim=512, num_heads=2, num_layers=2, text_dim=64, freq_dim=32)
    ref_keys = {k: tuple(v.shape) for k, v in model.WanModel(**kw).state_dict().items()}
    my_keys = {k: tuple(v.shape) for k, v in mine.WanModel(**kw).state_dict().items()}
    assert ref_keys == my_keys


def test_reference_context_wrapper_hooks_the_drop_in_modules():
    """Wan22ContextWrapper finds the drop-in WanCrossAttention by class name and pre-scales its context
    (model_pipeline.py:1742-1810); the drop-in forward is replaced by a recorder (no GPU here)."""
    from univid_b200.wan.modules import model as mine
    Wrapper = ref_loader.load_context_wrapper()
    seen = {}

    class Blk(nn.Module):
        def __init__(self):
            super().__init__()
            self.cross_attn = mine.WanCrossAttention(256, 2)

    class Pipe:
        def __init__(self):
            self.model = nn.ModuleList([Blk(), Blk()])
            self.text_encoder = type("T", (), {"__call__": lambda self, t, d: None})()

    pipe = Pipe()
    for i, blk in enumerate(pipe.model):
        blk.cross_attn.forward = (lambda x, context, context_lens, _i=i: seen.__setitem__(_i, context) or x)

    class Cfg:
        use_dynamic_text_weight = True
        total_sampling_steps = 50
        text_weight_transition_ratio = 0.4
        text_weight_max = 1.3
        text_weight_min = 1.0
        text_weight_schedule = "cosine"
        bagel_sequence_length = 128

    wr = Wrapper(pipe, None, logging.getLogger("t"), Cfg())
    assert len(wr.original_forward_methods) == 2
    ctx = torch.randn(1, 512, 256)
    wr.use_bagel_context, wr.bagel_context = True, [ctx]
    wr.set_timestep(0)
    x = torch.zeros(1, 4, 256)
    pipe.model[1].cross_attn(x, ctx, None)
    assert torch.equal(seen[1][:, :128], ctx[:, :128] * 1.3) and torch.equal(seen[1][:, 128:], ctx[:, 128:])
    assert torch.equal(seen[1], orc.weight_context(ctx, 1.3))
