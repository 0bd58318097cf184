"""WanAttentionBlock on the GPU: the fused glue kernel (uvb_block_glue) against its fp32 torch expression, and
the whole block (fused glue + attention kernels) against the oracle pinned to the reference block.  -m gpu."""
import importlib
import os

import pytest
import torch

from oracle import wan_attention_oracle as orc
from tests.golden.make_block_golden import DIM, EPS, FFN, HEADS, block_case

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _cos(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a @ b) / (a.norm() * b.norm() + 1e-30))


@pytest.mark.parametrize("dim", [256, 512, 1024, 1536, 2048, 3072, 4096, 5120])
@pytest.mark.parametrize("per_tok", [False, True])
def test_block_glue_kernel(dim, per_tok):
    from univid_b200 import _ext
    g = torch.Generator().manual_seed(dim + per_tok)
    B, L = 2, 37
    x = torch.randn(B, L, dim, generator=g).cuda() * 3 + 0.5
    y = torch.randn(B, L, dim, generator=g).to(torch.bfloat16).cuda()
    mod = (0.3 * torch.randn(B, L if per_tok else 1, 6, dim, generator=g)).cuda()
    w, b = (1 + 0.1 * torch.randn(dim, generator=g)).cuda(), (0.1 * torch.randn(dim, generator=g)).cuda()
    shift, scale, gate = mod[:, :, 0], mod[:, :, 1], mod[:, :, 2]
    ln = lambda t, ww=None, bb=None: torch.nn.functional.layer_norm(t, (dim,), ww, bb, 1e-6)
    # LN + modulation only
    _, h = _ext.block_glue(x, scale=scale, shift=shift, eps=1e-6)
    want = torch.addcmul(shift, ln(x), 1 + scale)
    assert h.dtype == torch.bfloat16 and (h.float() - want).abs().max() <= 0.0079 * want.abs().max()
    assert (h == want.to(torch.bfloat16)).float().mean() > 0.99          # same rounding point, rare 1-ulp flips
    # gated residual + affine LN, out of place
    x0 = x.clone()
    x1, h = _ext.block_glue(x, y=y, gate=gate, ln=(w, b), eps=1e-6)
    want_x = x0 + y.float() * gate
    assert torch.equal(x, x0) and torch.equal(x1, want_x)
    assert (h == ln(want_x, w, b).to(torch.bfloat16)).float().mean() > 0.99
    # plain residual, modulated LN, in place
    x2, h = _ext.block_glue(x1, y=y, gate=None, scale=scale, shift=shift, eps=1e-6, inplace=True)
    want_x2 = want_x + y.float()
    assert x2.data_ptr() == x1.data_ptr() and torch.equal(x2, want_x2)
    assert (h == torch.addcmul(shift, ln(want_x2), 1 + scale).to(torch.bfloat16)).float().mean() > 0.99
    # residual only
    x3, none = _ext.block_glue(x2, y=y, gate=gate, want_h=False, inplace=True)
    assert none is None and torch.equal(x3, want_x2 + y.float() * gate)


@pytest.mark.parametrize("dim", [256, 1536, 3072, 5120])
def test_block_glue_indexed_modulation_rows(dim):
    """index [B, L] + one modulation row per distinct timestep == the materialised per-token modulation, bit for bit."""
    from univid_b200 import _ext
    g = torch.Generator().manual_seed(dim)
    B, L, U = 2, 53, 3
    x = torch.randn(B, L, dim, generator=g).cuda()
    y = torch.randn(B, L, dim, generator=g).to(torch.bfloat16).cuda()
    rows = (0.3 * torch.randn(1, U, 6, dim, generator=g)).cuda()
    idx = torch.randint(0, U, (B, L), generator=g).to(torch.int32).cuda()
    full = rows[0][idx.long()]                                   # [B, L, 6, dim]
    for k_gate, k_scale, k_shift in ((2, 1, 0), (5, 4, 3)):
        x1, h1 = _ext.block_glue(x, y=y, gate=rows[:, :, k_gate], scale=rows[:, :, k_scale], shift=rows[:, :, k_shift],
                                 eps=1e-6, index=idx)
        x2, h2 = _ext.block_glue(x, y=y, gate=full[:, :, k_gate], scale=full[:, :, k_scale], shift=full[:, :, k_shift], eps=1e-6)
        assert torch.equal(x1, x2) and torch.equal(h1, h2)
    with pytest.raises(RuntimeError):
        _ext.block_glue(x, scale=rows[:, :, 1], shift=rows[:, :, 0], index=idx.long())


def test_model_forward_with_per_token_timesteps_deduplicated():
    """WanModel.forward with t [B, seq_len] as the reference sampling loop passes it (textimage2video.py:372-377):
    the de-duplicated form (one embedded row per distinct timestep + a row index) against the reference's
    materialised [B, L, 6, C] expansion (max_distinct_timesteps = 0) and, for uniform t, against the [B] form."""
    mdl = importlib.import_module("univid_b200.wan.modules.model")
    torch.manual_seed(0)
    model = mdl.WanModel(model_type="ti2v", dim=256, ffn_dim=512, num_heads=2, num_layers=2, text_len=32, text_dim=64,
                         freq_dim=64, in_dim=16, out_dim=16).cuda().eval()
    torch.nn.init.normal_(model.head.head.weight, std=0.05)
    g = torch.Generator(device="cuda").manual_seed(3)
    lat = [torch.randn(16, 3, 8, 12, device="cuda", generator=g)]       # 72 tokens
    ctx = [torch.randn(20, 64, device="cuda", generator=g)]
    L = 80                                                                # 8 padding tokens
    t_tok = torch.full((1, L), 640.0, device="cuda")
    t_tok[0, :24] = 0.0                                                   # first latent frame given (ti2v): t = 0
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        _, _, kw = model.embed(lat, t_tok, ctx, L)
        assert kw["e"].shape == (1, 2, 6, 256) and kw["e_index"].shape == (1, L) and kw["e_index"].dtype == torch.int32
        got = model(lat, t_tok, ctx, L)[0]
        model.max_distinct_timesteps = 0
        _, _, kw0 = model.embed(lat, t_tok, ctx, L)
        assert kw0["e"].shape == (1, L, 6, 256) and "e_index" not in kw0
        want = model(lat, t_tok, ctx, L)[0]
        model.max_distinct_timesteps = 8
        uni = model(lat, torch.full((1, L), 640.0, device="cuda"), ctx, L)[0]
        uni_b = model(lat, torch.tensor([640.0], device="cuda"), ctx, L)[0]
    assert torch.isfinite(got).all() and got.shape == lat[0].shape
    assert (got - want).abs().max().item() <= 2e-3 * max(1.0, want.abs().max().item())
    assert (uni - uni_b).abs().max().item() <= 2e-3 * max(1.0, uni_b.abs().max().item())
    assert (got - uni).abs().max() > 1e-3                                  # the per-token values really matter


def test_first_block_with_bf16_input_takes_the_fused_path():
    """The first block of the DiT receives the bf16 output of the patch embedding (model.py:447 under autocast):
    WanLayerNorm then returns bf16 (.type_as, model.py:98).  The fused glue reproduces that rounding point
    (ln_round_bf16); checked against the eager expression of the same module and against the oracle."""
    mdl = importlib.import_module("univid_b200.wan.modules.model")
    from univid_b200 import _ext
    case = block_case(3, False)
    freqs = orc.make_freqs(128)
    xb = case["x"].to(torch.bfloat16)
    want32 = orc.attention_block(xb, case["e"], case["prm"], case["seq_lens"], case["grid_sizes"], freqs,
                                 case["context"], None, HEADS, eps=EPS, bf16=False)
    blk = mdl.WanAttentionBlock(DIM, FFN, HEADS, cross_attn_norm=True, eps=EPS)
    blk.load_state_dict(case["prm"])
    blk = blk.cuda().eval()
    args = (xb.cuda(), case["e"].cuda(), case["seq_lens"], case["grid_sizes"], freqs.cuda(), case["context"].cuda(), None)
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        assert blk._fused_glue_ok(args[0], args[1])
        got = blk(*args)
        blk._fused_glue_ok = lambda *a: False
        eager = blk(*args)
    assert got.dtype == torch.float32 and eager.dtype == torch.float32
    assert (got - eager).abs().max().item() <= 2e-2
    assert (got.cpu() - want32).abs().max().item() <= 2e-2 and _cos(got.cpu(), want32) >= 0.9999
    # the kernel-level rounding point: LN rounded to bf16, then modulated
    g = torch.Generator().manual_seed(1)
    x = torch.randn(2, 19, 1536, generator=g).to(torch.bfloat16).float().cuda()
    mod = (0.3 * torch.randn(2, 1, 2, 1536, generator=g)).cuda()
    _, h = _ext.block_glue(x, scale=mod[:, :, 1], shift=mod[:, :, 0], eps=1e-6, ln_round_bf16=True)
    ln = torch.nn.functional.layer_norm(x, (1536,), None, None, 1e-6).to(torch.bfloat16).float()
    want = torch.addcmul(mod[:, :, 0], ln, 1 + mod[:, :, 1]).to(torch.bfloat16)
    assert (h == want).float().mean() > 0.99 and (h.float() - want.float()).abs().max() <= 0.0079 * want.float().abs().max()


@pytest.mark.parametrize("per_tok", [False, True])
def test_block_matches_oracle(per_tok):
    """Fused block vs the oracle (bit-exactly pinned to the reference block by tests/test_block_oracle_golden.py):
    north-star tolerance against the fp32 evaluation, and the eager-glue path of the same module as a cross-check."""
    mdl = importlib.import_module("univid_b200.wan.modules.model")
    case = block_case(3, per_tok)
    freqs = orc.make_freqs(128)
    want32 = orc.attention_block(case["x"], case["e"], case["prm"], case["seq_lens"], case["grid_sizes"], freqs,
                                 case["context"], None, HEADS, eps=EPS, bf16=False)
    blk = mdl.WanAttentionBlock(DIM, FFN, HEADS, cross_attn_norm=True, eps=EPS)
    blk.load_state_dict(case["prm"])
    blk = blk.cuda().eval()
    args = (case["x"].cuda(), case["e"].cuda(), case["seq_lens"], case["grid_sizes"], freqs.cuda(), case["context"].cuda(), None)
    x_before = args[0].clone()
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        assert blk._fused_glue_ok(args[0], args[1])
        got = blk(*args)
        blk._fused_glue_ok = lambda *a: False          # eager glue, same attention kernels
        eager = blk(*args)
    assert torch.equal(args[0], x_before), "the block must not modify its input"
    assert got.dtype == torch.float32
    err = (got.cpu() - want32).abs().max().item()
    assert err <= 2e-2 and _cos(got.cpu(), want32) >= 0.9999, (err, _cos(got.cpu(), want32))
    assert (got - eager).abs().max().item() <= 2e-2
