"""The drop-in modules on the GPU against (a) the frozen outputs of the reference's own source
(tests/golden) and (b) the oracle at BASELINE.json config #1 geometry.  -m gpu."""
import importlib

import pytest
import torch

from oracle import wan_attention_oracle as orc
from tests.golden.make_golden import DIM, EPS, HEADS

pytestmark = pytest.mark.gpu
MAX_ABS, MIN_COS = 2e-2, 0.9999
mdl = importlib.import_module("univid_b200.wan.modules.model")
att = importlib.import_module("univid_b200.wan.modules.attention")


def _cos(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a @ b) / (a.norm() * b.norm() + 1e-30))


def _check(got, gold_fp32, ref_bf16=None):
    got = got.float().cpu()
    assert torch.isfinite(got).all()
    err = (got - gold_fp32.float()).abs().max().item()
    assert err <= MAX_ABS and _cos(got, gold_fp32) >= MIN_COS, (err, _cos(got, gold_fp32))
    if ref_bf16 is not None:      # also as close to the reference's bf16 result as that is to fp32
        e2 = (got - ref_bf16.float()).abs().max().item()
        assert e2 <= MAX_ABS and _cos(got, ref_bf16) >= MIN_COS, (e2, _cos(got, ref_bf16))


def _module(cls, prm, dim=DIM, heads=HEADS):
    m = cls(dim, heads, eps=EPS)
    m.load_state_dict(prm)
    return m.cuda().eval()


def test_self_attention_matches_reference_golden(golden, small_case):
    c = small_case
    sa = _module(mdl.WanSelfAttention, c["prm_self"])
    freqs = orc.make_freqs(128).cuda()
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        got = sa(c["x"].cuda(), c["seq_lens"], c["grid_sizes"], freqs)
    assert got.dtype == torch.bfloat16 and got.shape == (2, 26, DIM)
    # the reference SDPA route ignores seq_lens (attention.py:165-168) while the kernel masks keys beyond
    # them like the flash route: compare the rows of sample 0 up to its 24 valid keys with a varlen oracle
    want = orc.self_attention(c["x"], c["prm_self"], c["seq_lens"], c["grid_sizes"], orc.make_freqs(128), HEADS, EPS,
                              bf16=False, route="varlen")
    _check(got, want)
    # without padding the two routes coincide: golden (reference) comparison on full-length samples
    full = torch.tensor([26, 26])
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        got = sa(c["x"].cuda(), full, c["grid_sizes"], freqs)
    _check(got, golden["self_fp32"], golden["self_bf16"])


def test_self_attention_fp32_modules_without_autocast(golden, small_case):
    c = small_case
    sa = _module(mdl.WanSelfAttention, c["prm_self"])
    with torch.no_grad():
        got = sa(c["x"].cuda(), torch.tensor([26, 26]), c["grid_sizes"], orc.make_freqs(128).cuda())
    assert got.dtype == torch.float32
    _check(got, golden["self_fp32"])


def test_cross_attention_matches_reference_golden(golden, small_case):
    c = small_case
    ca = _module(mdl.WanCrossAttention, c["prm_cross"])
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        got = ca(c["x"].cuda(), c["context"].cuda(), None)
    _check(got, golden["cross_fp32"], golden["cross_bf16"])


def test_text_weighted_cross_attention_hook_path_and_fused_path(golden, small_case):
    c = small_case
    ca = _module(mdl.WanCrossAttention, c["prm_cross"])
    w = golden["hook_w5"]
    x, ctx = c["x"].cuda(), c["context"].cuda()
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        hooked = ca(x, orc.weight_context(c["context"], w).cuda(), None)     # what the reference hook passes
        fused = ca(x, ctx, None, text_weight=w, text_len=16)
    _check(hooked, golden["cross_hook_w5_fp32"], golden["cross_hook_w5_bf16"])
    _check(fused, golden["cross_hook_w5_fp32"], golden["cross_hook_w5_bf16"])


def test_fused_schedule_equals_prescaled_context_over_the_flow(small_case):
    """cfg #5 in miniature: every weight of the 50-step schedule, fused path vs pre-scaled context."""
    from univid_b200 import tma
    c = small_case
    ca = _module(mdl.WanCrossAttention, c["prm_cross"])
    x, ctx = c["x"].cuda(), c["context"].cuda()
    cfg = tma.TextWeightConfig()
    for call in range(0, 22, 3):
        w = tma.calculate_text_weight(call, cfg)
        want = orc.cross_attention_text_weighted(c["x"], c["context"], c["prm_cross"], HEADS, w, eps=EPS, bf16=False)
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
            got = ca(x, ctx, None, text_weight=w, text_len=tma.text_len_for(ctx, cfg))
        _check(got, want)


@pytest.mark.parametrize("realistic", [False, True])
def test_config1_block_attention_stack(realistic):
    """BASELINE.json config #1: dim 1536, 12 heads, 5x30x52 latent -> (5,15,26) = 1950 tokens, 512 text tokens."""
    g = torch.Generator().manual_seed(0)
    dim, heads, L = 1536, 12, 1950
    prm_s = orc.init_attention_params(dim, g, realistic_bias=realistic)
    prm_c = orc.init_attention_params(dim, g, realistic_bias=realistic)
    x = torch.randn(1, L, dim, generator=g).to(torch.bfloat16).float()
    ctx = torch.randn(1, 512, dim, generator=g).to(torch.bfloat16).float()
    gs, sl = torch.tensor([[5, 15, 26]]), torch.tensor([L])
    freqs = orc.make_freqs(128)
    sa, ca = _module(mdl.WanSelfAttention, prm_s, dim, heads), _module(mdl.WanCrossAttention, prm_c, dim, heads)
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        got_s = sa(x.cuda(), sl, gs, freqs.cuda())
        got_c = ca(x.cuda(), ctx.cuda(), None)
        got_w = ca(x.cuda(), ctx.cuda(), None, text_weight=1.3, text_len=128)
    _check(got_s, orc.self_attention(x, prm_s, sl, gs, freqs, heads, EPS, bf16=False),
           orc.self_attention(x, prm_s, sl, gs, freqs, heads, EPS, bf16=True))
    _check(got_c, orc.cross_attention(x, ctx, prm_c, heads, None, EPS, bf16=False),
           orc.cross_attention(x, ctx, prm_c, heads, None, EPS, bf16=True))
    _check(got_w, orc.cross_attention_text_weighted(x, ctx, prm_c, heads, 1.3, eps=EPS, bf16=False))


def test_cfg_batched_forward_equals_two_single_forwards(small_case):
    """the cond / uncond passes of classifier-free guidance batched into one B = 2 call
    (different contexts, same latent) give exactly what two B = 1 calls give."""
    c = small_case
    freqs = orc.make_freqs(128).cuda()
    sa = _module(mdl.WanSelfAttention, c["prm_self"], DIM, HEADS)
    ca = _module(mdl.WanCrossAttention, c["prm_cross"], DIM, HEADS)
    x1 = c["x"][:1].cuda()
    x2 = torch.cat([x1, x1])
    ctx2 = torch.stack([c["context"][0], torch.zeros_like(c["context"][0])]).cuda()      # cond, "empty prompt"
    gs2, sl2 = c["grid_sizes"][:1].repeat(2, 1), c["seq_lens"][:1].repeat(2)
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        s_b = sa(x2, sl2, gs2, freqs)
        c_b = ca(x2, ctx2, None)
        for i in range(2):
            assert torch.equal(s_b[i:i + 1], sa(x1, sl2[:1], gs2[:1], freqs))
            assert torch.equal(c_b[i:i + 1], ca(x1, ctx2[i:i + 1], None))


def test_flash_attention_entry_point_dtype_contract():
    g = torch.Generator().manual_seed(4)
    q, k, v = (torch.randn(1, 200, 2, 128, generator=g) for _ in range(3))
    out = att.flash_attention(q.cuda(), k.cuda(), v.cuda().bfloat16(), k_lens=torch.tensor([150]))
    assert out.dtype == torch.float32                        # q's input dtype (attention.py:130)
    want = orc.attention_varlen(q, k, v, k_lens=torch.tensor([150]), compute_dtype=torch.bfloat16)
    assert (out.cpu() - want).abs().max() <= MAX_ABS
    out2 = att.attention(q.cuda().bfloat16(), k.cuda().bfloat16(), v.cuda().bfloat16(), q_scale=2.0, softmax_scale=0.04)
    want2 = orc.attention_varlen(q, k, v, softmax_scale=0.08, compute_dtype=torch.bfloat16)
    assert out2.dtype == torch.bfloat16 and (out2.float().cpu() - want2).abs().max() <= MAX_ABS
    for kw in (dict(causal=True), dict(dropout_p=0.1), dict(window_size=(8, 8)), dict(q_lens=torch.tensor([200]))):
        with pytest.raises(NotImplementedError):
            att.flash_attention(q.cuda(), k.cuda(), v.cuda(), **kw)


def test_training_mode_is_refused():
    sa = mdl.WanSelfAttention(256, 2).cuda()
    x = torch.randn(1, 8, 256, device="cuda", requires_grad=True)
    with pytest.raises(RuntimeError, match="forward-only"), torch.autocast("cuda", dtype=torch.bfloat16):
        sa(x, torch.tensor([8]), torch.tensor([[2, 2, 2]]), orc.make_freqs(128).cuda())


def test_tiny_dit_forward_is_finite_and_deterministic():
    """Two-block WanModel harness end to end: finite output of the right shape; deterministic.  (Parity of
    WanModel.forward with the reference: tests/test_reference_binding_gpu.py.)"""
    torch.manual_seed(0)
    m = mdl.WanModel(dim=256, ffn_dim=512, num_heads=2, num_layers=2, text_dim=64, freq_dim=32, in_dim=4, out_dim=4)
    torch.nn.init.normal_(m.head.head.weight, std=0.02)
    m = m.cuda().eval()
    lat = [torch.randn(4, 3, 8, 12, device="cuda")]
    ctx = [torch.randn(20, 64, device="cuda")]
    t = torch.tensor([500.0], device="cuda")
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        a = m(lat, t, ctx, seq_len=3 * 4 * 6 + 8)
        b = m(lat, t, ctx, seq_len=3 * 4 * 6 + 8)
    assert a[0].shape == (4, 3, 8, 12) and torch.isfinite(a[0]).all() and a[0].abs().max() > 0
    assert torch.equal(a[0], b[0])


@pytest.mark.parametrize("per_token", [False, True])
def test_cross_block_glue_fusion_is_bit_identical_to_the_block_loop(per_token):
    """WanModel._run_blocks defers the last residual update of block i into the first glue call of block i + 1: same
    arithmetic in the same order, so the stream after 3 blocks must equal the plain `for block in blocks` loop
    bit for bit (scalar and de-duplicated per-token timesteps); a patched block switches the fusion off."""
    torch.manual_seed(1)
    m = mdl.WanModel(dim=256, ffn_dim=512, num_heads=2, num_layers=3, text_dim=64, freq_dim=32, in_dim=4, out_dim=4)
    for blk in m.blocks:
        torch.nn.init.normal_(blk.modulation, std=0.3)
    m = m.cuda().eval()
    lat = [torch.randn(4, 3, 8, 12, device="cuda")]
    ctx = [torch.randn(20, 64, device="cuda")]
    seq_len = 3 * 4 * 6 + 8
    t = torch.full((1, seq_len), 700.0, device="cuda")
    if per_token:
        t[0, :24] = 0.0
    else:
        t = torch.tensor([700.0], device="cuda")
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        x, e, kwargs = m.embed(lat, t, ctx, seq_len)
        ref = x
        for blk in m.blocks:
            ref = blk(ref, **kwargs)
        n0 = _ext_launches()
        got = m._run_blocks(x, kwargs)
        fused_launches = _ext_launches() - n0
        n0 = _ext_launches()
        for blk in m.blocks:
            blk(x, **kwargs)
        loop_launches = _ext_launches() - n0
        assert torch.equal(got, ref)
        assert fused_launches == loop_launches - (len(m.blocks) - 1)       # one glue launch less per block boundary
        # a patched block (e.g. a hook wrapper bound onto the instance) keeps the reference loop
        import types
        m.blocks[1].forward = types.MethodType(type(m.blocks[1]).forward, m.blocks[1])
        assert torch.equal(m._run_blocks(x, kwargs), ref)


def _ext_launches():
    from univid_b200 import _ext
    return _ext.launch_count
