"""Host-side logic of the drop-in (signatures, argument policing, schedule, hook plumbing) on CPU."""
import inspect

import pytest
import torch
import torch.nn as nn

from oracle import wan_attention_oracle as orc
from univid_b200 import tma
import importlib

# `wan.modules` re-exports a function called `attention`, which shadows the submodule attribute
att = importlib.import_module('univid_b200.wan.modules.attention')
mdl = importlib.import_module('univid_b200.wan.modules.model')


def test_flash_attention_signature_matches_reference():
    want = ["q", "k", "v", "q_lens", "k_lens", "dropout_p", "softmax_scale", "q_scale", "causal",
            "window_size", "deterministic", "dtype", "version"]
    assert list(inspect.signature(att.flash_attention).parameters) == want
    sig = inspect.signature(att.attention)
    assert list(sig.parameters) == want[:-1] + ["fa_version"]
    assert sig.parameters["dtype"].default == torch.bfloat16
    assert sig.parameters["window_size"].default == (-1, -1)


def test_flash_attention_refuses_cpu_tensors_like_the_reference():
    q = torch.zeros(1, 4, 1, 128, dtype=torch.bfloat16)
    with pytest.raises(AssertionError):      # attention.py:54: assert q.device.type == 'cuda'
        att.flash_attention(q, q, q)
    with pytest.raises(AssertionError):      # attention.py:53: dtype must be a half type
        att.flash_attention(q, q, q, dtype=torch.float32)


def test_kernel_wrappers_have_no_cpu_path():
    from univid_b200 import _ext
    q = torch.zeros(1, 4, 1, 128, dtype=torch.bfloat16)
    with pytest.raises(RuntimeError, match="no CPU path"):
        _ext.fmha_fwd(q, q, q)
    with pytest.raises(RuntimeError, match="no CPU path"):
        _ext.qk_norm_rope(torch.zeros(1, 4, 128), None, torch.ones(128), None, 1e-6, 1)


def test_k_lens_without_masking_needs_no_device_copy():
    assert att._k_lens_arg(None, 2, 10, "cpu") is None
    assert att._k_lens_arg(torch.tensor([10, 10]), 2, 10, "cpu") is None
    got = att._k_lens_arg(torch.tensor([10, 7]), 2, 10, "cpu")
    assert got.dtype == torch.int32 and got.tolist() == [10, 7]
    with pytest.raises(ValueError):
        att._k_lens_arg(torch.tensor([1, 2, 3]), 2, 10, "cpu")


def test_rope_params_and_standalone_rope_apply_match_reference(golden):
    d = 128
    f = torch.cat([mdl.rope_params(1024, d - 4 * (d // 6)), mdl.rope_params(1024, 2 * (d // 6)),
                   mdl.rope_params(1024, 2 * (d // 6))], dim=1)
    assert torch.equal(f.real, golden["freqs_real"]) and torch.equal(f.imag, golden["freqs_imag"])
    x = (torch.arange(128, dtype=torch.float32) / 128).view(1, 1, 1, 128).expand(1, 6, 1, 128).contiguous()
    assert torch.equal(mdl.rope_apply(x, torch.tensor([[1, 2, 3]]), f), golden["rope_kat_grid123"])


def test_standalone_rmsnorm_matches_reference(golden):
    n = mdl.WanRMSNorm(8, eps=1e-6)
    y = n(torch.arange(1, 9, dtype=torch.float32).view(1, 1, 8))
    assert torch.equal(y.detach(), golden["rmsnorm_1to8"])
    assert (n.dim, n.eps) == (8, 1e-6)


def test_attention_modules_keep_reference_names():
    sa = mdl.WanSelfAttention(256, 2)
    assert [n for n, _ in sa.named_children()] == ["q", "k", "v", "o", "norm_q", "norm_k"]
    assert sorted(sa.state_dict()) == sorted(
        ["q.weight", "q.bias", "k.weight", "k.bias", "v.weight", "v.bias", "o.weight", "o.bias",
         "norm_q.weight", "norm_k.weight"])
    assert (sa.dim, sa.num_heads, sa.head_dim, sa.window_size, sa.qk_norm, sa.eps) == (256, 2, 128, (-1, -1), True, 1e-6)
    ca = mdl.WanCrossAttention(256, 2)
    assert ca.__class__.__name__ == "WanCrossAttention" and isinstance(ca, mdl.WanSelfAttention)
    assert isinstance(mdl.WanSelfAttention(256, 2, qk_norm=False).norm_q, nn.Identity)
    assert list(inspect.signature(mdl.WanSelfAttention.forward).parameters) == ["self", "x", "seq_lens", "grid_sizes", "freqs"]
    assert list(inspect.signature(mdl.WanCrossAttention.forward).parameters)[:4] == ["self", "x", "context", "context_lens"]


def test_wan_model_state_dict_layout():
    m = mdl.WanModel(dim=256, ffn_dim=512, num_heads=2, num_layers=2, text_dim=64, freq_dim=32)
    keys = set(m.state_dict())
    for k in ("patch_embedding.weight", "text_embedding.0.weight", "text_embedding.2.bias",
              "time_embedding.0.weight", "time_embedding.2.weight", "time_projection.1.weight",
              "blocks.0.modulation", "blocks.1.self_attn.norm_q.weight", "blocks.0.cross_attn.k.weight",
              "blocks.0.norm3.weight", "blocks.1.ffn.0.weight", "blocks.1.ffn.2.bias", "head.head.weight",
              "head.modulation"):
        assert k in keys, k
    assert m.freqs.shape == (1024, 64) and m.freqs.dtype == torch.complex128
    assert torch.count_nonzero(m.head.head.weight) == 0          # init_weights (model.py:546)
    assert torch.count_nonzero(m.blocks[0].self_attn.q.bias) == 0


def test_text_weight_schedule_equals_oracle(golden):
    for name in ("cosine", "linear", "exponential"):
        cfg = tma.TextWeightConfig(text_weight_schedule=name)
        got = torch.tensor([tma.calculate_text_weight(c, cfg) for c in range(25)], dtype=torch.float64)
        assert torch.equal(got, golden[f"schedule_{name}_0_24"])
    cfg = tma.TextWeightConfig(total_sampling_steps=100)
    assert [tma.calculate_text_weight(c, cfg) for c in (0, 39, 40)] == [orc.text_weight(c, 100) for c in (0, 39, 40)]
    assert tma.calculate_text_weight(0, tma.TextWeightConfig(use_dynamic_text_weight=False)) == 1.0
    assert tma.text_len_for(torch.zeros(1, 512, 8), tma.TextWeightConfig()) == 128
    assert tma.text_len_for(torch.zeros(1, 32, 8), tma.TextWeightConfig()) == 16


class _Recorder(nn.Module):
    """Stands in for the CUDA forward: records what the hook passes down."""

    def __init__(self):
        super().__init__()
        self.calls = []

    def forward(self, x, context, context_lens, **kw):
        self.calls.append(kw)
        return x


def test_fused_schedule_hooks_cross_attention_and_counts_dit_calls():
    class WanCrossAttention(_Recorder):
        pass

    class Dit(nn.Module):
        def __init__(self):
            super().__init__()
            self.blocks = nn.ModuleList([WanCrossAttention(), WanCrossAttention()])

        def forward(self, x, ctx):
            for b in self.blocks:
                x = b(x, ctx, None)
            return x

    dit = Dit()
    x, ctx = torch.zeros(1, 4, 8), torch.zeros(1, 512, 8)
    with tma.FusedTextWeightSchedule(dit, tma.TextWeightConfig()) as sched:
        for _ in range(22):
            dit(x, ctx)
        assert sched.call_index == 22
    calls = dit.blocks[1].calls
    assert calls[0] == {"text_weight": 1.3, "text_len": 128}
    assert abs(calls[5]["text_weight"] - orc.text_weight(5)) < 1e-15
    assert calls[20] == {} and calls[21] == {}         # weight back to 1.0 -> plain path
    assert "forward" not in dit.__dict__ and "forward" not in dit.blocks[0].__dict__   # hooks removed
    dit(x, ctx)
    assert dit.blocks[0].calls[-1] == {}


def test_grad_mode_is_refused():
    from univid_b200 import _ext
    t = torch.zeros(1, 4, 1, 128, dtype=torch.bfloat16, requires_grad=True)
    with pytest.raises(RuntimeError, match="forward-only"):
        _ext._no_grad_only(t)
    with torch.no_grad():
        _ext._no_grad_only(t)


def test_lin_helper_routes_cpu_and_wrapped_modules_through_the_module():
    """_lin / _ffn_forward (SURVEY.md sec. 8f rank 2 host side): anything that is not a plain nn.Linear on CUDA in the
    inference configuration is computed by the module itself; the bf16 operand cache follows parameter updates."""
    import torch
    import torch.nn as nn
    mdl = importlib.import_module("univid_b200.wan.modules.model")
    lin = nn.Linear(16, 24)
    x = torch.randn(3, 16)
    assert not mdl._gemm_ok(lin, x)
    assert torch.equal(mdl._lin(lin, x), lin(x))
    assert torch.equal(mdl._lin(lin, x, act=1), torch.nn.functional.gelu(lin(x), approximate="tanh"))
    ffn = nn.Sequential(nn.Linear(16, 32), nn.GELU(approximate="tanh"), nn.Linear(32, 16))
    assert torch.equal(mdl._ffn_forward(ffn, x), ffn(x))
    w, b = mdl._linear_operands(lin)
    assert w.dtype == torch.bfloat16 and b.dtype == torch.float32
    assert torch.equal(w, lin.weight.detach().to(torch.bfloat16))
    assert torch.equal(b, lin.bias.detach().to(torch.bfloat16).float())     # autocast rounds the bias to bf16
    assert mdl._linear_operands(lin)[0] is w                                 # cached
    with torch.no_grad():
        lin.weight.mul_(3.0)
    w2, _ = mdl._linear_operands(lin)
    assert w2 is not w and torch.equal(w2, lin.weight.detach().to(torch.bfloat16))


def test_unipc_dropin_has_the_reference_interface():
    """Drop-in FlowUniPCMultistepScheduler: constructor / method signatures of
    models/wan/utils/fm_solvers_unipc.py:79-97, :162-169, :657-662; schedule identical to the oracle; options the
    fused kernel does not cover raise instead of being ignored."""
    import torch
    from oracle import ref_loader
    from oracle import unipc_oracle as uo
    mod = importlib.import_module("univid_b200.wan.utils.fm_solvers_unipc")
    cls = mod.FlowUniPCMultistepScheduler
    want_init = ["self", "num_train_timesteps", "solver_order", "prediction_type", "shift", "use_dynamic_shifting",
                 "thresholding", "dynamic_thresholding_ratio", "sample_max_value", "predict_x0", "solver_type",
                 "lower_order_final", "disable_corrector", "solver_p", "timestep_spacing", "steps_offset",
                 "final_sigmas_type"]
    assert list(inspect.signature(cls.__init__).parameters) == want_init
    assert list(inspect.signature(cls.set_timesteps).parameters) == ["self", "num_inference_steps", "device", "sigmas", "mu", "shift"]
    assert list(inspect.signature(cls.step).parameters) == ["self", "model_output", "timestep", "sample", "return_dict", "generator"]
    if ref_loader.available():
        ref = ref_loader.load_unipc_scheduler()
        assert list(inspect.signature(ref.__init__).parameters) == want_init
        assert list(inspect.signature(ref.step).parameters) == list(inspect.signature(cls.step).parameters)
    s = cls(num_train_timesteps=1000, shift=1, use_dynamic_shifting=False)
    s.set_timesteps(50, device="cpu", shift=5.0)
    t, sig = uo.sampling_schedule(50, 5.0)
    assert torch.equal(s.timesteps, t) and torch.equal(s.sigmas, sig) and s.config.solver_order == 2
    for kw in (dict(solver_order=3), dict(thresholding=True), dict(predict_x0=False), dict(use_dynamic_shifting=True),
               dict(prediction_type="epsilon"), dict(final_sigmas_type="sigma_min")):
        with pytest.raises(NotImplementedError):
            cls(**kw)
    with pytest.raises(RuntimeError):                     # no CPU path
        s.step(torch.zeros(1, 4), s.timesteps[0], torch.zeros(1, 4))
    # the scalar coefficients are the oracle's, bit for bit (both mirror fm_solvers_unipc.py:395-455 / :549-606)
    for corr, (i_t, i_s0, hist, order) in ((True, (3, 2, [1], 2)), (False, (4, 3, [2], 2)), (True, (1, 0, [], 1)),
                                           (False, (50, 49, [], 1))):
        a, b, ab, rk, rhos = s._bh(i_t, i_s0, hist, order, corr)
        c = uo.bh_coefficients(s.sigmas, i_t, i_s0, hist, order, "bh2", corr)
        assert (a, b, ab) == (float(c["a"]), float(c["b"]), float(c["ab"]))
        assert rhos == [float(r) for r in c["rhos"]] and (not hist or rk == float(c["rks"][0]))


def test_embed_deduplicates_per_token_timesteps():
    """WanModel.embed (CPU: no kernels involved): t [B, seq_len] with few distinct values -> one embedded row per value
    plus an int32 row index; gathering the rows reproduces the reference's materialised expansion (model.py:460-468)."""
    import torch
    mdl = importlib.import_module("univid_b200.wan.modules.model")
    torch.manual_seed(0)
    model = mdl.WanModel(model_type="ti2v", dim=256, ffn_dim=256, num_heads=2, num_layers=1, text_len=8, text_dim=32,
                         freq_dim=32, in_dim=4, out_dim=4).eval()
    lat = [torch.randn(4, 2, 4, 4), torch.randn(4, 2, 4, 4)]
    ctx = [torch.randn(5, 32), torch.randn(3, 32)]
    L = 10
    t = torch.full((2, L), 500.0)
    t[0, :4] = 0.0
    t[1, :4] = 250.0
    with torch.no_grad():
        x, e, kw = model.embed(lat, t, ctx, L)
        assert e.shape == (1, 3, 256) and kw["e"].shape == (1, 3, 6, 256) and kw["e_index"].dtype == torch.int32
        model.max_distinct_timesteps = 0
        x0, e_full, kw_full = model.embed(lat, t, ctx, L)
        assert e_full.shape == (2, L, 256) and kw_full["e"].shape == (2, L, 6, 256) and "e_index" not in kw_full
    idx = kw["e_index"].long()
    assert torch.allclose(model.token_embedding(e, kw["e_index"]), e_full, atol=1e-6)
    assert torch.allclose(kw["e"][0][idx], kw_full["e"], atol=1e-6)
    assert torch.equal(x, x0)
    assert model.token_embedding(e_full, None) is e_full
    # more distinct values than the cap: the reference's expansion is kept
    model.max_distinct_timesteps = 2
    with torch.no_grad():
        _, _, kw2 = model.embed(lat, t, ctx, L)
    assert "e_index" not in kw2 and kw2["e"].shape == (2, L, 6, 256)


def test_every_kernel_wrapper_refuses_cpu_tensors():
    """No CPU fallback anywhere on the product path: each tensor-level wrapper of the C ABI raises on CPU tensors
    instead of computing something else (the drop-in modules then fail loudly on a machine without the GPU)."""
    import torch
    from univid_b200 import _ext
    bf = torch.bfloat16
    x = torch.zeros(1, 4, 256)
    with pytest.raises(RuntimeError, match="no CPU path"):
        _ext.linear(x.to(bf), torch.zeros(256, 256, dtype=bf))
    with pytest.raises(RuntimeError, match="no CPU path"):
        _ext.block_glue(x, scale=torch.zeros(1, 1, 256), shift=torch.zeros(1, 1, 256))
    with pytest.raises(RuntimeError, match="no CPU path"):
        _ext.head_scatter(torch.zeros(1, 4, 2, 128, dtype=bf), 2)
    with pytest.raises(RuntimeError, match="no CPU path"):
        _ext.unipc_step(torch.zeros(8), None, torch.zeros(8), None, None, None, _ext.UnipcCoef())
    q = torch.zeros(1, 4, 2, 128, dtype=bf)
    with pytest.raises(RuntimeError, match="no CPU path"):
        _ext.fmha_fwd(q, q, q)
    with pytest.raises(RuntimeError, match="no CPU path"):
        _ext.qk_norm_rope(x.to(bf), None, torch.ones(256), None, 1e-6, 2)
    # the modules built on them: a CPU call of the attention classes cannot silently run another implementation
    mdl = importlib.import_module("univid_b200.wan.modules.model")
    sa = mdl.WanSelfAttention(256, 2)
    with pytest.raises((RuntimeError, AssertionError)):
        sa(x, torch.tensor([4]), torch.tensor([[1, 2, 2]]), mdl.rope_params(1024, 128)[:, :64])
