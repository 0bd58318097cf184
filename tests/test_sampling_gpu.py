"""The denoising loop around the DiT on the GPU (univid_b200/wan/textimage2video.py):
one B = 2 classifier-free-guidance forward against the two B = 1 forwards of the reference loop
(models/wan/textimage2video.py:380-383), with and without UniVid's per-call text weights, and a short sampling loop
against the oracle scheduler driven by the same model outputs.  -m gpu."""
import importlib

import pytest
import torch

from oracle import unipc_oracle as uo

pytestmark = pytest.mark.gpu


def _cos(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a @ b) / (a.norm() * b.norm() + 1e-30))


@pytest.fixture(scope="module")
def setup():
    mdl = importlib.import_module("univid_b200.wan.modules.model")
    torch.manual_seed(0)
    model = mdl.WanModel(model_type="t2v", dim=256, ffn_dim=512, num_heads=2, num_layers=2, text_len=64, text_dim=128,
                         freq_dim=64, in_dim=16, out_dim=16).cuda().eval()
    torch.nn.init.normal_(model.head.head.weight, std=0.05)          # the reference zero-inits the head: make it count
    g = torch.Generator(device="cuda").manual_seed(1)
    latent = torch.randn(16, 3, 8, 12, device="cuda", generator=g)   # 3 x 4 x 6 = 72 tokens
    ctx = torch.randn(40, 128, device="cuda", generator=g)
    ctx_null = torch.randn(25, 128, device="cuda", generator=g)
    return model, latent, ctx, ctx_null, 72


@pytest.mark.parametrize("calls", [None, (0, 5)])
def test_batched_cfg_forward_equals_two_forwards(setup, calls):
    t2v = importlib.import_module("univid_b200.wan.textimage2video")
    tma = importlib.import_module("univid_b200.tma")
    weights = None if calls is None else tuple(tma.calculate_text_weight(c, tma.TextWeightConfig()) for c in calls)
    model, latent, ctx, ctx_null, seq_len = setup
    t = torch.tensor([700.0], device="cuda")
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        cond, uncond = t2v.cfg_batched_forward(model, latent, t, ctx, ctx_null, seq_len, weights, text_len=128)
        if weights is None:
            want_c = model([latent], t=t, context=[ctx], seq_len=seq_len)[0]
            want_u = model([latent], t=t, context=[ctx_null], seq_len=seq_len)[0]
        else:
            # the reference hook, through the fused schedule: call 0 -> w(0), call 1 -> w(1)
            cfg = tma.TextWeightConfig()
            assert weights[0] == 1.3 and abs(weights[1] - 1.2561) < 1e-4
            with tma.FusedTextWeightSchedule(model, cfg) as sched:
                sched.call_index = calls[0]
                want_c = model([latent], t=t, context=[ctx], seq_len=seq_len)[0]
                sched.call_index = calls[1]
                want_u = model([latent], t=t, context=[ctx_null], seq_len=seq_len)[0]
    for got, want in ((cond, want_c), (uncond, want_u)):
        assert got.shape == latent.shape and got.dtype == torch.float32
        assert (got - want).abs().max().item() <= 2e-2 and _cos(got, want) >= 0.9999
    assert (cond - uncond).abs().max() > 1e-3                          # the two branches really differ


def test_short_sampling_loop_follows_the_oracle_scheduler(setup):
    """4 UniPC steps, batched CFG + fused scheduler kernel: at every step the next latent equals the oracle scheduler
    fed with the SAME two model outputs (bit-exact: the update is op-by-op rounded, with the bf16 roundings of the
    history term that bf16 autocast implies)."""
    t2v = importlib.import_module("univid_b200.wan.textimage2video")
    sched_mod = importlib.import_module("univid_b200.wan.utils.fm_solvers_unipc")
    model, latent, ctx, ctx_null, seq_len = setup
    sch = sched_mod.FlowUniPCMultistepScheduler(num_train_timesteps=1000, shift=1, use_dynamic_shifting=False)
    sch.set_timesteps(4, device="cuda", shift=5.0)
    o = uo.UniPCOracle(history_bf16=True)      # the loop runs under bf16 autocast like the product's: bf16 history einsum
    o.set_timesteps(4, shift=5.0)
    x, xo = latent, latent.cpu().unsqueeze(0)
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        for k, t in enumerate(sch.timesteps):
            cond, uncond = t2v.cfg_batched_forward(model, x, t2v.expand_timestep(t, seq_len), ctx, ctx_null, seq_len)
            x_next = sch.step_cfg(cond.unsqueeze(0), uncond.unsqueeze(0), 5.0, t, x.unsqueeze(0))[0].squeeze(0)
            xo = o.step(uo.cfg_combine(cond.cpu().unsqueeze(0), uncond.cpu().unsqueeze(0), 5.0), o.timesteps[k],
                        x.cpu().unsqueeze(0))
            assert torch.equal(x_next.cpu(), xo.squeeze(0)), k
            x = x_next
        # the packaged loop gives the same final latent (same kernels, same order)
        sch2 = sched_mod.FlowUniPCMultistepScheduler(num_train_timesteps=1000, shift=1, use_dynamic_shifting=False)
        final = t2v.sample_loop(model, sch2, latent, ctx, ctx_null, seq_len, guide_scale=5.0, sampling_steps=4, shift=5.0,
                                batch_cfg=True)
    assert torch.equal(final, x)
    assert torch.isfinite(final).all()


def test_expand_timestep_matches_reference_expression():
    t2v = importlib.import_module("univid_b200.wan.textimage2video")
    t = torch.tensor(833)
    assert t2v.expand_timestep(t, 10).tolist() == [833]
    mask = torch.ones(2, 4, 6)
    mask[0] = 0                                                        # first latent frame is given (ti2v)
    ts = t2v.expand_timestep(t, 14, mask)                             # 2 x 2 x 3 = 12 tokens + 2 padding
    want = torch.cat([(mask[:, ::2, ::2] * t).flatten(), torch.ones(2) * t]).unsqueeze(0)   # textimage2video.py:372-377
    assert torch.equal(ts, want) and ts.shape == (1, 14)
