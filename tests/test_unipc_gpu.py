"""Fused sampler update on the GPU (uvb_unipc_step behind the drop-in FlowUniPCMultistepScheduler; SURVEY.md sec. 8f
rank 3) against the CPU oracle (pinned bit-exactly to the reference scheduler by tests/test_unipc_oracle_golden.py)
and the frozen outputs of the reference sampling loop.  The update is a chain of individually rounded fp32
operations, so the bar is BIT-EXACT against the oracle evaluated on this machine; against the golden file (made on
another CPU, whose libm may round log / expm1 of the schedule scalars differently) it is 1e-5.  -m gpu."""
import importlib
import os

import pytest
import torch

from oracle import unipc_oracle as uo
from tests.golden.make_unipc_golden import CASES, SHAPE, model_outputs

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _sched(order=2):
    mod = importlib.import_module("univid_b200.wan.utils.fm_solvers_unipc")
    return mod.FlowUniPCMultistepScheduler(num_train_timesteps=1000, shift=1, use_dynamic_shifting=False,
                                           solver_order=order)


@pytest.mark.parametrize("name", list(CASES))
@pytest.mark.parametrize("fused_cfg", [True, False])
def test_sampling_loop_matches_oracle_bit_exactly(name, fused_cfg):
    steps, shift, order, guide = CASES[name]
    gold = torch.load(os.path.join(ROOT, "tests", "golden", "unipc_golden.pt"), map_location="cpu", weights_only=False)[name]
    sch = _sched(order)
    sch.set_timesteps(steps, device="cuda", shift=shift)
    o = uo.UniPCOracle(solver_order=order)
    o.set_timesteps(steps, shift=shift)
    assert torch.equal(sch.timesteps.cpu(), o.timesteps) and torch.equal(sch.sigmas, o.sigmas)
    x0, outs = model_outputs(name, steps)
    x, xo = x0.cuda(), x0
    kept = []
    for k, (t, (vc, vu)) in enumerate(zip(sch.timesteps, outs)):
        x_in, x_before = x, x.clone()
        if fused_cfg:
            x = sch.step_cfg(vc.cuda(), vu.cuda(), guide, t, x)[0]
        else:
            x = sch.step(uo.cfg_combine(vc, vu, guide).cuda(), t, x, return_dict=False)[0]
        xo = o.step(uo.cfg_combine(vc, vu, guide), o.timesteps[k], xo)
        assert torch.equal(x_in, x_before)                          # the caller's sample is never written
        assert x.shape == x0.shape and x.dtype == torch.float32
        assert torch.equal(x.cpu(), xo), f"step {k}: max diff {(x.cpu() - xo).abs().max()}"
        assert sch.this_order == o.this_order and sch.step_index == o.step_index
        kept.append(x.cpu())
    kept = torch.stack(kept[:3] + kept[-2:]) if steps > 5 else torch.stack(kept)
    assert (kept - gold["samples"]).abs().max() <= 1e-5


def test_full_size_latent_and_ragged_tail():
    """The 1.3B latent (16 x 21 x 60 x 104) and a length that is not a multiple of 4, against the oracle."""
    for shape in ((1, 16, 21, 60, 104), (1, 3, 7, 11)):
        sch, o = _sched(), uo.UniPCOracle()
        sch.set_timesteps(4, device="cuda", shift=5.0)
        o.set_timesteps(4, shift=5.0)
        g = torch.Generator().manual_seed(len(shape))
        x = torch.randn(shape, generator=g)
        xg = x.cuda()
        for k, t in enumerate(sch.timesteps):
            vc, vu = torch.randn(shape, generator=g), torch.randn(shape, generator=g)
            xg = sch.step_cfg(vc.cuda(), vu.cuda(), 5.0, t, xg)[0]
            x = o.step(uo.cfg_combine(vc, vu, 5.0), o.timesteps[k], x)
        assert torch.equal(xg.cpu(), x)


def test_step_returns_the_reference_types_and_rejects_cpu():
    sch = _sched()
    sch.set_timesteps(3, device="cuda", shift=5.0)
    x = torch.randn(SHAPE, device="cuda")
    out = sch.step(torch.randn_like(x), sch.timesteps[0], x)
    assert hasattr(out, "prev_sample") and out.prev_sample.shape == x.shape
    with pytest.raises(RuntimeError):
        sch.step(torch.randn(SHAPE), sch.timesteps[1], torch.randn(SHAPE))


def test_history_bf16_mode_matches_the_oracle_bit_exactly():
    """Inside torch.amp.autocast('cuda', bfloat16) the drop-in mirrors the bf16 einsum of the history terms (the
    product's configuration); outside it stays on the fp32 chain.  Both against the oracle."""
    steps, shift = 12, 5.0
    for autocast in (True, False):
        sch, o = _sched(), uo.UniPCOracle(history_bf16=autocast)
        sch.set_timesteps(steps, device="cuda", shift=shift)
        o.set_timesteps(steps, shift=shift)
        g = torch.Generator().manual_seed(3)
        x = torch.randn(SHAPE, generator=g)
        xg = x.cuda()
        for k, t in enumerate(sch.timesteps):
            v = torch.randn(SHAPE, generator=g)
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
                xg = sch.step(v.cuda(), t, xg, return_dict=False)[0]
            x = o.step(v, o.timesteps[k], x)
            assert xg.dtype == torch.float32
            assert torch.equal(xg.cpu(), x), (autocast, k, (xg.cpu() - x).abs().max())


def test_history_bf16_mode_matches_the_reference_scheduler_under_cuda_autocast():
    """The UNMODIFIED reference FlowUniPCMultistepScheduler (staged under oracle/_ref) run on the GPU inside
    torch.amp.autocast('cuda', bfloat16) -- the way textimage2video.py:330-331 runs it -- against the drop-in under the
    same context: every step of a 20-step schedule.  The reference evaluates its scalar coefficients on the GPU
    (device libm), the drop-in on the host, so the bar is 1e-5 relative to the sample scale rather than bit equality;
    the bf16 roundings themselves are 4e-3 apart from the fp32 chain, i.e. the test separates the two modes."""
    from oracle import ref_loader
    if not ref_loader.available():
        pytest.skip("reference sources not staged (python oracle/make_ref.py)")
    Ref = ref_loader.load_unipc_scheduler()
    steps, shift = 20, 5.0
    ref = Ref(num_train_timesteps=1000, shift=1, use_dynamic_shifting=False)
    ref.set_timesteps(steps, device="cuda", shift=shift)
    sch, sch32 = _sched(), _sched()
    sch.set_timesteps(steps, device="cuda", shift=shift)
    sch32.set_timesteps(steps, device="cuda", shift=shift)
    sch32.history_dtype = "fp32"
    g = torch.Generator().manual_seed(4)
    x = torch.randn(SHAPE, generator=g).cuda()
    xr, xs, x32 = x, x, x
    worst, worst32 = 0.0, 0.0
    for t in ref.timesteps:
        v = torch.randn(SHAPE, generator=g).cuda()
        with torch.autocast("cuda", dtype=torch.bfloat16):
            xr = ref.step(v, t, xr, return_dict=False)[0]
            xs = sch.step(v, t, xs, return_dict=False)[0]
            x32 = sch32.step(v, t, x32, return_dict=False)[0]
        scale = xr.abs().max().item()
        worst = max(worst, (xr.float() - xs).abs().max().item() / scale)
        worst32 = max(worst32, (xr.float() - x32).abs().max().item() / scale)
    assert worst <= 1e-5, (worst, worst32)
    assert worst32 > 10 * max(worst, 1e-7), (worst, worst32)      # the fp32 chain is measurably NOT what autocast computes
