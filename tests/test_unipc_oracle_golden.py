"""The sampler-update oracle (oracle/unipc_oracle.py: CFG combine + FlowUniPCMultistepScheduler restated, SURVEY.md
sec. 8f rank 3) against the outputs of the unmodified reference scheduler frozen in tests/golden/unipc_golden.pt
(make_unipc_golden.py), and live against the reference source while it is mounted."""
import os
import warnings

import pytest
import torch

from oracle import ref_loader
from oracle import unipc_oracle as uo
from tests.golden.make_unipc_golden import CASES, model_outputs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def gold():
    return torch.load(os.path.join(ROOT, "tests", "golden", "unipc_golden.pt"), map_location="cpu", weights_only=False)


@pytest.mark.parametrize("name", list(CASES))
def test_oracle_reproduces_reference_sampling_loop(gold, name):
    steps, shift, order, guide = CASES[name]
    o = uo.UniPCOracle(solver_order=order)
    o.set_timesteps(steps, shift=shift)
    assert torch.equal(o.timesteps, gold[name]["timesteps"]) and torch.equal(o.sigmas, gold[name]["sigmas"])
    x, outs = model_outputs(name, steps)
    samples = []
    for t, (vc, vu) in zip(o.timesteps, outs):
        x = o.step(uo.cfg_combine(vc, vu, guide), t, x)
        samples.append(x)
    sums = torch.tensor([float(s.double().sum()) for s in samples], dtype=torch.float64)
    assert torch.equal(sums, gold[name]["sums"])                      # every step, bit-exact
    kept = torch.stack(samples[:3] + samples[-2:]) if steps > 5 else torch.stack(samples)
    assert torch.equal(kept, gold[name]["samples"])


def test_schedule_known_answers():
    """KATs derived from fm_solvers_unipc.py:108-122, :162-229: 8 steps, shift 5."""
    t, s = uo.sampling_schedule(8, 5.0)
    assert t.tolist() == [999, 972, 937, 892, 833, 749, 624, 416]
    assert s.dtype == torch.float32 and s[-1] == 0 and abs(s[0].item() - 0.9998) < 1e-4
    # sigma' = shift * sigma / (1 + (shift - 1) * sigma) is monotone and maps (0, 1) into (0, 1)
    assert (s[:-1] > s[1:]).all()
    assert abs(uo.training_sigmas()[0].item() - 0.999) < 1e-6 and uo.training_sigmas()[-1].item() == 0.0


def test_first_step_is_first_order_and_last_step_drops_the_order():
    o = uo.UniPCOracle(solver_order=2)
    o.set_timesteps(4, shift=5.0)
    x = torch.zeros(1, 2, 1, 2, 2)
    orders = []
    for t in o.timesteps:
        x = o.step(torch.ones_like(x), t, x)
        orders.append(o.this_order)
    assert orders == [1, 2, 2, 1]            # warm-up (:722) and lower_order_final (:714-717)
    # with a constant velocity field v = 1 the exact flow solution is x(sigma) = x0 + (sigma - sigma0) * 1
    assert torch.allclose(x, torch.full_like(x, -float(o.sigmas[0])), atol=1e-5)


@pytest.mark.skipif(not ref_loader.available(), reason="reference tree not mounted")
@pytest.mark.parametrize("order,steps,shift", [(2, 12, 5.0), (1, 5, 1.0), (3, 9, 5.0)])
def test_live_reference_scheduler(order, steps, shift):
    warnings.simplefilter("ignore")
    cls = ref_loader.load_unipc_scheduler()
    ref = cls(num_train_timesteps=1000, shift=1, use_dynamic_shifting=False, solver_order=order)
    ref.set_timesteps(steps, device="cpu", shift=shift)
    o = uo.UniPCOracle(solver_order=order)
    o.set_timesteps(steps, shift=shift)
    g = torch.Generator().manual_seed(order * 100 + steps)
    x = torch.randn(1, 4, 2, 3, 5, generator=g)
    xo = x.clone()
    for t in ref.timesteps:
        v = torch.randn(x.shape, generator=g)
        x = ref.step(v, t, x, return_dict=False)[0]
        xo = o.step(v, t, xo)
        if order <= 2:
            assert torch.equal(x, xo)
        else:                                   # order 3 sums two history terms: einsum vs explicit sum, 1-2 ulp
            assert (x - xo).abs().max() <= 4e-6
