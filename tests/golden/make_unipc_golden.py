"""Generate tests/golden/unipc_golden.pt: the UNMODIFIED reference FlowUniPCMultistepScheduler
(models/wan/utils/fm_solvers_unipc.py) driven like the sampling loop of models/wan/textimage2video.py:367-394
(CFG combine, then scheduler.step) on seeded stand-in model outputs, CPU fp32.

    python tests/golden/make_unipc_golden.py        (needs /root/reference)
"""
import os
import sys
import warnings

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref_loader  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "unipc_golden.pt")
SHAPE = (1, 16, 2, 6, 8)
CASES = {            # name -> (sampling_steps, shift, solver_order, guide_scale)
    "t2v_50": (50, 5.0, 2, 5.0),      # the product's defaults (textimage2video.py:168-171)
    "short_8": (8, 3.0, 2, 3.5),
    "order1_6": (6, 5.0, 1, 5.0),
    "single": (1, 5.0, 2, 5.0),
}


def model_outputs(name, steps):
    """Seeded stand-ins for (noise_pred_cond, noise_pred_uncond) of every step and the initial noise."""
    g = torch.Generator().manual_seed(sum(map(ord, name)))
    noise = torch.randn(SHAPE, generator=g)
    return noise, [(torch.randn(SHAPE, generator=g), torch.randn(SHAPE, generator=g)) for _ in range(steps)]


def main():
    assert ref_loader.available(), "reference tree not found"
    warnings.simplefilter("ignore")
    cls = ref_loader.load_unipc_scheduler()
    gold = {}
    for name, (steps, shift, order, guide) in CASES.items():
        sch = cls(num_train_timesteps=1000, shift=1, use_dynamic_shifting=False, solver_order=order)
        sch.set_timesteps(steps, device="cpu", shift=shift)
        x, outs = model_outputs(name, steps)
        samples = []
        for t, (vc, vu) in zip(sch.timesteps, outs):
            noise_pred = vu + guide * (vc - vu)                                         # textimage2video.py:385-386
            x = sch.step(noise_pred, t, x, return_dict=False)[0]                        # :388-393
            samples.append(x.clone())
        gold[name] = dict(timesteps=sch.timesteps.clone(), sigmas=sch.sigmas.clone(),
                          samples=torch.stack(samples[:3] + samples[-2:]) if steps > 5 else torch.stack(samples),
                          sums=torch.tensor([float(s.double().sum()) for s in samples], dtype=torch.float64))
    torch.save(gold, OUT)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
