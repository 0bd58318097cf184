"""One sampler step of the 1.3B config at 32 760 tokens: the two B = 1 DiT forwards of the reference loop
(textimage2video.py:380-383) against ONE B = 2 forward (wan/textimage2video.py::cfg_batched_forward), plus the fused
CFG + UniPC update kernel against its eager expression.  Run under gpurun:
    python scripts/bench_cfg_batch.py > gpurun_out/cfg_batch.log"""
import importlib
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

mdl = importlib.import_module("univid_b200.wan.modules.model")
t2v = importlib.import_module("univid_b200.wan.textimage2video")
sched_mod = importlib.import_module("univid_b200.wan.utils.fm_solvers_unipc")


def timed(fn, n):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    cfg = bench.CONFIGS["1.3B"]
    dev = torch.device("cuda")
    f, h, w = cfg["grid"]
    L = f * h * w
    torch.manual_seed(0)
    with torch.device(dev):
        model = mdl.WanModel(model_type="t2v", dim=cfg["dim"], ffn_dim=cfg["ffn"], num_heads=cfg["heads"],
                             num_layers=cfg["layers"], text_len=cfg["text_len"], in_dim=16, out_dim=16).eval()
    g = torch.Generator(device=dev).manual_seed(7)
    lat = torch.randn(16, f, h * 2, w * 2, device=dev, generator=g)
    ctx = torch.randn(cfg["text_len"], 4096, device=dev, generator=g)
    ctx0 = torch.randn(300, 4096, device=dev, generator=g)
    t = torch.tensor([500.0], device=dev)
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        two = timed(lambda: (model([lat], t=t, context=[ctx], seq_len=L), model([lat], t=t, context=[ctx0], seq_len=L)), 2)
        one = timed(lambda: t2v.cfg_batched_forward(model, lat, t, ctx, ctx0, L), 2)
    print(f"two B=1 DiT forwards: {two:.2f} ms   one B=2 forward: {one:.2f} ms   ratio {one / two:.4f}")

    # per-token timesteps [1, seq_len] (textimage2video.py:372-377; ti2v: first latent frame at t = 0): one embedded row
    # per distinct value + a row index, against the reference's materialised [1, L, 6, C] modulation
    tt = torch.full((1, L), 500.0, device=dev)
    tt[0, :h * w] = 0.0
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        scalar = timed(lambda: model([lat], t=t, context=[ctx], seq_len=L), 2)
        dedup = timed(lambda: model([lat], t=tt, context=[ctx], seq_len=L), 2)
        model.max_distinct_timesteps = 0
        expanded = timed(lambda: model([lat], t=tt, context=[ctx], seq_len=L), 2)
        model.max_distinct_timesteps = 8
    print(f"DiT forward: scalar t {scalar:.2f} ms   per-token t, de-duplicated rows {dedup:.2f} ms   "
          f"per-token t, materialised [1, L, 6, C] {expanded:.2f} ms")

    # sampler update: fused kernel vs the eager chain (CFG combine + the scheduler's elementwise ops)
    sch = sched_mod.FlowUniPCMultistepScheduler(num_train_timesteps=1000, shift=1, use_dynamic_shifting=False)
    sch.set_timesteps(50, device=dev, shift=5.0)
    x = lat.unsqueeze(0)
    vc, vu = torch.randn_like(x), torch.randn_like(x)
    for k in range(3):                                      # reach the steady state (corrector + order 2)
        x = sch.step_cfg(vc, vu, 5.0, sch.timesteps[k], x)[0]
    state = (sch._step_index, sch.lower_order_nums, list(sch.model_outputs), sch.last_sample, sch.this_order)

    def fused():
        sch._step_index, sch.lower_order_nums, sch.model_outputs, sch.last_sample, sch.this_order = (
            state[0], state[1], list(state[2]), state[3], state[4])
        sch.step_cfg(vc, vu, 5.0, sch.timesteps[3], x)

    m0, m1, last = state[2][-1], state[2][-2], state[3]

    def eager():            # the tensor ops of one steady-state step, as the reference issues them
        v = vu + 5.0 * (vc - vu)
        m_t = x - 0.9 * v
        x_t_ = 0.95 * last - 0.05 * m0
        d1 = (m1 - m0) / -1.4
        xc = x_t_ - 0.05 * (0.07 * d1 + 0.48 * (m_t - m0))
        p_ = 0.95 * xc - 0.05 * m_t
        return p_ - 0.05 * (0.5 * ((m0 - m_t) / -1.4))

    from univid_b200 import _ext
    coef = _ext.UnipcCoef()
    coef.guide_scale, coef.sigma, coef.corrector_order, coef.predictor_order = 5.0, 0.9, 2, 2
    coef.c_a, coef.c_b, coef.c_ab, coef.c_rk, coef.c_rho0, coef.c_rho_last = 0.95, -0.05, -0.05, -1.4, 0.07, 0.48
    coef.p_a, coef.p_b, coef.p_ab, coef.p_rk, coef.p_rho0 = 0.95, -0.05, -0.05, -1.4, 0.5
    kern = lambda: _ext.unipc_step(vc, vu, x, last, m0, m1, coef)
    fu, ke, ea = timed(fused, 50), timed(kern, 200), timed(eager, 50)
    n = x.numel()
    print(f"sampler update on {n} latent elements: scheduler.step_cfg (host schedule scalars + 1 kernel) {fu * 1e3:.1f} us; "
          f"the kernel alone {ke * 1e3:.1f} us = {36 * n / ke * 1e-6:.0f} GB/s of {36 * n / 1e6:.1f} MB; "
          f"the reference's tensor ops alone (22 eager kernels, no host scalars) {ea * 1e3:.1f} us")


if __name__ == "__main__":
    main()
