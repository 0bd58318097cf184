"""Drop-in for models/wan/utils/fm_solvers_unipc.py (logical path wan/utils/fm_solvers_unipc.py): the flow-matching
UniPC scheduler that advances the latent between two DiT forwards.

Same class name, constructor arguments, `set_timesteps` / `step` signatures and state attributes (`timesteps`,
`sigmas`, `model_outputs`, `last_sample`, `step_index`, `this_order`, `lower_order_nums`) as the reference, for the
configuration the product runs (textimage2video.py:335-341: solver_order <= 2, predict_x0, flow_prediction, bh1 /
bh2, final sigma 0).  The scalar schedule arithmetic stays on the host in the reference's own fp32 torch ops
(fm_solvers_unipc.py:395-455, :549-606); the elementwise update of the latent -- ~25 eager kernels in the reference
-- is ONE kernel (uvb_unipc_step), which with `step_cfg` also absorbs the classifier-free-guidance combine of
textimage2video.py:385-386.  Options the kernel does not cover raise NotImplementedError; CPU tensors raise.
"""
import types

import numpy as np
import torch

from ... import _ext

__all__ = ['FlowUniPCMultistepScheduler', 'SchedulerOutput']


class SchedulerOutput:
    """Stand-in for diffusers.schedulers.scheduling_utils.SchedulerOutput: `.prev_sample`."""

    def __init__(self, prev_sample):
        self.prev_sample = prev_sample


def _lam(sigma):
    return torch.log(1 - sigma) - torch.log(sigma)


class FlowUniPCMultistepScheduler:
    order = 1

    def __init__(self,
                 num_train_timesteps=1000,
                 solver_order=2,
                 prediction_type="flow_prediction",
                 shift=1.0,
                 use_dynamic_shifting=False,
                 thresholding=False,
                 dynamic_thresholding_ratio=0.995,
                 sample_max_value=1.0,
                 predict_x0=True,
                 solver_type="bh2",
                 lower_order_final=True,
                 disable_corrector=[],
                 solver_p=None,
                 timestep_spacing="linspace",
                 steps_offset=0,
                 final_sigmas_type="zero"):
        if solver_type not in ("bh1", "bh2"):
            if solver_type in ("midpoint", "heun", "logrho"):
                solver_type = "bh2"                                   # fm_solvers_unipc.py:99-101
            else:
                raise NotImplementedError(f"{solver_type} is not implemented for {self.__class__}")
        unsupported = [name for name, bad in (
            ("solver_order > 2", solver_order not in (1, 2)), ("prediction_type", prediction_type != "flow_prediction"),
            ("use_dynamic_shifting", use_dynamic_shifting), ("thresholding", thresholding),
            ("predict_x0=False", not predict_x0), ("solver_p", solver_p is not None),
            ("final_sigmas_type", final_sigmas_type != "zero")) if bad]
        if unsupported:
            raise NotImplementedError("univid_b200 UniPC step does not cover: " + ", ".join(unsupported))
        self.config = types.SimpleNamespace(
            num_train_timesteps=num_train_timesteps, solver_order=solver_order, prediction_type=prediction_type,
            shift=shift, use_dynamic_shifting=use_dynamic_shifting, thresholding=thresholding,
            dynamic_thresholding_ratio=dynamic_thresholding_ratio, sample_max_value=sample_max_value,
            predict_x0=predict_x0, solver_type=solver_type, lower_order_final=lower_order_final,
            disable_corrector=disable_corrector, solver_p=solver_p, timestep_spacing=timestep_spacing,
            steps_offset=steps_offset, final_sigmas_type=final_sigmas_type)
        self.predict_x0 = predict_x0
        # "auto": mirror the reference under the ambient autocast state (bf16 einsum of the history terms inside
        # torch.amp.autocast('cuda', bfloat16), fp32 otherwise); "fp32" / "bf16" pin it.  Not a reference argument.
        self.history_dtype = "auto"
        self.num_inference_steps = None
        alphas = np.linspace(1, 1 / num_train_timesteps, num_train_timesteps)[::-1].copy()
        sigmas = torch.from_numpy(1.0 - alphas).to(dtype=torch.float32)
        sigmas = shift * sigmas / (1 + (shift - 1) * sigmas)
        self.sigmas = sigmas.to("cpu")
        self.timesteps = sigmas * num_train_timesteps
        self.model_outputs = [None] * solver_order
        self.timestep_list = [None] * solver_order
        self.lower_order_nums = 0
        self.disable_corrector = disable_corrector
        self.solver_p = None
        self.last_sample = None
        self.this_order = None
        self._step_index = None
        self._begin_index = None
        self.sigma_min = self.sigmas[-1].item()
        self.sigma_max = self.sigmas[0].item()

    @property
    def step_index(self):
        return self._step_index

    @property
    def begin_index(self):
        return self._begin_index

    def set_begin_index(self, begin_index=0):
        self._begin_index = begin_index

    def set_timesteps(self, num_inference_steps=None, device=None, sigmas=None, mu=None, shift=None):
        """fm_solvers_unipc.py:162-229: linspace over the training sigma range, shift map, final sigma 0."""
        if sigmas is None:
            sigmas = np.linspace(self.sigma_max, self.sigma_min, num_inference_steps + 1).copy()[:-1]
        if shift is None:
            shift = self.config.shift
        sigmas = shift * sigmas / (1 + (shift - 1) * sigmas)
        timesteps = sigmas * self.config.num_train_timesteps
        sigmas = np.concatenate([sigmas, [0]]).astype(np.float32)
        self.sigmas = torch.from_numpy(sigmas)                       # stays on the host: scalar arithmetic only
        self._timesteps_host = torch.from_numpy(timesteps).to(dtype=torch.int64)
        self.timesteps = self._timesteps_host.to(device=device)
        self.num_inference_steps = len(timesteps)
        self.model_outputs = [None] * self.config.solver_order
        self.timestep_list = [None] * self.config.solver_order
        self.lower_order_nums = 0
        self.last_sample = None
        self.this_order = None
        self._step_index = None
        self._begin_index = None

    def index_for_timestep(self, timestep, schedule_timesteps=None):
        ts = self._timesteps_host if schedule_timesteps is None else schedule_timesteps.cpu()
        t = int(timestep)                                              # one device->host read on the first step only
        indices = (ts == t).nonzero()
        return indices[1 if len(indices) > 1 else 0].item()

    def _init_step_index(self, timestep):
        self._step_index = self.index_for_timestep(timestep) if self.begin_index is None else self._begin_index

    def _bh(self, i_t, i_s0, history_idx, order, corrector):
        """Scalar part of UniP / UniC from sigma index i_s0 to i_t (:395-455 / :549-606) in fp32 torch scalars:
        (a, b, ab, rk, rhos) with x_t = a x - b m0 - ab (rhos . D1s [+ rho_last D1_t])."""
        sigma_t, sigma_s0 = self.sigmas[i_t], self.sigmas[i_s0]
        alpha_t = 1 - sigma_t
        h = _lam(sigma_t) - _lam(sigma_s0)
        rks = [(_lam(self.sigmas[si]) - _lam(sigma_s0)) / h for si in history_idx]
        rks_all = torch.tensor(rks + [1.0])
        hh = -h
        h_phi_1 = torch.expm1(hh)
        h_phi_k = h_phi_1 / hh - 1
        b_h = hh if self.config.solver_type == "bh1" else torch.expm1(hh)
        factorial_i = 1
        rows, b = [], []
        for i in range(1, order + 1):
            rows.append(torch.pow(rks_all, i - 1))
            b.append(h_phi_k * factorial_i / b_h)
            factorial_i *= i + 1
            h_phi_k = h_phi_k / hh - 1 / factorial_i
        if corrector:
            rhos = torch.tensor([0.5]) if order == 1 else torch.linalg.solve(torch.stack(rows), torch.tensor(b))
        else:
            rhos = torch.tensor([0.5]) if order == 2 else torch.zeros(0)
        return (float(sigma_t / sigma_s0), float(alpha_t * h_phi_1), float(alpha_t * b_h),
                float(rks[0]) if rks else 1.0, [float(r) for r in rhos])

    def _advance(self, cond, uncond, guide_scale, timestep, sample):
        if self.num_inference_steps is None:
            raise ValueError("Number of inference steps is 'None', you need to run 'set_timesteps' after creating the scheduler")
        if not (torch.is_tensor(sample) and sample.is_cuda):
            raise RuntimeError("univid_b200 UniPC step runs on CUDA tensors; there is no CPU path")
        if self.step_index is None:
            self._init_step_index(timestep)
        i = self._step_index
        coef = _ext.UnipcCoef()
        # Inside torch.amp.autocast('cuda', bfloat16) -- where the product runs the sampling loop, textimage2video.py:
        # 330-331 -- the reference's einsum over the history terms executes in bf16 on the GPU; mirror its roundings.
        # (history_dtype = "fp32" / "bf16" overrides the detection.)
        if self.history_dtype == "auto":
            coef.history_bf16 = int(torch.is_autocast_enabled('cuda') and torch.get_autocast_dtype('cuda') == torch.bfloat16)
        else:
            coef.history_bf16 = int(self.history_dtype == "bf16")
        coef.guide_scale = float(guide_scale)
        coef.sigma = float(self.sigmas[i])
        use_corrector = i > 0 and (i - 1) not in self.disable_corrector and self.last_sample is not None
        m0, m1 = self.model_outputs[-1], self.model_outputs[-2] if self.config.solver_order > 1 else None
        if use_corrector:
            oc = self.this_order
            a, b, ab, rk, rhos = self._bh(i, i - 1, [i - 1 - k for k in range(1, oc)], oc, corrector=True)
            coef.corrector_order, coef.c_a, coef.c_b, coef.c_ab, coef.c_rk = oc, a, b, ab, rk
            coef.c_rho0, coef.c_rho_last = (rhos[0] if oc == 2 else 0.0), rhos[-1]
        else:
            coef.corrector_order = 0
        if self.config.lower_order_final:
            this_order = min(self.config.solver_order, len(self._timesteps_host) - i)
        else:
            this_order = self.config.solver_order
        self.this_order = min(this_order, self.lower_order_nums + 1)
        assert self.this_order > 0
        a, b, ab, rk, rhos = self._bh(i + 1, i, [i - k for k in range(1, self.this_order)], self.this_order,
                                      corrector=False)
        coef.predictor_order, coef.p_a, coef.p_b, coef.p_ab, coef.p_rk = self.this_order, a, b, ab, rk
        coef.p_rho0 = rhos[0] if self.this_order == 2 else 0.0
        m_t, x_c, x_next = _ext.unipc_step(
            cond, uncond, sample, self.last_sample if use_corrector else None,
            m0 if (use_corrector or self.this_order == 2) else None,
            m1 if (use_corrector and coef.corrector_order == 2) else None, coef)
        for k in range(self.config.solver_order - 1):
            self.model_outputs[k] = self.model_outputs[k + 1]
            self.timestep_list[k] = self.timestep_list[k + 1]
        self.model_outputs[-1] = m_t
        self.timestep_list[-1] = timestep
        self.last_sample = x_c
        if self.lower_order_nums < self.config.solver_order:
            self.lower_order_nums += 1
        self._step_index += 1
        return x_next.view(sample.shape)

    def step(self, model_output, timestep, sample, return_dict=True, generator=None):
        """fm_solvers_unipc.py:657-741.  model_output / sample: fp32 CUDA tensors of one shape."""
        prev = self._advance(model_output, None, 0.0, timestep, sample)
        return SchedulerOutput(prev_sample=prev) if return_dict else (prev,)

    def step_cfg(self, noise_pred_cond, noise_pred_uncond, guide_scale, timestep, sample, return_dict=False):
        """Opt-in fused form of textimage2video.py:385-393:
            step(noise_pred_uncond + guide_scale * (noise_pred_cond - noise_pred_uncond), timestep, sample)
        without materialising the combined prediction."""
        prev = self._advance(noise_pred_cond, noise_pred_uncond, guide_scale, timestep, sample)
        return SchedulerOutput(prev_sample=prev) if return_dict else (prev,)

    def scale_model_input(self, sample, *args, **kwargs):
        return sample

    def __len__(self):
        return self.config.num_train_timesteps
