"""CPU oracle of the sampler update around the DiT -- TEST INFRASTRUCTURE ONLY.

Restates, in plain PyTorch CPU ops and in the reference's order of floating-point operations:
  * the classifier-free-guidance combine  noise_pred = uncond + g * (cond - uncond)      (models/wan/textimage2video.py:380-386)
  * FlowUniPCMultistepScheduler            models/wan/utils/fm_solvers_unipc.py
      - training sigmas / shift             :108-122
      - set_timesteps                       :162-229
      - convert_model_output (flow, x0)     :281-333
      - UniP predictor  (B(h), bh1 / bh2)   :352-486
      - UniC corrector                      :488-628
      - step                                :657-741
for the configuration the product uses (textimage2video.py:335-341: solver_order 2, predict_x0, bh2, flow_prediction,
lower_order_final, final sigma 0, no thresholding, no dynamic shifting).  Pinned bit-exactly to the unmodified
reference class by tests/golden/make_unipc_golden.py / tests/test_unipc_oracle_golden.py.  Only tests/, smoke() and
bench.py's CPU legs may import this module.
"""
import numpy as np
import torch


def cfg_combine(cond, uncond, guide_scale):
    """textimage2video.py:385-386."""
    return uncond + guide_scale * (cond - uncond)


def training_sigmas(num_train_timesteps=1000, shift=1.0):
    """fm_solvers_unipc.py:108-118: sigma_k = 1 - alpha_k, alphas = linspace(1, 1/T, T) reversed, then the shift map."""
    alphas = np.linspace(1, 1 / num_train_timesteps, num_train_timesteps)[::-1].copy()
    sigmas = torch.from_numpy(1.0 - alphas).to(dtype=torch.float32)
    return shift * sigmas / (1 + (shift - 1) * sigmas)


def sampling_schedule(num_inference_steps, shift, num_train_timesteps=1000, init_shift=1.0):
    """set_timesteps (:162-229): (int64 timesteps [n], fp32 sigmas [n + 1] with a final 0)."""
    train = training_sigmas(num_train_timesteps, init_shift)
    sigma_max, sigma_min = train[0].item(), train[-1].item()
    sigmas = np.linspace(sigma_max, sigma_min, num_inference_steps + 1).copy()[:-1]
    sigmas = shift * sigmas / (1 + (shift - 1) * sigmas)
    timesteps = sigmas * num_train_timesteps
    sigmas = np.concatenate([sigmas, [0]]).astype(np.float32)
    return torch.from_numpy(timesteps).to(dtype=torch.int64), torch.from_numpy(sigmas)


def _lambda(sigma):
    return torch.log(1 - sigma) - torch.log(sigma)          # alpha = 1 - sigma (:274-275)


def bh_coefficients(sigmas, i_t, i_s0, history_idx, order, solver_type="bh2", corrector=False):
    """The scalar part of the UniP / UniC update from sigma index i_s0 to i_t (:395-455 / :549-606), all in fp32
    torch scalars exactly like the reference.  history_idx: sigma indices of the older model outputs (one per extra
    order).  Returns dict(a, b, ab, rks, rhos): x_t = a * x - b * m0 - ab * (sum_k rhos[k] * D1[k] (+ rho_last * D1_t))."""
    sigma_t, sigma_s0 = sigmas[i_t], sigmas[i_s0]
    alpha_t = 1 - sigma_t
    h = _lambda(sigma_t) - _lambda(sigma_s0)
    rks = []
    for si in history_idx[:order - 1]:
        rks.append((_lambda(sigmas[si]) - _lambda(sigma_s0)) / h)
    rks_all = torch.tensor(rks + [1.0])
    hh = -h                                                      # predict_x0
    h_phi_1 = torch.expm1(hh)
    h_phi_k = h_phi_1 / hh - 1
    b_h = hh if solver_type == "bh1" else torch.expm1(hh)
    factorial_i = 1
    R, b = [], []
    for i in range(1, order + 1):
        R.append(torch.pow(rks_all, i - 1))
        b.append(h_phi_k * factorial_i / b_h)
        factorial_i *= i + 1
        h_phi_k = h_phi_k / hh - 1 / factorial_i
    R, b = torch.stack(R), torch.tensor(b)
    if corrector:
        rhos = torch.tensor([0.5]) if order == 1 else torch.linalg.solve(R, b)
    elif order == 1:
        rhos = torch.zeros(0)
    elif order == 2:
        rhos = torch.tensor([0.5])
    else:
        rhos = torch.linalg.solve(R[:-1, :-1], b[:-1])
    return dict(a=sigma_t / sigma_s0, b=alpha_t * h_phi_1, ab=alpha_t * b_h, rks=rks, rhos=rhos)


class UniPCOracle:
    """Functional restatement of FlowUniPCMultistepScheduler.step for solver_order <= 3."""

    def __init__(self, num_train_timesteps=1000, solver_order=2, solver_type="bh2", init_shift=1.0,
                 lower_order_final=True, history_bf16=False):
        """history_bf16: restate what the reference computes inside torch.amp.autocast('cuda', bfloat16), where the
        product runs its sampling loop (textimage2video.py:330-331): torch.einsum over the history terms (:471, :614)
        is an autocast op, so rho and D1 are rounded to bf16 and the result is bf16; in the predictor that bf16 tensor
        then meets the 0-dim fp32 CPU tensor alpha_t * B_h (the reference keeps its sigmas on the host, :228-229) in a
        multiply whose result type is bf16 (a zero-dim operand of the same category does not promote; as a scalar
        operand it enters in fp32).  Pinned on the GPU box against the staged
        reference scheduler run under CUDA autocast (tests/test_unipc_gpu.py); the CPU reference does not take this
        route (einsum is not a CPU autocast op), so there is no CPU golden for it."""
        self.T, self.solver_order, self.solver_type = num_train_timesteps, solver_order, solver_type
        self.init_shift, self.lower_order_final = init_shift, lower_order_final
        self.history_bf16 = history_bf16

    def set_timesteps(self, num_inference_steps, shift=1.0):
        self.timesteps, self.sigmas = sampling_schedule(num_inference_steps, shift, self.T, self.init_shift)
        self.model_outputs = [None] * self.solver_order      # oldest ... newest (x0 predictions)
        self.lower_order_nums = 0
        self.last_sample = None
        self.step_index = None
        self.this_order = None

    @staticmethod
    def _einsum_bf16(rho, d1):
        """One term of einsum('k,bkc...->bc...') under bf16 autocast: operands rounded to bf16, product rounded to bf16."""
        return (rho.to(torch.bfloat16).float() * d1.to(torch.bfloat16).float()).to(torch.bfloat16)

    def _d1s(self, m0, rks):
        return [(self.model_outputs[-(k + 2)] - m0) / rk for k, rk in enumerate(rks)]

    def step(self, model_output, timestep, sample):
        if self.step_index is None:                               # :630-655 (unique timesteps: first match)
            idx = (self.timesteps == timestep).nonzero()
            self.step_index = idx[1 if len(idx) > 1 else 0].item()
        i = self.step_index
        m_t = sample - self.sigmas[i] * model_output              # convert_model_output (:320-323)
        if i > 0 and self.last_sample is not None:                # UniC (:688-699, :549-628)
            order = self.this_order
            m0 = self.model_outputs[-1]
            c = bh_coefficients(self.sigmas, i, i - 1, [i - 1 - k for k in range(1, order)], order, self.solver_type,
                                corrector=True)
            x_t_ = c["a"] * self.last_sample - c["b"] * m0
            corr = 0
            for rho, d1 in zip(c["rhos"][:-1], self._d1s(m0, c["rks"])):
                corr = corr + (self._einsum_bf16(rho, d1).float() if self.history_bf16 else rho * d1)
            sample = x_t_ - c["ab"] * (corr + c["rhos"][-1] * (m_t - m0))
        self.model_outputs = self.model_outputs[1:] + [m_t]       # :707-712
        order = min(self.solver_order, len(self.timesteps) - i) if self.lower_order_final else self.solver_order
        self.this_order = min(order, self.lower_order_nums + 1)   # :714-723
        self.last_sample = sample
        p = bh_coefficients(self.sigmas, i + 1, i, [i - k for k in range(1, self.this_order)], self.this_order,
                            self.solver_type)                      # UniP (:395-486)
        x_t_ = p["a"] * sample - p["b"] * m_t
        pred = 0
        for rho, d1 in zip(p["rhos"], self._d1s(m_t, p["rks"])):
            pred = pred + (self._einsum_bf16(rho, d1) if self.history_bf16 else rho * d1)
        if self.history_bf16 and torch.is_tensor(pred):
            prev = x_t_ - (p["ab"].float() * pred.float()).to(torch.bfloat16).float()   # scalar (fp32) x bf16 -> bf16
        else:
            prev = x_t_ - p["ab"] * pred
        if self.lower_order_nums < self.solver_order:
            self.lower_order_nums += 1
        self.step_index += 1
        return prev
