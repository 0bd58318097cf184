"""Temperature Modality Alignment: UniVid's per-DiT-call text weighting of the cross-attention context.

Reference: Wan22ContextWrapper in models/model_pipeline.py -- the schedule `_calculate_text_weight`
(:1699-1735), the DiT-forward counter `hooked_dit_forward` (:1856-1866) and the cross-attention hook
that multiplies context[:, :text_len] by the weight (:1756-1803).

The reference wrapper works unchanged on the drop-in modules (it discovers them by the class name
'WanCrossAttention' and pre-scales the context; the default forward path then runs).  This module is
the FUSED alternative: the same schedule and counter, but instead of materialising a scaled context
for every layer of every DiT call it hands (text_weight, text_len) to WanCrossAttention.forward,
which folds the weight into the k-norm prologue and the attention kernel.
"""
import math
from dataclasses import dataclass


@dataclass
class TextWeightConfig:
    """The text-weight knobs of CrossAttentionConfig (model_pipeline.py:169, :201-206)."""
    use_dynamic_text_weight: bool = True
    text_weight_max: float = 1.3
    text_weight_min: float = 1.0
    text_weight_schedule: str = "cosine"      # linear | cosine | exponential
    text_weight_transition_ratio: float = 0.4
    total_sampling_steps: int = 50
    bagel_sequence_length: int = 128


def calculate_text_weight(call_index, config):
    """Weight for the `call_index`-th DiT forward of a generation (model_pipeline.py:1699-1735).
    Note the index counts DiT forwards, not sampler steps: with classifier-free guidance there are two
    per step (textimage2video.py:380-383)."""
    if not config.use_dynamic_text_weight:
        return 1.0
    transition = int(config.total_sampling_steps * config.text_weight_transition_ratio)
    if call_index >= transition:
        return config.text_weight_min
    progress = call_index / max(transition, 1)
    lo, hi = config.text_weight_min, config.text_weight_max
    if config.text_weight_schedule == "linear":
        return hi - (hi - lo) * progress
    if config.text_weight_schedule == "cosine":
        return lo + (hi - lo) * (1 + math.cos(math.pi * progress)) / 2
    if config.text_weight_schedule == "exponential":
        return lo + (hi - lo) * math.exp(-5 * progress)
    return 1.0


def text_len_for(context, config):
    """min(bagel_sequence_length, seq_len // 2) (model_pipeline.py:1789)."""
    seq_len = context.shape[1] if context.dim() > 1 else context.shape[0]
    return min(config.bagel_sequence_length, seq_len // 2)


class TextWeightCounter:
    """The DiT-call counter of hooked_dit_forward (model_pipeline.py:1856-1866) on its own: `next_weight()` returns
    the weight of the next DiT call and advances the counter.  Used by the batched-CFG denoise step, which issues the
    conditional and the unconditional call of a sampler step as one B = 2 forward and therefore needs both weights
    up front (wan/textimage2video.py::cfg_batched_forward)."""

    def __init__(self, config=None):
        self.config = config or TextWeightConfig()
        self.call_index = 0

    def next_weight(self):
        w = calculate_text_weight(self.call_index, self.config)
        self.call_index += 1
        return w


class FusedTextWeightSchedule:
    """Arms every WanCrossAttention of `dit_model` with the fused text weighting.

    Usage mirrors Wan22ContextWrapper.generate: `with FusedTextWeightSchedule(dit, cfg): pipeline.generate()`.
    Each DiT forward advances the call counter and sets the weight; each cross-attention forward passes
    (text_weight, text_len) to the fused path.  `injection_layers` restricts the layers like the
    reference's `injection_layers` (model_pipeline.py:1763).
    """

    def __init__(self, dit_model, config=None, injection_layers=None):
        self.dit_model = dit_model
        self.config = config or TextWeightConfig()
        self.injection_layers = injection_layers
        self.call_index = 0
        self.text_weight_multiplier = 1.0
        self._saved = []

    def set_timestep(self, call_index):
        self.text_weight_multiplier = calculate_text_weight(call_index, self.config)

    def __enter__(self):
        self.call_index = 0
        sched = self
        layer = 0
        for name, module in self.dit_model.named_modules():
            if module.__class__.__name__ != 'WanCrossAttention':
                continue
            inner = module.forward

            def fused_forward(x, context, context_lens, *args, _inner=inner, _layer=layer, **kwargs):
                w = sched.text_weight_multiplier
                active = sched.injection_layers is None or _layer in sched.injection_layers
                if active and sched.config.use_dynamic_text_weight and w != 1.0 and context is not None:
                    kwargs.setdefault('text_weight', w)
                    kwargs.setdefault('text_len', text_len_for(context, sched.config))
                return _inner(x, context, context_lens, *args, **kwargs)

            self._saved.append((module, 'forward' in module.__dict__, module.__dict__.get('forward')))
            module.forward = fused_forward
            layer += 1

        dit_inner = self.dit_model.forward

        def counted_forward(*args, **kwargs):
            sched.set_timestep(sched.call_index)
            sched.call_index += 1
            return dit_inner(*args, **kwargs)

        self._saved.append((self.dit_model, 'forward' in self.dit_model.__dict__,
                            self.dit_model.__dict__.get('forward')))
        self.dit_model.forward = counted_forward
        return self

    def __exit__(self, *exc):
        for module, had, old in reversed(self._saved):
            if had:
                module.forward = old
            else:
                del module.__dict__['forward']
        self._saved.clear()
        return False
