"""The denoising loop around the DiT (models/wan/textimage2video.py:357-394, the `t2v` sampling loop), restated for the
B200 path.  Only the loop: text encoder, VAE, prompt handling and model loading of
WanTI2V are outside the hot path and not reproduced.

Per sampler step the reference runs the DiT twice (conditional / unconditional context, :380-383), combines the two
predictions with the guidance scale (:385-386) and calls the scheduler (:388-393).  Here:
  * `cfg_denoise_step(..., batch_cfg=False)` does exactly that, with the CFG combine fused into the scheduler
    kernel (FlowUniPCMultistepScheduler.step_cfg);
  * `batch_cfg=True` runs both branches as ONE B = 2 forward (same latent twice, the two contexts stacked): every
    kernel of the block sees twice the rows per launch and the patch / time embeddings are computed once per sample
    in the same pass.  UniVid's text weighting counts DiT *calls* (model_pipeline.py:1856-1866), so within one
    sampler step the conditional call gets w(2k) and the unconditional one w(2k+1); the batched form reproduces that
    exactly by scaling the first `text_len` rows of each sample's embedded context with its own weight -- the same
    product the hook forms inside every cross-attention layer (model_pipeline.py:1789-1797).
"""
import torch

__all__ = ['expand_timestep', 'cfg_batched_forward', 'cfg_denoise_step', 'sample_loop']


def expand_timestep(t, seq_len, mask=None):
    """Per-token timesteps of textimage2video.py:372-377: tokens covered by `mask` (a [F, H, W] latent mask, sampled
    at the patch stride) carry t, the padding up to seq_len carries t; returns [1, seq_len].  mask=None is the t2v
    case (all ones), for which every token carries t and a [1] timestep gives identical embeddings."""
    t = t.reshape(1)
    if mask is None:
        return t
    ts = (mask[:, ::2, ::2] * t).flatten()
    ts = torch.cat([ts, ts.new_ones(seq_len - ts.size(0)) * t])
    return ts.unsqueeze(0)


def cfg_batched_forward(model, latent, timestep, context, context_null, seq_len, text_weights=None, text_len=128):
    """(noise_pred_cond, noise_pred_uncond) from ONE B = 2 forward of a univid_b200 WanModel.
    text_weights = (w_cond, w_uncond) scales rows [:text_len] of each branch's embedded context (the dynamic text
    weighting of Wan22ContextWrapper, per DiT call); None = no weighting."""
    t2 = timestep.expand(2, *timestep.shape[1:]) if timestep.dim() == 2 else timestep.reshape(1).expand(2)
    x, e, kwargs = model.embed([latent, latent], t2, [context, context_null], seq_len)
    if text_weights is not None and (text_weights[0] != 1.0 or text_weights[1] != 1.0):
        ctx = kwargs['context']
        n = min(text_len, ctx.size(1) // 2)                    # model_pipeline.py:1789
        w = torch.ones(2, ctx.size(1), 1, dtype=ctx.dtype, device=ctx.device)
        w[0, :n] = float(text_weights[0])
        w[1, :n] = float(text_weights[1])
        kwargs['context'] = ctx * w
    for block in model.blocks:
        x = block(x, **kwargs)
    x = model.head(x, model.token_embedding(e, kwargs.get('e_index')))
    out = model.unpatchify(x, kwargs['grid_sizes'])
    return out[0].float(), out[1].float()


def cfg_denoise_step(model, scheduler, latent, t, context, context_null, seq_len, guide_scale, batch_cfg=False,
                     mask=None, text_weight_schedule=None, text_len=128):
    """One iteration of the sampling loop (textimage2video.py:367-394): latent [C, F, H, W] fp32 -> next latent.
    text_weight_schedule: a univid_b200.tma.TextWeightCounter (or None); the batched form advances it by the two DiT
    calls of the step (the unbatched form leaves the weighting to whatever hook is armed on the model)."""
    timestep = expand_timestep(t, seq_len, mask)
    if batch_cfg:
        w = None
        if text_weight_schedule is not None:
            w = (text_weight_schedule.next_weight(), text_weight_schedule.next_weight())
        cond, uncond = cfg_batched_forward(model, latent, timestep, context, context_null, seq_len, w, text_len)
    else:
        cond = model([latent], t=timestep, context=[context], seq_len=seq_len)[0]
        uncond = model([latent], t=timestep, context=[context_null], seq_len=seq_len)[0]
    step_cfg = getattr(scheduler, 'step_cfg', None)
    if step_cfg is not None:
        nxt = step_cfg(cond.unsqueeze(0), uncond.unsqueeze(0), guide_scale, t, latent.unsqueeze(0))[0]
    else:
        noise_pred = uncond + guide_scale * (cond - uncond)
        nxt = scheduler.step(noise_pred.unsqueeze(0), t, latent.unsqueeze(0), return_dict=False)[0]
    return nxt.squeeze(0)


def sample_loop(model, scheduler, noise, context, context_null, seq_len, guide_scale=5.0, sampling_steps=50, shift=5.0,
                batch_cfg=False, mask=None, text_weight_schedule=None, text_len=128):
    """textimage2video.py:335-394 for the 'unipc' solver: noise [C, F, H, W] -> denoised latent."""
    scheduler.set_timesteps(sampling_steps, device=noise.device, shift=shift)
    latent = noise
    for t in scheduler.timesteps:
        latent = cfg_denoise_step(model, scheduler, latent, t, context, context_null, seq_len, guide_scale,
                                  batch_cfg=batch_cfg, mask=mask, text_weight_schedule=text_weight_schedule,
                                  text_len=text_len)
    return latent
