"""CPU restatement of the LDM sampling loop (TEST INFRASTRUCTURE).

Follows ``/root/reference/src/sample_trials.py:136-169``: DDIM timesteps, per step
``model_output = unet(x, timesteps=[t])`` then ``x, _ = scheduler.step(model_output, t, x)``,
finally ``decode_stage_2_outputs(x / scale_factor)`` and the ``[36:-36]`` crop.
Batched over B windows (the reference uses B = 1 per seed; windows are independent).
"""
from __future__ import annotations

import torch

from . import aekl as _aekl
from . import unet as _unet
from .schedulers import DDIMScheduler

SAMPLER_DEFAULTS = dict(  # sample_trials.py:136-143
    num_train_timesteps=1000, beta_start=0.0015, beta_end=0.0205,
    schedule="scaled_linear_beta", prediction_type="v_prediction", clip_sample=False)


@torch.no_grad()
def ddim_sample(unet_cfg, unet_sd, noise, n_steps=50, aekl_cfg=None, aekl_sd=None, scale_factor=1.0,
                sched_kwargs=None, crop=0, return_latent=False):
    sched = DDIMScheduler(**(sched_kwargs or SAMPLER_DEFAULTS))
    sched.set_timesteps(n_steps)
    x = noise.clone()
    for t in sched.timesteps:
        out = _unet.unet_forward(unet_cfg, unet_sd, x, torch.tensor([int(t)], dtype=torch.long))
        x, _ = sched.step(out, int(t), x)
    if aekl_cfg is None or return_latent:
        return x
    y = _aekl.decode(aekl_cfg, aekl_sd, x / scale_factor)
    if crop:
        y = y[:, :, crop:-crop]
    return y
