"""CPU restatement of one latent-diffusion training step (TEST INFRASTRUCTURE).

Follows the batch body of ``train_epoch_ldm``, ``/root/reference/src/training/training.py:420-443``, with the scheduler of
``src/train_ldm.py:199-200`` and the optimiser of ``src/train_ldm.py:208``:

    noisy = scheduler.add_noise(e, noise, timesteps)            training.py:429
    pred  = model(noisy, timesteps)                             training.py:430   (oracle.unet.unet_forward_train)
    target = noise | scheduler.get_velocity(e, noise, t)        training.py:432-436
    loss = F.mse_loss(pred.float(), target.float())             training.py:437
    loss.backward(); Adam(lr).step()                            training.py:441-443

in fp32: the reference's ``autocast`` + ``GradScaler`` change the precision of its PyTorch path, not the function (the loss scale
cancels in ``scaler.step``); the fp32 value is what that path approximates.  The denoiser and its gradient are pinned against the
reference's own ``UNetModel`` by ``tests/golden/make_golden_ldm_train.py`` (``ldm_train_golden.npz``); ``add_noise`` /
``get_velocity`` are pinned against the in-tree ``src/models/ldm.py`` tables (``make_golden_sched.py``).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import unet as ou
from .schedulers import DDPMScheduler


def ldm_loss(cfg, params, z0, noise, timesteps, sched: DDPMScheduler):
    noisy = sched.add_noise(z0, noise, timesteps)
    pred = ou.unet_forward_train(cfg, params, noisy, timesteps)
    target = sched.get_velocity(z0, noise, timesteps) if sched.prediction_type == "v_prediction" else noise
    return F.mse_loss(pred.float(), target.float())


def ldm_train_step(cfg, sd, z0, noise, timesteps, sched: DDPMScheduler, lr=1e-4, betas=(0.9, 0.999), eps=1e-8):
    """-> (loss, {name: grad}, {name: updated parameter})"""
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    loss = ldm_loss(cfg, params, z0, noise, timesteps, sched)
    loss.backward()
    grads = {k: p.grad.clone() for k, p in params.items()}
    if lr > 0:
        torch.optim.Adam(list(params.values()), lr=lr, betas=betas, eps=eps).step()
    return float(loss), grads, {k: p.detach() for k, p in params.items()}
