"""The oracle's latent-diffusion training step (oracle/ldm_train.py) against golden losses / gradients produced by torch autograd
through the REFERENCE's own UNetModel (tests/golden/make_golden_ldm_train.py).  CPU only."""
import os

import numpy as np
import pytest
import torch

from oracle import ldm_train as ol
from oracle import unet as ou
from oracle.schedulers import DDPMScheduler

from conftest import GOLDEN
import sys
sys.path.insert(0, GOLDEN)
from train_cases import FULL_GRADS, TRAIN_CASES  # noqa: E402

_G = np.load(os.path.join(GOLDEN, "ldm_train_golden.npz"))


def golden_case(name):
    over, B, T, pred, schedule, (b0, b1) = TRAIN_CASES[name]
    cfg = ou.full_cfg(**over)
    sd = ou.make_unet_state_dict(cfg, seed=0)
    sched = DDPMScheduler(1000, b0, b1, schedule, pred)
    z0, noise, t = (torch.from_numpy(_G[f"{name}/{k}"]) for k in ("z0", "noise", "t"))
    return cfg, sd, sched, z0, noise, t, float(_G[name + "/loss"])


def check_against_golden(name, grads, rtol, floor):
    """grads: {key: tensor} -> compared with the stored gradients (small cases) or their digests (full network).
    Absolute floor: `floor` x the tensor's largest reference entry, plus 1e-6 x the largest gradient entry of the whole model
    (some gradients are exactly zero in exact arithmetic -- a conv bias in front of a GroupNorm whose groups are single channels
    -- and come out as rounding noise on both sides)."""
    if name in FULL_GRADS:
        gmax = max(float(np.abs(_G[f"{name}/grad/{k}"]).max()) for k in grads)
    else:
        gmax = max(float(np.abs(_G[f"{name}/digest/{k}/top"]).max()) for k in grads)
    for k, g in grads.items():
        g = g.detach().cpu()
        if name in FULL_GRADS:
            ref = torch.from_numpy(_G[f"{name}/grad/{k}"])
            torch.testing.assert_close(g, ref, rtol=rtol, atol=floor * float(ref.abs().max()) + 1e-6 * gmax, msg=lambda m: f"{k}: {m}")
        else:
            f = g.flatten()
            norm = float(_G[f"{name}/digest/{k}/norm"])
            head = torch.from_numpy(_G[f"{name}/digest/{k}/head"])
            idx = torch.from_numpy(_G[f"{name}/digest/{k}/top_idx"])
            top = torch.from_numpy(_G[f"{name}/digest/{k}/top"])
            scale = float(top.abs().max())
            assert abs(float(f.double().norm()) - norm) <= 2 * rtol * norm + 1e-6 * gmax * f.numel() ** 0.5, k
            torch.testing.assert_close(f[:32], head, rtol=rtol, atol=floor * scale + 1e-6 * gmax, msg=lambda m: f"{k} head: {m}")
            torch.testing.assert_close(f[idx], top, rtol=rtol, atol=floor * scale + 1e-6 * gmax, msg=lambda m: f"{k} top: {m}")


@pytest.mark.parametrize("name", [n for n in TRAIN_CASES if n != "ldm_full"])
def test_oracle_training_step_matches_reference_autograd(name):
    torch.set_num_threads(4)
    cfg, sd, sched, z0, noise, t, loss_ref = golden_case(name)
    loss, grads, _ = ol.ldm_train_step(cfg, sd, z0, noise, t, sched, lr=0.0)
    assert abs(loss - loss_ref) <= 1e-5 * abs(loss_ref)
    check_against_golden(name, grads, 1e-4, 1e-5)


def test_oracle_adam_step_moves_every_parameter():
    cfg, sd, sched, z0, noise, t, _ = golden_case("small_eps")
    _, grads, new = ol.ldm_train_step(cfg, sd, z0, noise, t, sched, lr=1e-4)
    for k in sd:   # first Adam step: |delta| = lr * |g| / (|g| + eps) ~ lr wherever the gradient is non-zero
        d = (new[k] - sd[k]).abs()
        nz = grads[k].abs() > 1e-6
        assert torch.all(d[nz] > 0.9e-4) and torch.all(d <= 1.01e-4), k
