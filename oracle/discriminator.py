"""CPU restatement of the adversarial half of the autoencoder training step (TEST INFRASTRUCTURE).

Reference call sites: ``src/train_autoencoderkl.py:135-137`` (``PatchDiscriminator(**config.patchdiscriminator.params)``),
``:156`` (``PatchAdversarialLoss(criterion="least_squares")``), ``:213-214`` (generator term), ``:223-234`` (discriminator
step); parameters ``config/config_aekl_eeg.yaml:30-40`` (1-D, 3 layers, 64 channels, kernel 3, BatchNorm, no conv bias, padding 1).

Both classes live in ``monai-generative`` (``generative/networks/nets/patchgan_discriminator.py``,
``generative/losses/adversarial_loss.py``), which is absent from ``/root/reference`` and not installable here (version
unpinned, ``requirements.txt:12``) -> **parity unpinned** against upstream; restated from the published code:

* ``PatchDiscriminator`` is an ``nn.Sequential`` of MONAI ``Convolution`` blocks (conv -> norm -> dropout -> act, "NDA"):
  ``initial_conv`` (in -> C, stride 2, bias, LeakyReLU(0.2), no norm); layers ``0 .. n-1`` (C*2^l -> C*2^(l+1), stride 2
  except the last which has stride 1, ``bias=False``, BatchNorm, LeakyReLU(0.2)); ``final_conv`` (-> out_channels, stride 1,
  bias, conv only, padding ``(k-1)//2``).  ``forward`` returns the list of every block's output; the reference uses ``[-1]``.
  Weights: conv N(0, 0.02), BatchNorm weight N(1, 0.02), bias 0 (``initialise_weights``).
* ``PatchAdversarialLoss(criterion="least_squares")``: target tensor filled with 1.0 (real) / 0.0 (fake); unless
  ``no_activation_leastsq`` the logits first pass through ``LeakyReLU(0.05)``; loss = ``MSELoss(mean)``; a list of
  discriminator outputs is averaged (one output here).

state_dict keys (MONAI ``Convolution``: children ``conv`` and ``adn`` with ``N`` = the norm):
``initial_conv.conv.{weight,bias}``, ``{l}.conv.weight``, ``{l}.adn.N.{weight,bias,running_mean,running_var,num_batches_tracked}``,
``final_conv.conv.{weight,bias}`` (unverified: no checkpoint ships).
"""
from __future__ import annotations

from collections import OrderedDict

import torch
import torch.nn.functional as F

DEFAULT_CFG = dict(spatial_dims=1, num_layers_d=3, num_channels=64, in_channels=1, out_channels=1, kernel_size=3,
                   norm="BATCH", bias=False, padding=1)   # config/config_aekl_eeg.yaml:30-40
BN_EPS, BN_MOMENTUM, LEAKY = 1e-5, 0.1, 0.2


def full_cfg(**over) -> dict:
    cfg = dict(DEFAULT_CFG)
    cfg.update(over)
    return cfg


def disc_plan(cfg):
    """[(prefix, cin, cout, stride, padding, has_bias, has_norm, has_act)] in module order."""
    if cfg.get("spatial_dims", 1) != 1 or str(cfg.get("norm", "BATCH")).upper() != "BATCH" or cfg.get("bias", False):
        raise ValueError("restated for the reference's configuration: 1-D, BatchNorm, bias=False")
    k, pad, c = cfg["kernel_size"], cfg["padding"], cfg["num_channels"]
    n = cfg["num_layers_d"]
    plan = [("initial_conv", cfg["in_channels"], c, 2, pad, True, False, True)]
    cin, cout = c, 2 * c
    for l in range(n):
        plan.append((str(l), cin, cout, 1 if l == n - 1 else 2, pad, False, True, True))
        cin, cout = cout, 2 * cout
    plan.append(("final_conv", cin, cfg["out_channels"], 1, (k - 1) // 2, True, False, False))
    return plan


def disc_param_shapes(cfg) -> "OrderedDict[str, tuple]":
    out: "OrderedDict[str, tuple]" = OrderedDict()
    k = cfg["kernel_size"]
    for prefix, cin, cout, _s, _p, has_bias, has_norm, _a in disc_plan(cfg):
        out[prefix + ".conv.weight"] = (cout, cin, k)
        if has_bias:
            out[prefix + ".conv.bias"] = (cout,)
        if has_norm:
            out[prefix + ".adn.N.weight"] = (cout,)
            out[prefix + ".adn.N.bias"] = (cout,)
            out[prefix + ".adn.N.running_mean"] = (cout,)
            out[prefix + ".adn.N.running_var"] = (cout,)
            out[prefix + ".adn.N.num_batches_tracked"] = ()
    return out


def is_buffer(name: str) -> bool:
    return name.endswith(("running_mean", "running_var", "num_batches_tracked"))


def make_disc_state_dict(cfg, seed: int = 7, weight_std: float = 0.02) -> "OrderedDict[str, torch.Tensor]":
    """``initialise_weights`` statistics (conv N(0, 0.02), BatchNorm N(1, 0.02) / 0) with small random biases so that every
    parameter is exercised; ``weight_std`` can be raised for tests (0.02 makes the logits tiny)."""
    g = torch.Generator().manual_seed(seed)
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    for name, shape in disc_param_shapes(cfg).items():
        if name.endswith("conv.weight"):
            t = weight_std * torch.randn(shape, generator=g)
        elif name.endswith("conv.bias"):
            t = 0.05 * torch.randn(shape, generator=g)
        elif name.endswith("N.weight"):
            t = 1.0 + 0.02 * torch.randn(shape, generator=g)
        elif name.endswith("N.bias"):
            t = 0.05 * torch.randn(shape, generator=g)
        elif name.endswith("running_mean"):
            t = torch.zeros(shape)
        elif name.endswith("running_var"):
            t = torch.ones(shape)
        else:
            t = torch.zeros(shape, dtype=torch.long)
        sd[name] = t.contiguous() if t.dtype == torch.long else t.float().contiguous()
    return sd


def forward(cfg, sd, x, training=True, update_running=None):
    """PatchDiscriminator.forward(x) -> list of the outputs of every block (the reference takes ``[-1]``).
    ``training``: BatchNorm uses batch statistics (the training loop never calls ``discriminator.eval()``);
    ``update_running``: optional dict that receives the updated running statistics (momentum 0.1, unbiased variance)."""
    outs = []
    h = x
    for prefix, _cin, _cout, stride, pad, has_bias, has_norm, has_act in disc_plan(cfg):
        h = F.conv1d(h, sd[prefix + ".conv.weight"], sd.get(prefix + ".conv.bias") if has_bias else None, stride=stride, padding=pad)
        if has_norm:
            p = prefix + ".adn.N."
            if training:
                rm = sd[p + "running_mean"].detach().clone()
                rv = sd[p + "running_var"].detach().clone()
                h = F.batch_norm(h, rm, rv, sd[p + "weight"], sd[p + "bias"], training=True, momentum=BN_MOMENTUM, eps=BN_EPS)
                if update_running is not None:
                    update_running[p + "running_mean"], update_running[p + "running_var"] = rm, rv
            else:
                h = F.batch_norm(h, sd[p + "running_mean"], sd[p + "running_var"], sd[p + "weight"], sd[p + "bias"], training=False,
                                 eps=BN_EPS)
        if has_act:
            h = F.leaky_relu(h, LEAKY)
        outs.append(h)
    return outs


def patch_adversarial_loss(logits, target_is_real: bool, for_discriminator: bool, no_activation_leastsq: bool = False):
    """PatchAdversarialLoss(criterion="least_squares", reduction="mean")(logits, target_is_real, for_discriminator)."""
    if not for_discriminator and not target_is_real:
        target_is_real = True      # upstream: "with a generator loss the target is always real"
    if isinstance(logits, (list, tuple)):
        return torch.stack([patch_adversarial_loss(l, target_is_real, for_discriminator, no_activation_leastsq) for l in logits]).mean()
    y = logits if no_activation_leastsq else F.leaky_relu(logits, 0.05)
    target = torch.full_like(y, 1.0 if target_is_real else 0.0)
    return F.mse_loss(y, target)
