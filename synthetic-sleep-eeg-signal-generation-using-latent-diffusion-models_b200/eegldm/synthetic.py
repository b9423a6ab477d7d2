"""Synthetic configurations and seeded weights for benchmarks (no trained checkpoint ships with the reference, and there is
no network): the reference's configs as plain dicts, and a seeded initialiser that works on any ``state_dict`` by name and
shape.  Product-side helper: nothing here touches ``oracle/`` (which is test infrastructure)."""
from __future__ import annotations

import math
from collections import OrderedDict

import torch

# config/config_ldm.yaml:30-43 with in / out channels = latent_channels = 1 (src/sample_trials.py:71-76)
LDM_UNET_CFG = dict(image_size=768, in_channels=1, out_channels=1, model_channels=128, attention_resolutions=[8, 4],
                    num_res_blocks=2, channel_mult=[1, 2, 4], dropout=0.0, conv_resample=True, num_heads=1,
                    num_head_channels=-1, use_scale_shift_norm=False, resblock_updown=True)
# config/config_aekl_eeg_2_2_4_spec.yaml:19-30
AEKL_224_CFG = dict(spatial_dims=1, in_channels=1, out_channels=1, num_channels=[2, 2, 4], latent_channels=1, num_res_blocks=2,
                    norm_num_groups=1, attention_levels=[False, False, False], with_encoder_nonlocal_attn=False,
                    with_decoder_nonlocal_attn=False)
# src/sample_trials.py:136-143
DDIM_CFG = dict(num_train_timesteps=1000, beta_start=0.0015, beta_end=0.0205, schedule="scaled_linear_beta",
                prediction_type="v_prediction", clip_sample=False)


def seeded_state_dict(module: torch.nn.Module, seed: int = 0) -> "OrderedDict[str, torch.Tensor]":
    """Random weights for every entry of ``module.state_dict()``: fan-in scaled convolutions / linears (residual-branch
    outputs at half gain so the stream variance stays bounded through 21 ResBlocks), GroupNorm affine near (1, 0), small
    biases.  The tensors the reference zero-initialises (``zero_module``, unet.py:39-45) are random too -- a fresh reference
    UNet outputs exactly 0, which would make any measurement vacuous."""
    g = torch.Generator().manual_seed(seed)
    out: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    for name, ref in module.state_dict().items():
        shape = tuple(ref.shape)
        if len(shape) == 1:
            is_norm_w = name.endswith("weight")
            t = (1.0 + 0.1 * torch.randn(shape, generator=g)) if is_norm_w else 0.05 * torch.randn(shape, generator=g)
        else:
            fan_in = 1
            for s in shape[1:]:
                fan_in *= s
            gain = 0.5 if (".out_layers.3." in name or ".proj_out." in name or ".conv2." in name) else 1.0
            t = torch.randn(shape, generator=g) * (gain / math.sqrt(fan_in))
        out[name] = t.float().contiguous()
    return out
