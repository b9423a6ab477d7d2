"""CPU restatement of the KL autoencoder (TEST INFRASTRUCTURE).

Upstream: ``monai-generative`` ``generative/networks/nets/autoencoderkl.py`` (classes
Upsample, Downsample, ResBlock, Encoder, Decoder, AutoencoderKL) -- not installable
here, version unpinned (``requirements.txt:12``) -> **parity unpinned against upstream,
pinned against its in-tree ancestor**: golden vectors produced by the reference's
``src/models/ae_kl.py`` (tests/golden/make_golden_aekl.py, tests/test_oracle_aekl.py).  Restated
from the published algorithm as specified in SURVEY.md section 8a rows 10-12 and
cross-checked against the in-tree ancestor ``/root/reference/src/models/ae_kl.py``:
ResBlock ``:48-80``, Downsample pad-right-1 + stride-2 ``:33-45``, Upsample
nearest x2 + conv ``:20-30``, final GroupNorm -> conv with no SiLU ``:173-176,236-239``,
encode clamp/exp ``:259-267``, sampling ``:269-272``, decode ``:279-282``.
Differences from the ancestor that follow upstream: ``norm_num_groups`` is a
parameter (1 in every reference config), channels are the absolute ``num_channels``
list, attention is optional (off in every reference config; not implemented here).

Call sites: ``src/train_autoencoderkl.py:133,204`` (forward), ``src/train_ldm.py:103-104,148``
(encode_stage_2_inputs), ``src/sample_trials.py:100,166`` (decode_stage_2_outputs).

State-dict naming follows MONAI (``Convolution(conv_only=True)`` wraps the conv as a
child called ``conv``): ``encoder.blocks.N.conv.{weight,bias}``,
``encoder.blocks.N.{norm1,norm2}.*``, ``encoder.blocks.N.{conv1,conv2,nin_shortcut}.conv.*``,
down/upsample ``...blocks.N.conv.conv.*``, ``quant_conv_mu.conv.*``,
``quant_conv_log_sigma.conv.*``, ``post_quant_conv.conv.*`` (unverified: no checkpoint ships).
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Dict, List

import torch
import torch.nn.functional as F

DEFAULT_CFG = dict(  # config/config_aekl_eeg_2_2_4_spec.yaml:19-30
    spatial_dims=1,
    in_channels=1,
    out_channels=1,
    num_channels=[2, 2, 4],
    latent_channels=1,
    num_res_blocks=2,
    norm_num_groups=1,
    attention_levels=[False, False, False],
    with_encoder_nonlocal_attn=False,
    with_decoder_nonlocal_attn=False,
)


def full_cfg(**over) -> dict:
    cfg = dict(DEFAULT_CFG)
    cfg.update(over)
    return cfg


def _check(cfg):
    if cfg.get("spatial_dims", 1) != 1:
        raise ValueError("only spatial_dims=1")
    if any(cfg.get("attention_levels", [])) or cfg.get("with_encoder_nonlocal_attn") or cfg.get(
            "with_decoder_nonlocal_attn"):
        raise ValueError("attention in the autoencoder is off in every reference config; not restated")


def _nres(cfg) -> List[int]:
    n = cfg["num_res_blocks"]
    return [n] * len(cfg["num_channels"]) if isinstance(n, int) else list(n)


def aekl_plan(cfg: dict) -> dict:
    """Encoder / Decoder block lists (upstream Encoder.__init__ / Decoder.__init__)."""
    _check(cfg)
    nc, nres, z = list(cfg["num_channels"]), _nres(cfg), cfg["latent_channels"]
    enc: List[dict] = [dict(kind="conv", prefix="encoder.blocks.0", cin=cfg["in_channels"], cout=nc[0], k=3)]
    out_ch = nc[0]
    for i in range(len(nc)):
        in_ch, out_ch = out_ch, nc[i]
        for _ in range(nres[i]):
            enc.append(dict(kind="res", prefix=f"encoder.blocks.{len(enc)}", cin=in_ch, cout=out_ch))
            in_ch = out_ch
        if i != len(nc) - 1:
            enc.append(dict(kind="down", prefix=f"encoder.blocks.{len(enc)}", ch=in_ch))
    enc.append(dict(kind="norm", prefix=f"encoder.blocks.{len(enc)}", ch=nc[-1]))
    enc.append(dict(kind="conv", prefix=f"encoder.blocks.{len(enc)}", cin=nc[-1], cout=z, k=3))

    rc, rres = nc[::-1], nres[::-1]
    dec: List[dict] = [dict(kind="conv", prefix="decoder.blocks.0", cin=z, cout=rc[0], k=3)]
    out_ch = rc[0]
    for i in range(len(rc)):
        in_ch, out_ch = out_ch, rc[i]
        for _ in range(rres[i]):
            dec.append(dict(kind="res", prefix=f"decoder.blocks.{len(dec)}", cin=in_ch, cout=out_ch))
            in_ch = out_ch
        if i != len(rc) - 1:
            dec.append(dict(kind="up", prefix=f"decoder.blocks.{len(dec)}", ch=in_ch))
    dec.append(dict(kind="norm", prefix=f"decoder.blocks.{len(dec)}", ch=in_ch))
    dec.append(dict(kind="conv", prefix=f"decoder.blocks.{len(dec)}", cin=in_ch, cout=cfg["out_channels"], k=3))
    return dict(encoder=enc, decoder=dec)


def aekl_param_shapes(cfg: dict) -> "OrderedDict[str, tuple]":
    plan = aekl_plan(cfg)
    z = cfg["latent_channels"]
    out: "OrderedDict[str, tuple]" = OrderedDict()

    def conv(p, i, o, k):
        out[p + ".conv.weight"] = (o, i, k)
        out[p + ".conv.bias"] = (o,)

    def gn(p, c):
        out[p + ".weight"] = (c,)
        out[p + ".bias"] = (c,)

    def blocks(lst):
        for l in lst:
            p = l["prefix"]
            if l["kind"] == "conv":
                conv(p, l["cin"], l["cout"], l["k"])
            elif l["kind"] == "res":
                gn(p + ".norm1", l["cin"])
                conv(p + ".conv1", l["cin"], l["cout"], 3)
                gn(p + ".norm2", l["cout"])
                conv(p + ".conv2", l["cout"], l["cout"], 3)
                if l["cin"] != l["cout"]:
                    conv(p + ".nin_shortcut", l["cin"], l["cout"], 1)
            elif l["kind"] in ("down", "up"):
                conv(p + ".conv", l["ch"], l["ch"], 3)
            elif l["kind"] == "norm":
                gn(p, l["ch"])

    blocks(plan["encoder"])
    blocks(plan["decoder"])
    conv("quant_conv_mu", z, z, 1)
    conv("quant_conv_log_sigma", z, z, 1)
    conv("post_quant_conv", z, z, 1)
    return out


def make_aekl_state_dict(cfg: dict, seed: int = 42) -> "OrderedDict[str, torch.Tensor]":
    """Seeded weights with PyTorch's default Conv1d init statistics
    (U(-1/sqrt(fan_in), 1/sqrt(fan_in)) for weight and bias; GroupNorm 1/0 perturbed
    so that the affine parameters are exercised)."""
    g = torch.Generator().manual_seed(seed)
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    shapes = aekl_param_shapes(cfg)
    for name, shape in shapes.items():
        if ".conv." in name or name.endswith("conv.weight") or name.endswith("conv.bias"):
            wshape = shapes[name.rsplit(".", 1)[0] + ".weight"]
            bound = 1.0 / math.sqrt(wshape[1] * wshape[2])
            t = (torch.rand(shape, generator=g) * 2 - 1) * bound
        elif name.endswith("weight"):
            t = 1.0 + 0.1 * torch.randn(shape, generator=g)
        else:
            t = 0.1 * torch.randn(shape, generator=g)
        sd[name] = t.float().contiguous()
    return sd


# --------------------------------------------------------------------- functional
def _conv(x, sd, p, padding=1, stride=1):
    return F.conv1d(x, sd[p + ".conv.weight"], sd[p + ".conv.bias"], stride=stride, padding=padding)


def _gn(x, sd, p, groups):
    return F.group_norm(x, groups, sd[p + ".weight"], sd[p + ".bias"], eps=1e-6)


def _run(x, sd, blocks, groups):
    for l in blocks:
        p, k = l["prefix"], l["kind"]
        if k == "conv":
            x = _conv(x, sd, p, padding=l["k"] // 2)
        elif k == "res":       # ae_kl.py:66-80
            h = _conv(F.silu(_gn(x, sd, p + ".norm1", groups)), sd, p + ".conv1")
            h = _conv(F.silu(_gn(h, sd, p + ".norm2", groups)), sd, p + ".conv2")
            if l["cin"] != l["cout"]:
                x = _conv(x, sd, p + ".nin_shortcut", padding=0)
            x = x + h
        elif k == "down":      # ae_kl.py:41-45: pad right by 1, k3 stride 2 padding 0
            x = _conv(F.pad(x, (0, 1), mode="constant", value=0.0), sd, p + ".conv", padding=0, stride=2)
        elif k == "up":        # ae_kl.py:27-30: nearest x2 then k3 conv
            x = _conv(F.interpolate(x, scale_factor=2.0, mode="nearest"), sd, p + ".conv")
        elif k == "norm":      # final GroupNorm, no SiLU before the last conv (ae_kl.py:173-176)
            x = _gn(x, sd, p, groups)
    return x


def encode(cfg, sd, x):
    """-> (z_mu, z_sigma); ae_kl.py:259-267."""
    plan = aekl_plan(cfg)
    h = _run(x, sd, plan["encoder"], cfg["norm_num_groups"])
    z_mu = _conv(h, sd, "quant_conv_mu", padding=0)
    z_log_var = torch.clamp(_conv(h, sd, "quant_conv_log_sigma", padding=0), -30.0, 20.0)
    return z_mu, torch.exp(z_log_var / 2)


def sampling(z_mu, z_sigma, eps=None):
    """z = mu + eps * sigma (ae_kl.py:269-272); eps may be supplied for reproducibility."""
    if eps is None:
        eps = torch.randn_like(z_sigma)
    return z_mu + eps * z_sigma


def decode(cfg, sd, z):
    """ae_kl.py:279-282."""
    plan = aekl_plan(cfg)
    return _run(_conv(z, sd, "post_quant_conv", padding=0), sd, plan["decoder"], cfg["norm_num_groups"])


def forward(cfg, sd, x, eps=None):
    z_mu, z_sigma = encode(cfg, sd, x)
    return decode(cfg, sd, sampling(z_mu, z_sigma, eps)), z_mu, z_sigma


def kl_loss(z_mu, z_sigma):
    """src/train_autoencoderkl.py:210-211 (sum over channel dim, then sum / B)."""
    kl = 0.5 * torch.sum(z_mu.pow(2) + z_sigma.pow(2) - torch.log(z_sigma.pow(2)) - 1, dim=[1])
    return torch.sum(kl) / kl.shape[0]
