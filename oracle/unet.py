"""Functional CPU restatement of the reference denoiser (TEST INFRASTRUCTURE).

Follows ``/root/reference/src/models/unet.py``:

* ``timestep_embedding``            unet.py:12-36
* ``Normalize`` = GroupNorm(32,eps=1e-6) unet.py:71-74
* ``QKVAttentionLegacy.forward``    unet.py:107-125
* ``AttentionBlock._forward``       unet.py:168-174
* ``Downsample`` / ``Upsample``     unet.py:177-224
* ``ResBlock._forward``             unet.py:307-327
* ``UNetModel.__init__`` (topology) unet.py:372-505
* ``UNetModel.forward``             unet.py:512-563

It is written as a *plan* (list of layer records derived from the config) plus a
functional executor over a plain ``state_dict`` -- no nn.Module classes -- so it
can run on the GPU box where ``/root/reference`` does not exist.  The key grammar
of the state dict is the reference's and is verified by
loading our generated dict into the real ``UNetModel`` with ``strict=True``
(tests/test_oracle_unet.py).
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Dict, List

import torch
import torch.nn.functional as F

DEFAULT_CFG = dict(  # config/config_ldm.yaml:30-43 with in/out = latent_channels = 1
    image_size=768,
    in_channels=1,
    out_channels=1,
    model_channels=128,
    attention_resolutions=[8, 4],
    num_res_blocks=2,
    channel_mult=[1, 2, 4],
    dropout=0.0,
    conv_resample=True,
    num_heads=1,
    num_head_channels=-1,
    use_scale_shift_norm=False,
    resblock_updown=True,
)


def full_cfg(**over) -> dict:
    cfg = dict(DEFAULT_CFG)
    cfg.update(over)
    return cfg


# --------------------------------------------------------------------------- plan
def unet_plan(cfg: dict) -> dict:
    """Topology of UNetModel.__init__ (unet.py:382-505) as plain records.

    Each block is a list of layers; a layer is a dict with ``kind`` in
    {"conv_in", "res", "attn", "down_conv", "up_conv"} and a ``prefix`` that is
    the state_dict prefix of that module.
    """
    mc = cfg["model_channels"]
    mult = list(cfg["channel_mult"])
    nres = cfg["num_res_blocks"]
    att = set(cfg["attention_resolutions"])
    heads = cfg.get("num_heads", 1)
    nhc = cfg.get("num_head_channels", -1)
    updown = cfg.get("resblock_updown", False)
    conv_resample = cfg.get("conv_resample", True)

    def n_heads(ch):
        return heads if nhc == -1 else ch // nhc

    input_blocks: List[List[dict]] = [[dict(kind="conv_in", prefix="input_blocks.0.0",
                                            cin=cfg["in_channels"], cout=mc)]]
    chans = [mc]
    ch, ds = mc, 1
    for level, m in enumerate(mult):
        for _ in range(nres):
            idx = len(input_blocks)
            layers = [dict(kind="res", prefix=f"input_blocks.{idx}.0", cin=ch, cout=m * mc, mode="none")]
            ch = m * mc
            if ds in att:
                layers.append(dict(kind="attn", prefix=f"input_blocks.{idx}.1", ch=ch, heads=n_heads(ch)))
            input_blocks.append(layers)
            chans.append(ch)
        if level != len(mult) - 1:
            idx = len(input_blocks)
            if updown:
                input_blocks.append([dict(kind="res", prefix=f"input_blocks.{idx}.0", cin=ch, cout=ch, mode="down")])
            else:
                input_blocks.append([dict(kind="down_conv", prefix=f"input_blocks.{idx}.0", ch=ch,
                                          use_conv=conv_resample)])
            chans.append(ch)
            ds *= 2
    middle = [
        dict(kind="res", prefix="middle_block.0", cin=ch, cout=ch, mode="none"),
        dict(kind="attn", prefix="middle_block.1", ch=ch, heads=n_heads(ch)),
        dict(kind="res", prefix="middle_block.2", cin=ch, cout=ch, mode="none"),
    ]
    output_blocks: List[List[dict]] = []
    for level, m in list(enumerate(mult))[::-1]:
        for i in range(nres + 1):
            ich = chans.pop()
            idx = len(output_blocks)
            layers = [dict(kind="res", prefix=f"output_blocks.{idx}.0", cin=ch + ich, cout=mc * m, mode="none")]
            ch = mc * m
            if ds in att:
                layers.append(dict(kind="attn", prefix=f"output_blocks.{idx}.{len(layers)}", ch=ch,
                                   heads=n_heads(ch)))
            if level and i == nres:
                if updown:
                    layers.append(dict(kind="res", prefix=f"output_blocks.{idx}.{len(layers)}", cin=ch, cout=ch,
                                       mode="up"))
                else:
                    layers.append(dict(kind="up_conv", prefix=f"output_blocks.{idx}.{len(layers)}", ch=ch,
                                       use_conv=conv_resample))
                ds //= 2
            output_blocks.append(layers)
    return dict(input_blocks=input_blocks, middle=middle, output_blocks=output_blocks, final_ch=ch,
                time_embed_dim=mc * 4)


def unet_param_shapes(cfg: dict) -> "OrderedDict[str, tuple]":
    """Every state_dict entry (name -> shape) in the reference's registration order."""
    plan = unet_plan(cfg)
    mc, ted = cfg["model_channels"], plan["time_embed_dim"]
    emb_mult = 2 if cfg.get("use_scale_shift_norm", False) else 1
    out: "OrderedDict[str, tuple]" = OrderedDict()

    def lin(p, i, o):
        out[p + ".weight"] = (o, i)
        out[p + ".bias"] = (o,)

    def conv(p, i, o, k):
        out[p + ".weight"] = (o, i, k)
        out[p + ".bias"] = (o,)

    def gn(p, c):
        out[p + ".weight"] = (c,)
        out[p + ".bias"] = (c,)

    def layer(l):
        p = l["prefix"]
        if l["kind"] == "conv_in":
            conv(p, l["cin"], l["cout"], 3)
        elif l["kind"] == "res":
            gn(p + ".in_layers.0", l["cin"])
            conv(p + ".in_layers.2", l["cin"], l["cout"], 3)
            lin(p + ".emb_layers.1", ted, emb_mult * l["cout"])
            gn(p + ".out_layers.0", l["cout"])
            conv(p + ".out_layers.3", l["cout"], l["cout"], 3)
            if l["cin"] != l["cout"]:
                conv(p + ".skip_connection", l["cin"], l["cout"], 1)
        elif l["kind"] == "attn":
            gn(p + ".norm", l["ch"])
            conv(p + ".qkv", l["ch"], 3 * l["ch"], 1)
            conv(p + ".proj_out", l["ch"], l["ch"], 1)
        elif l["kind"] == "down_conv":
            if l["use_conv"]:
                conv(p + ".op", l["ch"], l["ch"], 3)
        elif l["kind"] == "up_conv":
            if l["use_conv"]:
                conv(p + ".conv", l["ch"], l["ch"], 3)

    lin("time_embed.0", mc, ted)
    lin("time_embed.2", ted, ted)
    for blk in plan["input_blocks"]:
        for l in blk:
            layer(l)
    for l in plan["middle"]:
        layer(l)
    for blk in plan["output_blocks"]:
        for l in blk:
            layer(l)
    gn("out.0", plan["final_ch"])
    conv("out.2", mc, cfg["out_channels"], 3)
    return out


def make_unet_state_dict(cfg: dict, seed: int = 0) -> "OrderedDict[str, torch.Tensor]":
    """Seeded synthetic weights (no trained checkpoint ships with the reference).

    Every tensor is random, *including* the 28 tensors that ``zero_module``
    (unet.py:39-45) zeroes at construction -- otherwise the network output is
    identically 0 and parity passes vacuously.  Scales are
    fan-in based so activations stay O(1) through the 21 residual blocks.
    """
    g = torch.Generator().manual_seed(seed)
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    for name, shape in unet_param_shapes(cfg).items():
        is_norm = (".in_layers.0." in name or ".out_layers.0." in name or ".norm." in name
                   or name.startswith("out.0."))
        if is_norm:
            if name.endswith("weight"):
                t = 1.0 + 0.1 * torch.randn(shape, generator=g)
            else:
                t = 0.1 * torch.randn(shape, generator=g)
        elif name.endswith("bias"):
            t = 0.05 * torch.randn(shape, generator=g)
        else:
            fan_in = 1
            for s in shape[1:]:
                fan_in *= s
            gain = 1.0
            if ".out_layers.3." in name or ".proj_out." in name:
                gain = 0.5  # residual branches: keep the stream variance bounded
            t = torch.randn(shape, generator=g) * (gain / math.sqrt(fan_in))
        sd[name] = t.float().contiguous()
    return sd


# ------------------------------------------------------------------- functional ops
def timestep_embedding(timesteps: torch.Tensor, dim: int, max_period: int = 10000) -> torch.Tensor:
    """unet.py:12-36 (cos half first, then sin half; odd dim zero-padded)."""
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(0, half, dtype=torch.float32) / half)
    args = timesteps[:, None].float() * freqs[None]
    emb = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)
    if dim % 2:
        emb = torch.cat([emb, torch.zeros_like(emb[:, :1])], dim=-1)
    return emb


def _gn(x, sd, p):
    return F.group_norm(x, 32, sd[p + ".weight"], sd[p + ".bias"], eps=1e-6)


def _conv(x, sd, p, padding, stride=1):
    return F.conv1d(x, sd[p + ".weight"], sd[p + ".bias"], stride=stride, padding=padding)


def _resample(x, mode):
    if mode == "down":  # Downsample(use_conv=False) -> AvgPool1d(2,2)   unet.py:195
        return F.avg_pool1d(x, kernel_size=2, stride=2)
    if mode == "up":    # Upsample(use_conv=False) -> nearest x2         unet.py:221
        return F.interpolate(x, scale_factor=2, mode="nearest")
    return x


def res_block(x, emb, sd, l, use_scale_shift_norm=False):
    """ResBlock._forward, unet.py:307-327."""
    p = l["prefix"]
    h = F.silu(_gn(x, sd, p + ".in_layers.0"))
    if l["mode"] != "none":       # up/down applied to BOTH branches between SiLU and conv
        h = _resample(h, l["mode"])
        x = _resample(x, l["mode"])
    h = _conv(h, sd, p + ".in_layers.2", 1)
    emb_out = F.linear(F.silu(emb), sd[p + ".emb_layers.1.weight"], sd[p + ".emb_layers.1.bias"])[..., None]
    if use_scale_shift_norm:
        scale, shift = torch.chunk(emb_out, 2, dim=1)
        h = _gn(h, sd, p + ".out_layers.0") * (1 + scale) + shift
        h = _conv(F.silu(h), sd, p + ".out_layers.3", 1)
    else:
        h = h + emb_out
        h = _conv(F.silu(_gn(h, sd, p + ".out_layers.0")), sd, p + ".out_layers.3", 1)
    if l["cin"] != l["cout"]:
        x = _conv(x, sd, p + ".skip_connection", 0)
    return x + h


def attention_block(x, sd, l):
    """AttentionBlock._forward + QKVAttentionLegacy.forward, unet.py:107-125,168-174."""
    p, nh = l["prefix"], l["heads"]
    b, c, t = x.shape
    qkv = _conv(_gn(x, sd, p + ".norm"), sd, p + ".qkv", 0)
    ch = c // nh
    q, k, v = qkv.reshape(b * nh, 3 * ch, t).split(ch, dim=1)
    scale = 1.0 / math.sqrt(math.sqrt(ch))
    w = torch.einsum("bct,bcs->bts", q * scale, k * scale)
    w = torch.softmax(w.float(), dim=-1)
    a = torch.einsum("bts,bcs->bct", w, v).reshape(b, c, t)
    return x + _conv(a, sd, p + ".proj_out", 0)


def _run_layer(h, emb, sd, l, cfg):
    k = l["kind"]
    if k == "conv_in":
        return _conv(h, sd, l["prefix"], 1)
    if k == "res":
        return res_block(h, emb, sd, l, cfg.get("use_scale_shift_norm", False))
    if k == "attn":
        return attention_block(h, sd, l)
    if k == "down_conv":
        if l["use_conv"]:
            return _conv(h, sd, l["prefix"] + ".op", 1, stride=2)
        return F.avg_pool1d(h, 2, 2)
    if k == "up_conv":
        h = F.interpolate(h, scale_factor=2, mode="nearest")
        if l["use_conv"]:
            h = _conv(h, sd, l["prefix"] + ".conv", 1)
        return h
    raise ValueError(k)


def unet_forward_train(cfg: dict, sd: Dict[str, torch.Tensor], x: torch.Tensor, timesteps: torch.Tensor,
                       taps: dict | None = None) -> torch.Tensor:
    """UNetModel.forward, unet.py:512-563, with autograd enabled (dropout = 0 in every reference config, so training and eval
    mode compute the same function).  ``timesteps`` has shape [1] or [B]."""
    plan = unet_plan(cfg)
    t_emb = timestep_embedding(timesteps, cfg["model_channels"])
    emb = F.linear(t_emb, sd["time_embed.0.weight"], sd["time_embed.0.bias"])
    emb = F.linear(F.silu(emb), sd["time_embed.2.weight"], sd["time_embed.2.bias"])
    hs = []
    h = x
    for bi, blk in enumerate(plan["input_blocks"]):
        for l in blk:
            h = _run_layer(h, emb, sd, l, cfg)
        hs.append(h)
        if taps is not None:
            taps[f"input_blocks.{bi}"] = h
    for l in plan["middle"]:
        h = _run_layer(h, emb, sd, l, cfg)
    if taps is not None:
        taps["middle_block"] = h
    for bi, blk in enumerate(plan["output_blocks"]):
        h = torch.cat([h, hs.pop()], dim=1)
        for l in blk:
            h = _run_layer(h, emb, sd, l, cfg)
        if taps is not None:
            taps[f"output_blocks.{bi}"] = h
    h = F.silu(_gn(h, sd, "out.0"))
    return _conv(h, sd, "out.2", 1)


@torch.no_grad()
def unet_forward(cfg: dict, sd: Dict[str, torch.Tensor], x: torch.Tensor, timesteps: torch.Tensor,
                 taps: dict | None = None) -> torch.Tensor:
    """UNetModel.forward, unet.py:512-563 (inference).  ``timesteps`` has shape [1] or [B]."""
    return unet_forward_train(cfg, sd, x, timesteps, taps)
