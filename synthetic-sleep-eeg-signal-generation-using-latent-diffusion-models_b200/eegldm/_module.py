"""nn.Module plumbing shared by the wrappers: a parameter tree whose ``state_dict`` keys follow the
reference's grammar and lazy upload of the weights to the engine."""
from __future__ import annotations

import math

import torch
from torch import nn


class _Node(nn.Module):
    """Name-only container so that dotted reference keys (``input_blocks.3.0.in_layers.2.weight``)
    come out of ``state_dict()`` unchanged."""


def register_tree(root: nn.Module, infos, init_fn) -> None:
    for name, shape in infos:
        parts = name.split(".")
        mod = root
        for p in parts[:-1]:
            if p not in mod._modules:
                mod.add_module(p, _Node())
            mod = mod._modules[p]
        mod.register_parameter(parts[-1], nn.Parameter(init_fn(name, shape)))


def default_init(name: str, shape, weight_shape=None, zero=False) -> torch.Tensor:
    """PyTorch's default Conv1d / Linear initialisation statistics (kaiming_uniform(a=sqrt(5)) ==
    U(-1/sqrt(fan_in), 1/sqrt(fan_in)) for weight and bias); norm layers 1 / 0."""
    if zero:
        return torch.zeros(shape)
    ws = weight_shape if weight_shape is not None else shape
    if len(ws) == 1:  # GroupNorm affine
        return torch.ones(shape) if name.endswith("weight") else torch.zeros(shape)
    fan_in = 1
    for s in ws[1:]:
        fan_in *= s
    bound = 1.0 / math.sqrt(fan_in)
    return (torch.rand(shape) * 2 - 1) * bound


class EngineModule(nn.Module):
    """Tracks whether the engine's copy of the weights is stale."""

    def __init__(self):
        super().__init__()
        self._uploaded_key = None

    def _weights_key(self):
        return tuple((p.data_ptr(), p._version) for p in self.parameters())

    def _sync_weights(self) -> None:
        key = self._weights_key()
        if key != self._uploaded_key:
            self._upload(self.state_dict())
            self._uploaded_key = key

    def _upload(self, state_dict) -> None:  # pragma: no cover - overridden
        raise NotImplementedError

    def load_state_dict(self, state_dict, strict: bool = True, **kw):
        # DataParallel checkpoints carry a "module." prefix (testing/MSSIM_reconstruction.py:66-69)
        if any(k.startswith("module.") for k in state_dict):
            state_dict = {(k[7:] if k.startswith("module.") else k): v for k, v in state_dict.items()}
        out = super().load_state_dict(state_dict, strict=strict, **kw)
        self._uploaded_key = None
        return out


def check_cuda_f32(t: torch.Tensor, what: str) -> torch.Tensor:
    if not t.is_cuda:
        raise RuntimeError(f"eegldm: {what} must be a CUDA tensor (there is no CPU path)")
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()
