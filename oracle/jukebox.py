"""CPU restatement of the spectral (Jukebox) loss (TEST INFRASTRUCTURE).

Upstream: ``monai-generative`` ``generative/losses/spectral_loss.py`` (JukeboxLoss) --
not installable here -> **parity unpinned**; pinned by Parseval / known-answer
identities (tests/test_oracle_misc.py).  Call sites:
``src/train_autoencoderkl.py:158,208`` -- ``JukeboxLoss(spatial_dims=1, reduction="sum")``.

amplitude(x) = |fftn(x, dim=(1,2), norm="ortho")| ; loss = reduce((A(target)-A(input))^2).
"""
from __future__ import annotations

import torch


def fft_amplitude(x: torch.Tensor, spatial_dims: int = 1) -> torch.Tensor:
    dims = tuple(range(1, spatial_dims + 2))
    f = torch.fft.fftn(x, dim=dims, norm="ortho")
    return torch.sqrt(torch.real(f) ** 2 + torch.imag(f) ** 2)


def jukebox_loss(inp: torch.Tensor, target: torch.Tensor, spatial_dims: int = 1,
                 reduction: str = "sum") -> torch.Tensor:
    d = (fft_amplitude(target, spatial_dims) - fft_amplitude(inp, spatial_dims)) ** 2
    if reduction == "sum":
        return d.sum()
    if reduction == "mean":
        return d.mean()
    return d
