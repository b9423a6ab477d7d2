"""Config 5: the autoencoder training step (src/train_autoencoderkl.py:204-220, adversarial term excluded) and the spectral
loss, CUDA path through the C ABI vs the CPU oracle with torch autograd.  Tolerances: losses rtol 1e-4; gradients and
updated parameters rtol 2e-3 with an absolute floor of 1e-5 x the tensor's max magnitude (fp32 atomics reorder sums)."""
import numpy as np
import pytest
import torch

from oracle import aekl as oa
from oracle import jukebox as oj

pytestmark = pytest.mark.gpu


def _close(a, b, rtol, floor):
    scale = float(b.abs().max())
    torch.testing.assert_close(a, b, rtol=rtol, atol=floor * max(scale, 1e-30))


@pytest.mark.parametrize("reduction", ["sum", "mean"])
@pytest.mark.parametrize("B,N", [(3, 3072), (2, 256), (5, 1000)])
def test_jukebox_loss_and_gradient(built_lib, cuda_device, B, N, reduction):
    import eegldm
    g = torch.Generator().manual_seed(N)
    x = torch.rand(B, 1, N, generator=g)
    y = torch.rand(B, 1, N, generator=g)
    xr = x.clone().requires_grad_(True)
    ref = oj.jukebox_loss(xr, y, reduction=reduction)
    ref.backward()
    xd = x.to(cuda_device).requires_grad_(True)
    loss = eegldm.JukeboxLoss(spatial_dims=1, reduction=reduction)(xd, y.to(cuda_device))
    loss.backward()
    torch.testing.assert_close(loss.cpu(), ref.detach(), rtol=1e-4, atol=1e-6)
    _close(xd.grad.cpu(), xr.grad, 2e-3, 1e-4)
    # identical signals: zero loss
    z = eegldm.JukeboxLoss(spatial_dims=1, reduction=reduction)(xd.detach(), xd.detach())
    assert float(z) == 0.0


def _oracle_step(cfg, sd, x, eps, kl_w, spec_w, lr):
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    recon, mu, sigma = oa.forward(cfg, params, x, eps)
    l1 = torch.nn.functional.l1_loss(recon, x)
    kl = oa.kl_loss(mu, sigma)
    spec = oj.jukebox_loss(recon, x, reduction="sum")
    total = l1 + kl_w * kl + spec_w * spec
    total.backward()
    grads = {k: p.grad.clone() for k, p in params.items()}
    opt = torch.optim.Adam(list(params.values()), lr=lr)
    opt.step()
    return dict(l1=float(l1), kl=float(kl), spectral=float(spec), total=float(total)), grads, {k: p.detach() for k, p in params.items()}


@pytest.mark.parametrize("nc,B", [([2, 2, 4], 6), ([32, 32, 64], 2), ([8, 16], 3)])
def test_aekl_train_step_matches_autograd(built_lib, cuda_device, nc, B):
    import eegldm
    cfg = oa.full_cfg(num_channels=nc, attention_levels=[False] * len(nc))
    sd = oa.make_aekl_state_dict(cfg, seed=42)
    L = 3072 if len(nc) == 3 else 512
    x = torch.rand(B, 1, L, generator=torch.Generator().manual_seed(0))
    x[..., :36] = 0
    x[..., -36:] = 0
    T = L // (1 << (len(nc) - 1))
    eps = torch.randn(B, 1, T, generator=torch.Generator().manual_seed(1))
    kl_w, spec_w, lr = 1e-9, 1e4, 5e-3
    ref_losses, ref_grads, ref_params = _oracle_step(cfg, sd, x, eps, kl_w, spec_w, lr)
    m = eegldm.AutoencoderKL(**cfg)
    m.load_state_dict(sd)
    m = m.to(cuda_device)
    losses = m.train_step(x.to(cuda_device), eps.to(cuda_device), kl_weight=kl_w, spectral_weight=spec_w, lr=lr)
    for k in ("l1", "kl", "spectral", "total"):
        assert abs(losses[k] - ref_losses[k]) <= 1e-4 * abs(ref_losses[k]) + 1e-7, (k, losses[k], ref_losses[k])
    grads = m.grad_dict()
    assert set(grads) == set(ref_grads)
    for k in ref_grads:
        _close(grads[k], ref_grads[k], 2e-3, 1e-4)
    # parameters after one Adam step (lr 5e-3: every element moves by ~lr, sign errors would show)
    m.sync_trained()
    for k, p in m.state_dict().items():
        _close(p.cpu(), ref_params[k], 2e-3, 2e-4)
    # the inference path now uses the trained weights
    mu_ref, _ = oa.encode(cfg, ref_params, x)
    mu, _ = m.encode(x.to(cuda_device))
    torch.testing.assert_close(mu.cpu(), mu_ref, rtol=2e-3, atol=2e-4)


def test_aekl_train_loss_decreases(built_lib, cuda_device):
    """A few steps on a fixed batch reduce the objective (size-independent sanity at config 5's batch 512)."""
    import eegldm
    cfg = oa.full_cfg()
    m = eegldm.AutoencoderKL(**cfg)
    m.load_state_dict(oa.make_aekl_state_dict(cfg, seed=42))
    m = m.to(cuda_device)
    x = torch.rand(512, 1, 3072, generator=torch.Generator().manual_seed(0)).to(cuda_device)
    eps = torch.randn(512, 1, 768, generator=torch.Generator().manual_seed(1)).to(cuda_device)
    first = m.train_step(x, eps)["total"]
    for _ in range(20):
        last = m.train_step(x, eps)["total"]
    assert np.isfinite(last) and last < first
