"""Self-checks of the restated [upstream] pieces (monai-generative is not installable here, so these
are known-answer / identity checks): DDIM/DDPM schedulers, AutoencoderKL, JukeboxLoss."""
import math

import numpy as np
import pytest
import torch

from oracle import aekl as oa
from oracle import jukebox as oj
from oracle.schedulers import DDIMScheduler, DDPMScheduler
from oracle.sample import SAMPLER_DEFAULTS


def test_ddim_known_answers():
    s = DDIMScheduler(**SAMPLER_DEFAULTS)
    s.set_timesteps(50)
    assert s.timesteps.tolist() == list(range(980, -1, -20))
    ac = s.alphas_cumprod
    for t, v in ((0, 0.99849999), (20, 0.96728837), (500, 0.10767679), (980, 0.00014290)):
        assert abs(float(ac[t]) - v) < 2e-7 * max(1, v / 1e-3), (t, float(ac[t]))
    for t, (c1, c2) in ((980, (0.9999966, -0.0026253)), (20, (0.9897751, -0.1426366)), (0, (0.9992497, -0.0387300))):
        a, b = s.step_coefficients(t)
        assert abs(a - c1) < 2e-6 and abs(b - c2) < 2e-6
    e = DDIMScheduler(**dict(SAMPLER_DEFAULTS, prediction_type="epsilon"))
    e.set_timesteps(50)
    for t, (c1, c2) in ((980, (1.2195930, -0.2196122)), (0, (1.0007509, -0.0387590))):
        a, b = e.step_coefficients(t)
        assert abs(a - c1) < 3e-6 and abs(b - c2) < 3e-6


@pytest.mark.parametrize("pred", ["v_prediction", "epsilon"])
def test_ddim_step_equals_two_scalar_form(pred):
    s = DDIMScheduler(**dict(SAMPLER_DEFAULTS, prediction_type=pred))
    s.set_timesteps(50)
    g = torch.Generator().manual_seed(0)
    x, m = torch.randn(2, 1, 64, generator=g), torch.randn(2, 1, 64, generator=g)
    for t in (980, 500, 20, 0):
        prev, x0 = s.step(m, t, x)
        c1, c2 = s.step_coefficients(t)
        torch.testing.assert_close(prev, c1 * x + c2 * m, rtol=1e-5, atol=1e-5)
    # last step lands on x0 (alpha_prev := 1)
    prev, x0 = s.step(m, 0, x)
    torch.testing.assert_close(prev, x0, rtol=1e-6, atol=1e-6)


def test_ddpm_add_noise_velocity_identities():
    s = DDPMScheduler(1000, 0.0015, 0.0195, "linear_beta")   # train_ldm.py:199-200
    g = torch.Generator().manual_seed(0)
    x0, eps = torch.randn(4, 1, 32, generator=g), torch.randn(4, 1, 32, generator=g)
    t = torch.tensor([0, 10, 500, 999])
    xt = s.add_noise(x0, eps, t)
    v = s.get_velocity(x0, eps, t)
    a = s.alphas_cumprod[t].sqrt()[:, None, None]
    b = (1 - s.alphas_cumprod[t]).sqrt()[:, None, None]
    torch.testing.assert_close(a * xt - b * v, x0, rtol=1e-4, atol=1e-5)   # x0 recovered from (x_t, v)
    torch.testing.assert_close(a * v + b * xt, eps, rtol=1e-4, atol=1e-5)


def test_jukebox_parseval_and_zero():
    g = torch.Generator().manual_seed(0)
    x = torch.rand(3, 1, 3072, generator=g)
    A = oj.fft_amplitude(x)
    torch.testing.assert_close((A ** 2).sum(), (x ** 2).sum(), rtol=1e-4, atol=1e-3)   # ortho norm
    assert float(oj.jukebox_loss(x, x)) == 0.0
    # amplitude spectrum is shift-invariant -> loss 0 for a circular shift
    assert float(oj.jukebox_loss(torch.roll(x, 7, dims=-1), x)) < 1e-6
    # a pure cosine: amplitude sqrt(N)/2 at bins k and N-k
    n, k = 3072, 5
    c = torch.cos(2 * math.pi * k * torch.arange(n) / n)[None, None]
    A = oj.fft_amplitude(c)[0, 0]
    assert abs(float(A[k]) - math.sqrt(n) / 2) < 1e-2 and abs(float(A[n - k]) - math.sqrt(n) / 2) < 1e-2


@pytest.mark.parametrize("nc,z", [([2, 2, 4], 1), ([32, 32, 64], 1), ([32, 32, 64], 3)])
def test_aekl_shapes_and_param_counts(nc, z):
    cfg = oa.full_cfg(num_channels=nc, latent_channels=z)
    sd = oa.make_aekl_state_dict(cfg)
    x = torch.rand(2, 1, 3072, generator=torch.Generator().manual_seed(0))
    mu, sigma = oa.encode(cfg, sd, x)
    assert mu.shape == (2, z, 768) and sigma.shape == (2, z, 768)          # 3072 -> 1536 -> 768
    assert float(sigma.min()) > 0
    y = oa.decode(cfg, sd, mu)
    assert y.shape == (2, 1, 3072)
    enc = sum(v.numel() for k, v in sd.items() if k.startswith(("encoder", "quant")))
    dec = sum(v.numel() for k, v in sd.items() if k.startswith(("decoder", "post_quant")))
    if nc == [2, 2, 4] and z == 1:
        assert (enc, dec) == (429, 505)
    # MONAI key grammar
    assert "encoder.blocks.0.conv.weight" in sd and "quant_conv_log_sigma.conv.bias" in sd
    assert "decoder.blocks.1.norm1.weight" in sd and "decoder.blocks.1.conv1.conv.weight" in sd


def test_aekl_kl_and_sampling():
    mu = torch.zeros(2, 1, 8)
    sigma = torch.ones(2, 1, 8)
    assert float(oa.kl_loss(mu, sigma)) == 0.0
    eps = torch.full((2, 1, 8), 2.0)
    torch.testing.assert_close(oa.sampling(mu + 1, sigma * 3, eps), torch.full((2, 1, 8), 7.0))


@pytest.mark.parametrize("name", ["scaled_linear", "linear"])
def test_noise_schedule_matches_reference_in_tree_golden(name):
    """betas / cumulative alphas / add_noise against the reference's in-tree ancestor of the MONAI schedulers
    (src/models/ldm.py make_beta_schedule + DDPM.q_sample; tests/golden/make_golden_sched.py).  The reference computes
    the tables in float64, upstream (and the oracle) in float32: agreement to fp32 round-off of a 1000-term product."""
    import os
    from conftest import GOLDEN
    from golden.make_golden_sched import SCHEDULES
    g = np.load(os.path.join(GOLDEN, "sched_golden.npz"))
    _, oname, b0, b1 = SCHEDULES[name]
    s = DDPMScheduler(num_train_timesteps=1000, beta_start=b0, beta_end=b1, schedule=oname)
    np.testing.assert_allclose(s.betas.numpy(), g[name + "/betas"], rtol=2e-6)
    np.testing.assert_allclose(s.alphas_cumprod.numpy(), g[name + "/alphas_cumprod"], rtol=2e-4)   # 1000 fp32 factors
    xt = s.add_noise(torch.from_numpy(g["x0"]), torch.from_numpy(g["noise"]), torch.from_numpy(g["t"]))
    torch.testing.assert_close(xt, torch.from_numpy(g[name + "/x_t"]), rtol=1e-4, atol=1e-5)
