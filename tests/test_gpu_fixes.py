"""Regression tests for the round-1 review items: step-count changes on one handle (cached denoise-step graph vs the
per-step tables), device-resident timesteps, the f16x3 operand-range guard, the spectral loss on a sliced input."""
import pytest
import torch

from oracle import aekl as oa
from oracle import jukebox as oj
from oracle import sample as osamp
from oracle import unet as ou
from oracle.sample import SAMPLER_DEFAULTS

pytestmark = pytest.mark.gpu
RTOL, ATOL = 1e-3, 1e-4
# tensor-pipe eligible (channels % 128 == 0, attention at T = 128 with 256 channels) and small enough for the CPU oracle
SMALL = dict(model_channels=128, channel_mult=[1, 2], attention_resolutions=[2], num_res_blocks=1)


def _unet(cfg, sd, dev, math):
    import eegldm
    m = eegldm.UNetModel(**cfg, math=math)
    m.load_state_dict(sd)
    return m.to(dev).eval()


@pytest.mark.parametrize("math", ["fp32", "f16x3"])
def test_step_count_changes_on_one_handle(built_lib, cuda_device, math):
    """10, then 50, then 200 (the reference's own count, sample_trials.py:144), then 10 steps on the SAME model and (B, T):
    the cached step graph must follow the re-built timestep-embedding / coefficient tables (they used to be freed under it)."""
    import eegldm
    cfg = ou.full_cfg(**SMALL)
    sd = ou.make_unet_state_dict(cfg, 0)
    unet = _unet(cfg, sd, cuda_device, math)
    noise = torch.randn(2, 1, 256, generator=torch.Generator().manual_seed(0))
    for n in (10, 50, 200, 10):
        sched = eegldm.DDIMScheduler(**SAMPLER_DEFAULTS)
        sched.set_timesteps(n)
        y = eegldm.ddim_sample(unet, sched, noise.to(cuda_device)).cpu()   # step count taken from the scheduler
        ref = osamp.ddim_sample(cfg, sd, noise, n)
        torch.testing.assert_close(y, ref, rtol=RTOL, atol=ATOL)


def test_step_count_mismatch_is_refused(built_lib, cuda_device):
    import eegldm
    cfg = ou.full_cfg(**SMALL)
    unet = _unet(cfg, ou.make_unet_state_dict(cfg, 0), cuda_device, "fp32")
    sched = eegldm.DDIMScheduler(**SAMPLER_DEFAULTS)
    sched.set_timesteps(200)
    with pytest.raises(ValueError):
        eegldm.ddim_sample(unet, sched, torch.zeros(1, 1, 256, device=cuda_device), 50)
    with pytest.raises(ValueError):   # caller-supplied host output of the wrong shape
        eegldm.ddim_sample_host(unet, sched, torch.zeros(1, 1, 256), out_host=torch.zeros(1, 1, 128), device=cuda_device)


@pytest.mark.parametrize("math", ["fp32", "f16x3"])
def test_device_timesteps_equal_host_timesteps(built_lib, cuda_device, math):
    """timesteps as a CUDA tensor (training.py:430) take the on-device embedding kernel: same result as the host path,
    int64 and float, shared and per-sample."""
    cfg = ou.full_cfg(**SMALL)
    sd = ou.make_unet_state_dict(cfg, 0)
    m = _unet(cfg, sd, cuda_device, math)
    x = torch.randn(4, 1, 256, generator=torch.Generator().manual_seed(0)).to(cuda_device)
    for t in (torch.tensor([500]), torch.randint(0, 1000, (4,), generator=torch.Generator().manual_seed(1)),
              torch.tensor([0.0, 13.5, 999.0, 250.25])):
        y_host = m(x, timesteps=t)
        y_dev = m(x, timesteps=t.to(cuda_device))
        torch.testing.assert_close(y_dev, y_host, rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(y_dev.cpu(), ou.unet_forward(cfg, sd, x.cpu(), t), rtol=RTOL, atol=ATOL)


def test_f16x3_range_guard(built_lib, cuda_device):
    """A weight set whose residual stream exceeds the fp16 range (the raw stream feeds the 1x1 skip_connection un-normalised):
    fp32 stays finite and correct; f16x3 must REPORT it (flag / error), never return silent inf."""
    import eegldm
    cfg = ou.full_cfg(**SMALL)
    sd = ou.make_unet_state_dict(cfg, 0)
    sd = {k: v.clone() for k, v in sd.items()}
    sd["input_blocks.0.0.weight"] *= 3e6      # |h| ~ 1e6 on the stream that input_blocks.3's skip_connection (128 -> 256) reads raw
    x = torch.randn(2, 1, 256, generator=torch.Generator().manual_seed(0))
    t = torch.tensor([10])
    ref = ou.unet_forward(cfg, sd, x, t)
    assert torch.isfinite(ref).all()
    m32 = _unet(cfg, sd, cuda_device, "fp32")
    y32 = m32(x.to(cuda_device), timesteps=t)
    assert not m32.range_overflow()
    torch.testing.assert_close(y32.cpu(), ref, rtol=5e-3, atol=1e-3 * float(ref.abs().max()))
    m16 = _unet(cfg, sd, cuda_device, "f16x3")
    m16(x.to(cuda_device), timesteps=t)
    assert m16.range_overflow()               # raised ...
    assert not m16.range_overflow()           # ... and cleared by the read
    sched = eegldm.DDIMScheduler(**SAMPLER_DEFAULTS)
    sched.set_timesteps(2)
    with pytest.raises(eegldm.EegldmError):
        eegldm.ddim_sample(m16, sched, x.to(cuda_device))
    with pytest.raises(eegldm.EegldmError):
        eegldm.ddim_sample_host(m16, sched, x.contiguous(), device=cuda_device)
    # in-range weights: the flag stays down
    ok = _unet(cfg, ou.make_unet_state_dict(cfg, 0), cuda_device, "f16x3")
    ok(x.to(cuda_device), timesteps=t)
    assert not ok.range_overflow()
    # a WEIGHT outside the range keeps its layer on the fp32 SIMT kernel: still correct
    sd2 = {k: v.clone() for k, v in ou.make_unet_state_dict(cfg, 0).items()}
    sd2["input_blocks.1.0.in_layers.2.weight"][0, 0, 0] = 7e4
    y = _unet(cfg, sd2, cuda_device, "f16x3")(x.to(cuda_device), timesteps=t).cpu()
    torch.testing.assert_close(y, ou.unet_forward(cfg, sd2, x, t), rtol=RTOL, atol=ATOL)


def test_jukebox_gradient_through_a_sliced_input(built_lib, cuda_device):
    """recon[:, :, 36:-36] is non-contiguous: the gradient must flow through the slice (it used to come back as zero)."""
    import eegldm
    g = torch.Generator().manual_seed(0)
    x = torch.rand(3, 1, 1072, generator=g)
    y = torch.rand(3, 1, 1000, generator=g)
    xr = x.clone().requires_grad_(True)
    ref = oj.jukebox_loss(xr[:, :, 36:-36], y, reduction="sum")
    ref.backward()
    xd = x.to(cuda_device).requires_grad_(True)
    loss = eegldm.JukeboxLoss(spatial_dims=1, reduction="sum")(xd[:, :, 36:-36], y.to(cuda_device))
    loss.backward()
    torch.testing.assert_close(loss.cpu(), ref.detach(), rtol=1e-4, atol=1e-6)
    assert xd.grad is not None and float(xd.grad.abs().max()) > 0
    scale = float(xr.grad.abs().max())
    torch.testing.assert_close(xd.grad.cpu(), xr.grad, rtol=2e-3, atol=1e-4 * scale)
    # fp64 input: the cast happens under autograd
    x64 = x.double().to(cuda_device).requires_grad_(True)
    eegldm.JukeboxLoss(spatial_dims=1, reduction="sum")(x64[:, :, 36:-36], y.to(cuda_device)).backward()
    torch.testing.assert_close(x64.grad.float().cpu(), xr.grad, rtol=2e-3, atol=1e-4 * scale)
