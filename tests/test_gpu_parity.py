"""Parity of the CUDA path (through the C ABI) against the CPU oracle and the committed golden vectors.
Tolerance is BASELINE.json's: rtol 1e-3 / atol 1e-4, fp32."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from golden.cases import CASES
from oracle import aekl as oa
from oracle import sample as osamp
from oracle import unet as ou
from oracle.sample import SAMPLER_DEFAULTS

pytestmark = pytest.mark.gpu
RTOL, ATOL = 1e-3, 1e-4
MATH = os.environ.get("EEGLDM_TEST_MATH", "fp32,f16x3").split(",")


def _unet(cfg, sd, dev, math="fp32"):
    import eegldm
    try:
        m = eegldm.UNetModel(**cfg, math=math)
    except eegldm.EegldmError:
        pytest.skip(f"math mode {math} not built")
    m.load_state_dict(sd)
    return m.to(dev).eval()


def _aekl(cfg, sd, dev):
    import eegldm
    m = eegldm.AutoencoderKL(**cfg)
    m.load_state_dict(sd)
    return m.to(dev).eval()


@pytest.mark.parametrize("math", MATH)
@pytest.mark.parametrize("name", list(CASES))
def test_unet_forward_matches_reference_golden(built_lib, cuda_device, name, math):
    over, B, T, ts = CASES[name]
    cfg = ou.full_cfg(**over)
    sd = ou.make_unet_state_dict(cfg, seed=0)
    g = np.load(os.path.join(GOLDEN, "unet_golden.npz"))
    x = torch.from_numpy(g[name + "/x"]).to(cuda_device)
    t = torch.from_numpy(g[name + "/t"])
    y = _unet(cfg, sd, cuda_device, math)(x, timesteps=t).cpu()
    torch.testing.assert_close(y, torch.from_numpy(g[name + "/y"]), rtol=RTOL, atol=ATOL)


# (CTA-pair mask, fuse_epilogues mask) of eegldm_set_conv_tuning: every switch off in turn next to the default (1, 15) -- pairs, operand-image
# qkv epilogue (2), fused producer (4), attention -> proj_out image (8), wide GroupNorm records (256), polyphase up-conv (512),
# per-segment records everywhere (1024), two-warpgroup conv epilogue (64), fused producer for the qkv conv (128), pairs on 128-wide tiles
TUNINGS = [(1, 15), (0, 15), (3, 15), (1, 13), (1, 11), (1, 7), (1, 15 + 256), (1, 15 + 512), (1, 15 + 1024), (1, 15 + 64), (1, 15 + 128)]


@pytest.mark.parametrize("tuning", TUNINGS)
def test_unet_tuning_switches_match_reference_golden(built_lib, cuda_device, tuning):
    """The full config_ldm.yaml UNet against the reference's own output under every tuning switch: the switches choose kernels and
    data paths (CTA pairs, epilogue forms, record widths, polyphase up-conv), never results beyond the parity tolerance."""
    from eegldm import _lib
    over, B, T, ts = CASES["ldm_per_sample_t"]
    cfg = ou.full_cfg(**over)
    sd = ou.make_unet_state_dict(cfg, seed=0)
    g = np.load(os.path.join(GOLDEN, "unet_golden.npz"))
    x = torch.from_numpy(g["ldm_per_sample_t/x"]).to(cuda_device)
    t = torch.from_numpy(g["ldm_per_sample_t/t"])
    _lib.check(built_lib.eegldm_set_conv_tuning(tuning[0], 1, tuning[1]))
    try:
        y = _unet(cfg, sd, cuda_device, "f16x3")(x, timesteps=t).cpu()
    finally:
        _lib.check(built_lib.eegldm_set_conv_tuning(1, 1, 15))
    torch.testing.assert_close(y, torch.from_numpy(g["ldm_per_sample_t/y"]), rtol=RTOL, atol=ATOL)


@pytest.mark.parametrize("math", MATH)
def test_unet_forward_matches_oracle_config2_slice(built_lib, cuda_device, math):
    """Config 2 (one denoise step, config_ldm.yaml) on a B=8 slice, per-sample and shared timesteps."""
    cfg = ou.full_cfg()
    sd = ou.make_unet_state_dict(cfg, seed=0)
    m = _unet(cfg, sd, cuda_device, math)
    x = torch.randn(8, 1, 768, generator=torch.Generator().manual_seed(0))
    t = torch.randint(0, 1000, (8,), generator=torch.Generator().manual_seed(1))
    for tt in (t, torch.tensor([500])):
        ref = ou.unet_forward(cfg, sd, x, tt)
        y = m(x.to(cuda_device), timesteps=tt).cpu()
        torch.testing.assert_close(y, ref, rtol=RTOL, atol=ATOL)
    # float timesteps (DiffusionInferer passes floats; unet.py:28 .float())
    ref = ou.unet_forward(cfg, sd, x[:2], torch.tensor([17.0]))
    torch.testing.assert_close(m(x[:2].to(cuda_device), timesteps=torch.tensor([17.0])).cpu(), ref, rtol=RTOL, atol=ATOL)


def test_unet_edge_shapes(built_lib, cuda_device):
    cfg = ou.full_cfg(model_channels=32, channel_mult=[1, 2], attention_resolutions=[2], image_size=32)
    sd = ou.make_unet_state_dict(cfg, seed=0)
    m = _unet(cfg, sd, cuda_device)
    # empty batch
    y = m(torch.zeros(0, 1, 32, device=cuda_device), timesteps=torch.tensor([3]))
    assert y.shape == (0, 1, 32)
    # ragged lengths: any even T runs (not only image_size); odd T is rejected like the reference's skip mismatch
    for T in (8, 30, 130, 258):   # (T=2 leaves 1-2 elements per GroupNorm group: rstd ~ 1e3 amplifies round-off)
        x = torch.randn(3, 1, T, generator=torch.Generator().manual_seed(T))
        ref = ou.unet_forward(cfg, sd, x, torch.tensor([9]))
        torch.testing.assert_close(m(x.to(cuda_device), timesteps=torch.tensor([9])).cpu(), ref, rtol=RTOL, atol=ATOL)
    import eegldm
    with pytest.raises(eegldm.EegldmError):
        m(torch.zeros(1, 1, 31, device=cuda_device), timesteps=torch.tensor([3]))
    with pytest.raises(ValueError):
        m(torch.zeros(2, 1, 32, device=cuda_device), timesteps=torch.tensor([3, 4, 5]))


@pytest.mark.parametrize("nc,z", [([2, 2, 4], 1), ([32, 32, 64], 1), ([32, 32, 64], 3), ([8, 16], 2)])
def test_aekl_encode_decode_matches_oracle(built_lib, cuda_device, nc, z):
    """Config 1: x = rand(4,1,3072) in [0,1] with the 36-sample constant pads (dataset.py:15-18)."""
    cfg = oa.full_cfg(num_channels=nc, latent_channels=z, attention_levels=[False] * len(nc))
    sd = oa.make_aekl_state_dict(cfg, seed=42)
    m = _aekl(cfg, sd, cuda_device)
    x = torch.rand(4, 1, 3072, generator=torch.Generator().manual_seed(0))
    x[..., :36] = 0
    x[..., -36:] = 0
    mu, sigma = oa.encode(cfg, sd, x)
    gmu, gsigma = m.encode(x.to(cuda_device))
    torch.testing.assert_close(gmu.cpu(), mu, rtol=RTOL, atol=ATOL)
    torch.testing.assert_close(gsigma.cpu(), sigma, rtol=RTOL, atol=ATOL)
    eps = torch.randn(mu.shape, generator=torch.Generator().manual_seed(1))
    zlat = oa.sampling(mu, sigma, eps)
    torch.testing.assert_close(m.decode(zlat.to(cuda_device)).cpu(), oa.decode(cfg, sd, zlat), rtol=RTOL, atol=ATOL)
    torch.testing.assert_close(m.reconstruct(x.to(cuda_device)).cpu(), oa.decode(cfg, sd, mu), rtol=RTOL, atol=2 * ATOL)
    torch.testing.assert_close(m.decode_stage_2_outputs(zlat.to(cuda_device)).cpu(), oa.decode(cfg, sd, zlat), rtol=RTOL, atol=ATOL)
    rec, fmu, fsig = m(x.to(cuda_device))
    assert rec.shape == x.shape and fmu.shape == mu.shape and fsig.shape == sigma.shape
    assert m.encode_stage_2_inputs(x.to(cuda_device)).shape == mu.shape


@pytest.mark.parametrize("name", ["c32_112_z1", "c32_12_z3"])
def test_aekl_matches_reference_golden(built_lib, cuda_device, name):
    """CUDA autoencoder against vectors produced by the reference's in-tree class (tests/golden/make_golden_aekl.py)."""
    from test_oracle_aekl import golden_case
    cfg, sd, t = golden_case(name)
    m = _aekl(cfg, sd, cuda_device)
    gmu, gsigma = m.encode(t["x"].to(cuda_device))
    torch.testing.assert_close(gmu.cpu(), t["mu"], rtol=RTOL, atol=ATOL)
    torch.testing.assert_close(gsigma.cpu(), t["sigma"], rtol=RTOL, atol=ATOL)
    z = t["mu"] + t["eps"] * t["sigma"]
    torch.testing.assert_close(m.decode(z.to(cuda_device)).cpu(), t["recon"], rtol=RTOL, atol=ATOL)


def test_aekl_forward_abi_with_supplied_eps(built_lib, cuda_device):
    import ctypes as C
    from eegldm import _lib
    cfg = oa.full_cfg()
    sd = oa.make_aekl_state_dict(cfg, seed=42)
    m = _aekl(cfg, sd, cuda_device)
    m._sync_weights()
    x = torch.rand(3, 1, 3072, generator=torch.Generator().manual_seed(0))
    eps = torch.randn(3, 1, 768, generator=torch.Generator().manual_seed(1))
    ref, rmu, rsig = oa.forward(cfg, sd, x, eps)
    xd, ed = x.to(cuda_device), eps.to(cuda_device)
    rec = torch.empty(3, 1, 3072, device=cuda_device)
    mu = torch.empty(3, 1, 768, device=cuda_device)
    sig = torch.empty_like(mu)
    p = lambda t: C.c_void_p(t.data_ptr())
    _lib.check(built_lib.eegldm_aekl_forward(m._h, p(xd), p(ed), p(rec), p(mu), p(sig), 3, 3072,
                                             C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    torch.cuda.synchronize()
    torch.testing.assert_close(rec.cpu(), ref, rtol=RTOL, atol=ATOL)
    torch.testing.assert_close(mu.cpu(), rmu, rtol=RTOL, atol=ATOL)
    torch.testing.assert_close(sig.cpu(), rsig, rtol=RTOL, atol=ATOL)


@pytest.mark.parametrize("math", MATH)
@pytest.mark.parametrize("pred,schedule,be", [("v_prediction", "scaled_linear_beta", 0.0205), ("epsilon", "linear_beta", 0.0195)])
def test_ddim50_matches_oracle_config3_slice(built_lib, cuda_device, pred, schedule, be, math):
    """Config 3 on a B=4 slice: DDIM-50 + AEKL 2-2-4 decode + crop to 3000, vs the CPU oracle."""
    import eegldm
    ucfg, acfg = ou.full_cfg(), oa.full_cfg()
    usd, asd = ou.make_unet_state_dict(ucfg, 0), oa.make_aekl_state_dict(acfg, 42)
    unet, aekl = _unet(ucfg, usd, cuda_device, math), _aekl(acfg, asd, cuda_device)
    kw = dict(SAMPLER_DEFAULTS, prediction_type=pred, schedule=schedule, beta_end=be)
    noise = torch.randn(4, 1, 768, generator=torch.Generator().manual_seed(0))
    ref = osamp.ddim_sample(ucfg, usd, noise, 50, acfg, asd, scale_factor=0.9, sched_kwargs=kw, crop=36)
    sched = eegldm.DDIMScheduler(**kw)
    sched.set_timesteps(50)
    y = eegldm.ddim_sample(unet, sched, noise.to(cuda_device), 50, aekl, scale_factor=0.9, crop=36)
    assert y.shape == (4, 1, 3000)
    torch.testing.assert_close(y.cpu(), ref, rtol=RTOL, atol=ATOL)
    # latent-only output, host-buffer entry point, and the un-fused drop-in loop all agree
    lat_ref = osamp.ddim_sample(ucfg, usd, noise, 50, sched_kwargs=kw)
    lat = eegldm.ddim_sample(unet, sched, noise.to(cuda_device), 50)
    torch.testing.assert_close(lat.cpu(), lat_ref, rtol=RTOL, atol=ATOL)
    yh = eegldm.ddim_sample_host(unet, sched, noise.pin_memory(), 50, aekl, scale_factor=0.9)
    torch.testing.assert_close(yh[:, :, 36:-36], y.cpu(), rtol=0, atol=0)
    x = noise.to(cuda_device)
    for t in sched.timesteps:                      # sample_trials.py:155-163 verbatim
        out = unet(x, timesteps=torch.tensor([int(t)]))
        x, _ = sched.step(out, int(t), x)
    torch.testing.assert_close(x.cpu(), lat_ref, rtol=RTOL, atol=ATOL)


def test_ddim_z3_and_graph_vs_eager(built_lib, cuda_device):
    import eegldm
    ucfg = ou.full_cfg(model_channels=32, channel_mult=[1, 2], attention_resolutions=[2], in_channels=3, out_channels=3,
                       image_size=64)
    acfg = oa.full_cfg(num_channels=[8, 16], latent_channels=3, attention_levels=[False, False])
    usd, asd = ou.make_unet_state_dict(ucfg, 0), oa.make_aekl_state_dict(acfg, 42)
    unet, aekl = _unet(ucfg, usd, cuda_device), _aekl(acfg, asd, cuda_device)
    noise = torch.randn(5, 3, 64, generator=torch.Generator().manual_seed(0))
    ref = osamp.ddim_sample(ucfg, usd, noise, 10, acfg, asd, scale_factor=1.3)
    sched = eegldm.DDIMScheduler(**SAMPLER_DEFAULTS)
    sched.set_timesteps(10)
    y = eegldm.ddim_sample(unet, sched, noise.to(cuda_device), 10, aekl, scale_factor=1.3)
    torch.testing.assert_close(y.cpu(), ref, rtol=RTOL, atol=ATOL)
    built_lib.eegldm_set_graphs(0)
    try:
        y2 = eegldm.ddim_sample(unet, sched, noise.to(cuda_device), 10, aekl, scale_factor=1.3)
    finally:
        built_lib.eegldm_set_graphs(1)
    assert torch.equal(y, y2)     # graph replay == eager launches, bit for bit
    # two-lane graph (batch halves on two captured streams, B = 5 -> 3 + 2): rows are independent, so bit-identical again
    built_lib.eegldm_set_sample_lanes(2)
    try:
        unet2 = _unet(ucfg, usd, cuda_device)     # fresh handle: graphs are cached per (B, T)
        y3 = eegldm.ddim_sample(unet2, sched, noise.to(cuda_device), 10, aekl, scale_factor=1.3)
    finally:
        built_lib.eegldm_set_sample_lanes(1)
    assert torch.equal(y, y3)


@pytest.mark.parametrize("B,T", [(5, 192), (3, 320), (1, 64)])
def test_tensor_pipe_vs_simt_odd_shapes(built_lib, cuda_device, B, T):
    """Full config_ldm.yaml UNet at shapes whose 128-row tiles straddle samples and end in partial tiles (T = 192: 12
    sixteen-position segments per sample at level 0, 3 at level 2; B odd): the tcgen05 path (fused producer, two-segment
    convs with virtual concat, nearest-x2 / AvgPool ResBlocks, attention where eligible) against the fp32 SIMT path of the
    same engine, per-sample timesteps."""
    ucfg = ou.full_cfg()
    usd = ou.make_unet_state_dict(ucfg, 0)
    x = torch.randn(B, 1, T, generator=torch.Generator().manual_seed(B * 1000 + T)).to(cuda_device)
    t = torch.randint(0, 1000, (B,), generator=torch.Generator().manual_seed(7))
    y32 = _unet(ucfg, usd, cuda_device, "fp32")(x, timesteps=t)
    y16 = _unet(ucfg, usd, cuda_device, "f16x3")(x, timesteps=t)
    assert torch.isfinite(y16).all()
    torch.testing.assert_close(y16, y32, rtol=RTOL, atol=ATOL)


def test_large_batch_rows_match_small_batch(built_lib, cuda_device):
    """Bench-scale batch through the fused sampling call (B = 3000: 2.3e9 activation elements at level 0, beyond 2^31): finite,
    and the last rows equal the same rows sampled as a small batch bit for bit (64-bit offsets, persistent tile walks with
    thousands of tiles per CTA)."""
    import eegldm
    ucfg, acfg = ou.full_cfg(), oa.full_cfg()
    unet = _unet(ucfg, ou.make_unet_state_dict(ucfg, 0), cuda_device, "f16x3")
    aekl = _aekl(acfg, oa.make_aekl_state_dict(acfg, 42), cuda_device)
    sched = eegldm.DDIMScheduler(**SAMPLER_DEFAULTS)
    sched.set_timesteps(2)
    B = 3000
    noise = torch.randn(B, 1, 768, generator=torch.Generator().manual_seed(3)).to(cuda_device)
    y = eegldm.ddim_sample(unet, sched, noise, 2, aekl)
    assert y.shape == (B, 1, 3072) and torch.isfinite(y).all()
    tail = eegldm.ddim_sample(unet, sched, noise[B - 77:], 2, aekl)
    assert torch.equal(y[B - 77:], tail)


@pytest.mark.parametrize("math", MATH)
def test_raw_signal_dm_variant(built_lib, cuda_device, math):
    """the raw-signal diffusion model (config_dm.yaml: the same UNet on [B,1,3072], self-attention at T = 768)
    sampled as sample_trials_ddpm.py:83-102 does -- DDIM steps on the signal itself, no autoencoder, crop [36:-36].
    T = 768 exceeds the tcgen05 attention's 256-key tile, so attention takes the fp32 SIMT kernel; convs take the math mode."""
    import eegldm
    ucfg = ou.full_cfg()
    usd = ou.make_unet_state_dict(ucfg, 0)
    unet = _unet(ucfg, usd, cuda_device, math)
    noise = torch.randn(2, 1, 3072, generator=torch.Generator().manual_seed(2))
    ref = osamp.ddim_sample(ucfg, usd, noise, 3)[:, :, 36:-36]
    sched = eegldm.DDIMScheduler(**SAMPLER_DEFAULTS)
    sched.set_timesteps(3)
    y = eegldm.ddim_sample(unet, sched, noise.to(cuda_device), 3, None)[:, :, 36:-36]
    assert y.shape == (2, 1, 3000)
    torch.testing.assert_close(y.cpu(), ref, rtol=RTOL, atol=ATOL)


@pytest.mark.parametrize("math", MATH)
def test_full_size_properties(built_lib, cuda_device, math):
    """Size-independent properties at bench batch sizes (the oracle is too slow here):
    windows are independent, so any batch split gives bit-identical rows; results are deterministic."""
    import eegldm
    ucfg, acfg = ou.full_cfg(), oa.full_cfg()
    unet = _unet(ucfg, ou.make_unet_state_dict(ucfg, 0), cuda_device, math)
    aekl = _aekl(acfg, oa.make_aekl_state_dict(acfg, 42), cuda_device)
    sched = eegldm.DDIMScheduler(**SAMPLER_DEFAULTS)
    sched.set_timesteps(3)
    B = 64
    noise = torch.randn(B, 1, 768, generator=torch.Generator().manual_seed(0)).to(cuda_device)
    full = eegldm.ddim_sample(unet, sched, noise, 3, aekl, crop=36)
    again = eegldm.ddim_sample(unet, sched, noise, 3, aekl, crop=36)
    assert torch.equal(full, again)
    parts = torch.cat([eegldm.ddim_sample(unet, sched, noise[lo:hi], 3, aekl, crop=36)
                       for lo, hi in ((0, 8), (8, 40), (40, 64))])
    assert torch.equal(full, parts)
    assert torch.isfinite(full).all()
    # one UNet forward at config 2's full batch (256): finite, and equal to the B=8 slices
    x = torch.randn(256, 1, 768, generator=torch.Generator().manual_seed(0)).to(cuda_device)
    t = torch.randint(0, 1000, (256,), generator=torch.Generator().manual_seed(1))
    y = unet(x, timesteps=t)
    assert torch.isfinite(y).all()
    assert torch.equal(y[8:16], unet(x[8:16], timesteps=t[8:16]))
