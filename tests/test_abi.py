"""C-ABI checks that need no GPU: the library loads, exports every symbol include/eegldm.h declares,
its host-side logic (topology / state_dict grammar / scheduler tables / timestep embedding / error
codes) agrees with the oracle, and compute entry points fail loudly without a device."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import torch

from conftest import ROOT
from oracle import aekl as oa
from oracle import unet as ou
from oracle.sample import SAMPLER_DEFAULTS
from oracle.schedulers import DDIMScheduler as ODDIM


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "eegldm.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(eegldm_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(built_lib):
    from eegldm import _lib
    syms = _header_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(built_lib, s), f"{s} declared in include/eegldm.h but not exported"
    assert set(_lib.SIGNATURES) == set(syms), "ctypes binding and header disagree"
    assert built_lib.eegldm_version().startswith(b"eegldm")


def test_struct_sizes_match_header(built_lib):
    from eegldm import _lib
    assert C.sizeof(_lib.UNetCfg) == 4 * (5 + 1 + 8 + 1 + 8 + 6)
    assert C.sizeof(_lib.AeklCfg) == 4 * (3 + 8 + 8 + 2)
    assert C.sizeof(_lib.SchedCfg) == 4 * 7


UNET_CFGS = [dict(), dict(in_channels=3, out_channels=3),
             dict(model_channels=32, channel_mult=[1, 2], attention_resolutions=[1, 2], num_heads=4),
             dict(model_channels=32, channel_mult=[1, 2, 2], attention_resolutions=[4], resblock_updown=False),
             dict(model_channels=32, channel_mult=[1, 2], attention_resolutions=[], resblock_updown=False,
                  conv_resample=False, num_res_blocks=1)]


@pytest.mark.parametrize("over", UNET_CFGS)
def test_unet_param_grammar_matches_oracle(built_lib, over):
    import eegldm
    cfg = ou.full_cfg(**over)
    m = eegldm.UNetModel(**cfg)
    want = ou.unet_param_shapes(cfg)
    got = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    assert list(got.keys()) == list(want.keys())
    assert got == dict(want)
    # zero_module tensors start at zero like the reference's (unet.py:39-45)
    sd = m.state_dict()
    assert float(sd["out.2.weight"].abs().max()) == 0.0
    assert float(sd["input_blocks.1.0.out_layers.3.weight"].abs().max()) == 0.0
    assert float(sd["input_blocks.1.0.in_layers.2.weight"].abs().max()) > 0.0


@pytest.mark.parametrize("nc,z", [([2, 2, 4], 1), ([32, 32, 64], 3)])
def test_aekl_param_grammar_matches_oracle(built_lib, nc, z):
    import eegldm
    cfg = oa.full_cfg(num_channels=nc, latent_channels=z)
    m = eegldm.AutoencoderKL(**cfg)
    want = oa.aekl_param_shapes(cfg)
    got = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    assert set(got.keys()) == set(want.keys())
    assert got == dict(want)


def test_load_errors(built_lib):
    from eegldm import _lib
    import eegldm
    m = eegldm.UNetModel(**ou.full_cfg(model_channels=32, channel_mult=[1], attention_resolutions=[]))
    w = np.zeros((3, 3), dtype=np.float32)
    shape = (C.c_int64 * 2)(3, 3)
    assert built_lib.eegldm_unet_load(m._h, b"no.such.key", C.c_void_p(w.ctypes.data), shape, 2) == -3
    assert b"no.such.key" in built_lib.eegldm_last_error()
    assert built_lib.eegldm_unet_load(m._h, b"time_embed.0.weight", C.c_void_p(w.ctypes.data), shape, 2) == -2
    # forward before finalize -> EEGLDM_ERR_MISSING, never a crash
    one = (C.c_float * 1)(5.0)
    assert built_lib.eegldm_unet_forward(m._h, C.c_void_p(16), one, 1, C.c_void_p(16), 1, 32, None) == -3
    assert built_lib.eegldm_unet_forward(None, C.c_void_p(16), one, 1, C.c_void_p(16), 1, 32, None) == -1
    # DataParallel "module." prefix is accepted (testing/MSSIM_reconstruction.py:66-69)
    w = np.zeros((128, 32), dtype=np.float32)
    shape = (C.c_int64 * 2)(128, 32)
    assert built_lib.eegldm_unet_load(m._h, b"module.time_embed.0.weight", C.c_void_p(w.ctypes.data), shape, 2) == 0


def test_bad_configs_rejected(built_lib):
    import eegldm
    with pytest.raises(eegldm.EegldmError):
        eegldm.UNetModel(**ou.full_cfg(model_channels=48))            # GroupNorm32 needs C % 32 == 0
    with pytest.raises(NotImplementedError):
        eegldm.UNetModel(**ou.full_cfg(use_scale_shift_norm=True))
    with pytest.raises(NotImplementedError):
        eegldm.AutoencoderKL(**oa.full_cfg(with_encoder_nonlocal_attn=True))
    with pytest.raises(eegldm.EegldmError):
        eegldm.AutoencoderKL(**oa.full_cfg(num_channels=[3, 4, 4], norm_num_groups=2))


@pytest.mark.parametrize("schedule,pred,bs,be", [("scaled_linear_beta", "v_prediction", 0.0015, 0.0205),
                                                 ("scaled_linear_beta", "epsilon", 0.0015, 0.0205),
                                                 ("linear_beta", "epsilon", 0.0015, 0.0195)])
@pytest.mark.parametrize("n_steps", [50, 200, 1000, 7])
def test_scheduler_tables_match_oracle(built_lib, schedule, pred, bs, be, n_steps):
    import eegldm
    o = ODDIM(1000, bs, be, schedule, pred, clip_sample=False)
    o.set_timesteps(n_steps)
    s = eegldm.DDIMScheduler(1000, bs, be, schedule, pred, clip_sample=False)
    s.set_timesteps(n_steps)
    assert s.timesteps.tolist() == o.timesteps.tolist()
    np.testing.assert_allclose(s.alphas_cumprod.numpy(), o.alphas_cumprod.numpy(), rtol=2e-6, atol=0)
    for t in o.timesteps.tolist():
        c = s.step_coefficients(t)
        oc = o.step_coefficients(t)
        assert abs(c[0] - oc[0]) < 2e-6 * max(1, abs(oc[0])) and abs(c[1] - oc[1]) < 2e-6 * max(1, abs(oc[1]))
    # python-level step() agrees with the oracle's
    g = torch.Generator().manual_seed(0)
    x, m = torch.randn(2, 1, 16, generator=g), torch.randn(2, 1, 16, generator=g)
    t = int(o.timesteps[len(o.timesteps) // 2])
    torch.testing.assert_close(s.step(m, t, x)[0], o.step(m, t, x)[0], rtol=1e-5, atol=1e-5)


def test_scheduler_known_answers_via_abi(built_lib):
    import eegldm
    s = eegldm.DDIMScheduler(**SAMPLER_DEFAULTS)
    s.set_timesteps(50)
    assert s.timesteps.tolist() == list(range(980, -1, -20))
    for t, (c1, c2) in ((980, (0.9999966, -0.0026253)), (20, (0.9897751, -0.1426366)), (0, (0.9992497, -0.0387300))):
        a, b = s.step_coefficients(t)
        assert abs(a - c1) < 2e-6 and abs(b - c2) < 2e-6
    with pytest.raises(eegldm.EegldmError):
        s.set_timesteps(1001)


def test_timestep_embedding_abi(built_lib):
    ts = np.array([0, 1, 20, 500, 980, 999, 12.5], dtype=np.float32)
    out = np.empty((len(ts), 128), dtype=np.float32)
    assert built_lib.eegldm_timestep_embedding(ts.ctypes.data_as(C.POINTER(C.c_float)), len(ts), 128,
                                               out.ctypes.data_as(C.POINTER(C.c_float))) == 0
    ref = ou.timestep_embedding(torch.from_numpy(ts), 128).numpy()
    np.testing.assert_allclose(out, ref, rtol=0, atol=1e-4)   # fp32 range reduction at |arg| ~ 1e3
    np.testing.assert_allclose(out[4, 0:3], [0.98439258, 0.91511506, 0.97219217], atol=2e-6)
    np.testing.assert_allclose(out[4, 64:67], [-0.17598660, 0.40319273, -0.23418452], atol=2e-6)


def test_compute_fails_loudly_without_gpu(built_lib):
    """No CPU fallback: without a CUDA device finalize / forward return EEGLDM_ERR_CUDA."""
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import eegldm
    cfg = ou.full_cfg(model_channels=32, channel_mult=[1], attention_resolutions=[])
    m = eegldm.UNetModel(**cfg)
    with pytest.raises(eegldm.EegldmError) as ei:
        m._upload(ou.make_unet_state_dict(cfg))
    assert ei.value.code == -4
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 1, 32), timesteps=torch.tensor([1]))


def test_shard_range_partitions():
    from eegldm.sampler import shard_range
    for n in (0, 1, 7, 8, 1024, 8192, 8193):
        for w in (1, 2, 3, 8):
            r = [shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[i][1] == r[i + 1][0] for i in range(w - 1))
            assert max(h - l for l, h in r) - min(h - l for l, h in r) <= 1
