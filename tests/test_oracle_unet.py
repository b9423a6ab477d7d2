"""The denoiser oracle (oracle/unet.py) is pinned against (a) the reference's own module, imported
from /root/reference when present, and (b) golden vectors that module produced
(tests/golden/make_golden.py)."""
import os
import sys

import numpy as np
import pytest
import torch

from conftest import GOLDEN, REFERENCE
from oracle import unet as ou
from golden.cases import CASES


def _golden():
    return np.load(os.path.join(GOLDEN, "unet_golden.npz"))


@pytest.mark.parametrize("name", list(CASES))
def test_oracle_matches_golden(name):
    over, B, T, ts = CASES[name]
    cfg = ou.full_cfg(**over)
    sd = ou.make_unet_state_dict(cfg, seed=0)
    g = _golden()
    x = torch.from_numpy(g[name + "/x"])
    t = torch.from_numpy(g[name + "/t"])
    y = ou.unet_forward(cfg, sd, x, t)
    ref = torch.from_numpy(g[name + "/y"])
    assert y.shape == ref.shape
    # same algorithm, same fp32 library ops: agreement to fp32 round-off, far inside rtol 1e-3 / atol 1e-4
    torch.testing.assert_close(y, ref, rtol=1e-5, atol=1e-5)
    assert float(ref.abs().mean()) > 0.05  # the 28 zero-initialised tensors were re-randomised (non-vacuous)


@pytest.mark.reference
def test_oracle_matches_reference_module_live():
    sys.path.insert(0, os.path.join(REFERENCE, "src"))
    from models.unet import UNetModel
    cfg = ou.full_cfg(model_channels=64, channel_mult=[1, 2, 4], attention_resolutions=[2, 4], num_heads=2, image_size=128)
    sd = ou.make_unet_state_dict(cfg, seed=3)
    m = UNetModel(**cfg).eval()
    m.load_state_dict(sd, strict=True)
    # key grammar + shapes + registration order
    ref_sd = m.state_dict()
    assert list(ref_sd.keys()) == list(sd.keys())
    for k in sd:
        assert tuple(ref_sd[k].shape) == tuple(sd[k].shape)
    x = torch.randn(3, 1, 128, generator=torch.Generator().manual_seed(5))
    for t in (torch.tensor([17]), torch.tensor([0, 499, 999]), torch.tensor([12.5])):
        with torch.no_grad():
            ref = m(x, timesteps=t)
        torch.testing.assert_close(ou.unet_forward(cfg, sd, x, t), ref, rtol=1e-5, atol=1e-5)


@pytest.mark.reference
def test_fresh_reference_model_outputs_zero():
    """SURVEY item 3: zero_module makes a freshly built reference UNet output exactly 0, hence the
    re-randomisation in make_unet_state_dict."""
    sys.path.insert(0, os.path.join(REFERENCE, "src"))
    from models.unet import UNetModel
    cfg = ou.full_cfg(model_channels=32, channel_mult=[1, 2], attention_resolutions=[2], image_size=32)
    m = UNetModel(**cfg).eval()
    with torch.no_grad():
        y = m(torch.randn(1, 1, 32), timesteps=torch.tensor([5]))
    assert float(y.abs().max()) == 0.0


def test_param_count_config_ldm():
    shapes = ou.unet_param_shapes(ou.full_cfg())
    n = sum(int(np.prod(s)) for s in shapes.values())
    assert n == 30_533_121 and len(shapes) == 278
    n3 = sum(int(np.prod(s)) for s in ou.unet_param_shapes(ou.full_cfg(in_channels=3, out_channels=3)).values())
    assert n3 == 30_534_659


def test_timestep_embedding_known_answers():
    e = ou.timestep_embedding(torch.tensor([980]), 128)[0]
    np.testing.assert_allclose(e[0:3].numpy(), [0.98439258, 0.91511506, 0.97219217], rtol=0, atol=2e-6)
    np.testing.assert_allclose(e[64:67].numpy(), [-0.17598660, 0.40319273, -0.23418452], rtol=0, atol=2e-6)
