"""The autoencoder oracle (oracle/aekl.py) is pinned against the reference's in-tree class
``src/models/ae_kl.py`` -- the ancestor of the MONAI class the reference scripts instantiate -- with its always-on
non-local trio removed (what ``with_*_nonlocal_attn=False`` does upstream): (a) golden vectors that class produced
(tests/golden/make_golden_aekl.py, committed), (b) the live module when /root/reference is present."""
import os
import sys

import numpy as np
import pytest
import torch

from conftest import GOLDEN, REFERENCE
from oracle import aekl as oa
from golden.make_golden_aekl import CASES, oracle_cfg


def golden_case(name):
    g = np.load(os.path.join(GOLDEN, "aekl_golden.npz"))
    nch, mult, z, nres, B = CASES[name]
    cfg = oracle_cfg(nch, mult, z, nres)
    sd = {k: torch.from_numpy(g[f"{name}/w/{k}"]) for k in oa.aekl_param_shapes(cfg)}
    t = {k: torch.from_numpy(g[f"{name}/{k}"]) for k in ("x", "eps", "mu", "sigma", "recon")}
    return cfg, sd, t


@pytest.mark.parametrize("name", list(CASES))
def test_oracle_matches_reference_golden(name):
    cfg, sd, t = golden_case(name)
    mu, sigma = oa.encode(cfg, sd, t["x"])
    # same algorithm, same fp32 library ops: agreement to fp32 round-off, far inside rtol 1e-3 / atol 1e-4
    torch.testing.assert_close(mu, t["mu"], rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(sigma, t["sigma"], rtol=1e-5, atol=1e-5)
    z = oa.sampling(mu, sigma, t["eps"])
    torch.testing.assert_close(oa.decode(cfg, sd, z), t["recon"], rtol=1e-5, atol=1e-5)
    rec, _, _ = oa.forward(cfg, sd, t["x"], t["eps"])
    torch.testing.assert_close(rec, t["recon"], rtol=1e-5, atol=1e-5)
    assert float(t["recon"].abs().mean()) > 0.05 and float(t["sigma"].std()) > 0   # non-vacuous


@pytest.mark.reference
def test_oracle_matches_reference_module_live():
    sys.path.insert(0, os.path.join(REFERENCE, "src"))
    from golden.make_golden_aekl import build_reference, to_monai_keys
    nch, mult, z, nres = 32, (1, 2, 1), 2, 1      # ch_mult[0] must be 1: upstream's first conv already outputs num_channels[0]
    cfg = oracle_cfg(nch, mult, z, nres)
    m = build_reference(nch, mult, z, nres, seed=3)
    sd = to_monai_keys(cfg, m.state_dict())          # asserts a 1:1 key map and matching shapes
    x = torch.rand(2, 1, 256, generator=torch.Generator().manual_seed(5))
    with torch.no_grad():
        mu, sigma = m.encode(x)
        rec = m.reconstruct(x)
    omu, osig = oa.encode(cfg, sd, x)
    torch.testing.assert_close(omu, mu, rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(osig, sigma, rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(oa.decode(cfg, sd, omu), rec, rtol=1e-5, atol=1e-5)
