"""Golden input/output vectors for the KL autoencoder from the REFERENCE's in-tree class.

Run in the build container (needs /root/reference):   python tests/golden/make_golden_aekl.py

The class the reference scripts instantiate (monai-generative's AutoencoderKL) is not installable here, but its
in-tree ancestor ``/root/reference/src/models/ae_kl.py`` is: same ResBlock / Downsample (pad right 1, stride 2) /
Upsample (nearest x2 + conv) / final GroupNorm -> conv / quant convs / clamp-exp encode.  It differs in three
constructor conventions only: GroupNorm has 32 groups (``ae_kl.py:16-17``; upstream: ``norm_num_groups``), channels
are ``n_channels * ch_mult`` (upstream: the absolute ``num_channels`` list), and a ResBlock-AttnBlock-ResBlock trio
is always inserted at the lowest resolution (``ae_kl.py:168-171, 220-223``; upstream: only with
``with_{en,de}coder_nonlocal_attn=True``, which every reference config sets to False).  This script builds the
in-tree model UNCHANGED, deletes that trio from the two ``nn.ModuleList``s (what the reference configs switch off),
renames its parameters to the MONAI key grammar the oracle uses and stores weights, inputs and outputs.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import aekl as oa  # noqa: E402


def _ref():
    """The reference's in-tree module (build container only)."""
    if "/root/reference/src" not in sys.path:
        sys.path.insert(0, "/root/reference/src")
    from models import ae_kl
    return ae_kl

# name: (n_channels, ch_mult, z, num_res_blocks, B)
CASES = {
    "c32_112_z1": (32, (1, 1, 2), 1, 2, 2),     # the [32, 32, 64] autoencoder of config_aekl_eeg.yaml, GroupNorm(32)
    "c32_12_z3": (32, (1, 2), 3, 1, 3),
}
L = 3072


def oracle_cfg(n_channels, ch_mult, z, nres):
    return oa.full_cfg(num_channels=[n_channels * m for m in ch_mult], latent_channels=z, num_res_blocks=nres,
                       norm_num_groups=32, attention_levels=[False] * len(ch_mult))


def build_reference(n_channels, ch_mult, z, nres, seed):
    ref = _ref()
    torch.manual_seed(seed)
    hp = dict(in_channels=1, n_channels=n_channels, z_channels=z, out_channels=1, ch_mult=ch_mult, num_res_blocks=nres,
              resolution=(L,), attn_resolutions=())
    m = ref.AutoencoderKL(embed_dim=z, hparams=hp).eval()
    # drop the always-on non-local trio (ResBlock, AttnBlock, ResBlock): with_*_nonlocal_attn=False in the configs
    enc, dec = m.encoder.blocks, m.decoder.blocks
    assert isinstance(enc[-4], ref.AttnBlock) and isinstance(dec[2], ref.AttnBlock)
    m.encoder.blocks = torch.nn.ModuleList(list(enc[:-5]) + list(enc[-2:]))
    m.decoder.blocks = torch.nn.ModuleList([dec[0]] + list(dec[4:]))
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():   # exercise the GroupNorm affine parameters (default init is 1 / 0)
        for mod in m.modules():
            if isinstance(mod, torch.nn.GroupNorm):
                mod.weight.add_(0.1 * torch.randn(mod.weight.shape, generator=g))
                mod.bias.add_(0.1 * torch.randn(mod.bias.shape, generator=g))
    return m


def to_monai_keys(cfg, ref_sd):
    """oracle (MONAI) key -> reference tensor: MONAI wraps every conv as a child called ``conv``."""
    out = {}
    used = set()
    for k in oa.aekl_param_shapes(cfg):
        head, leaf = k.rsplit(".", 1)
        rk = (head[:-len(".conv")] if head.endswith(".conv") else head) + "." + leaf
        assert rk in ref_sd, (k, rk)
        out[k] = ref_sd[rk].detach().clone()
        used.add(rk)
    assert used == set(ref_sd.keys()), set(ref_sd.keys()) - used
    return out


def main():
    torch.set_num_threads(4)
    blob = {}
    for name, (nch, mult, z, nres, B) in CASES.items():
        cfg = oracle_cfg(nch, mult, z, nres)
        m = build_reference(nch, mult, z, nres, seed=7)
        sd = to_monai_keys(cfg, m.state_dict())
        for k, shape in oa.aekl_param_shapes(cfg).items():
            assert tuple(sd[k].shape) == tuple(shape), k
        g = torch.Generator().manual_seed(11)
        x = torch.rand(B, 1, L, generator=g)
        x[..., :36] = 0
        x[..., -36:] = 0                                  # dataset.py:15,18: min-max scaled, constant-padded windows
        with torch.no_grad():
            mu, sigma = m.encode(x)
            eps = torch.randn(mu.shape, generator=g)
            zz = mu + eps * sigma                          # ae_kl.py:269-272 with the noise made explicit
            recon = m.decode(zz)
        for k, v in sd.items():
            blob[f"{name}/w/{k}"] = v.numpy()
        for k, v in dict(x=x, eps=eps, mu=mu, sigma=sigma, recon=recon).items():
            blob[f"{name}/{k}"] = v.numpy()
        # the oracle must agree with the module it restates
        r2, mu2, s2 = oa.forward(cfg, sd, x, eps)
        for a, b in ((mu2, mu), (s2, sigma), (r2, recon)):
            torch.testing.assert_close(a, b, rtol=1e-5, atol=1e-5)
        print(name, tuple(recon.shape), float(recon.abs().mean()), "params", sum(v.numel() for v in sd.values()))
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "aekl_golden.npz"), **blob)


if __name__ == "__main__":
    main()
