"""Generate golden input/output vectors for the denoiser from the REFERENCE's own module.

Run in the build container (needs /root/reference):   python tests/golden/make_golden.py
It imports ``/root/reference/src/models/unet.py`` unchanged, loads the seeded synthetic weights of
``oracle.unet.make_unet_state_dict`` with ``strict=True`` (so the oracle's key grammar is checked
against the real module) and stores x, timesteps and UNetModel.forward's output.  The weights are
re-generated from the seed at test time and are not stored.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, "/root/reference/src")

from models.unet import UNetModel  # noqa: E402  (the reference)
from oracle import unet as ou  # noqa: E402

from cases import CASES  # noqa: E402


def main():
    torch.set_num_threads(4)
    out = {}
    for name, (over, B, T, ts) in CASES.items():
        cfg = ou.full_cfg(**over)
        sd = ou.make_unet_state_dict(cfg, seed=0)
        model = UNetModel(**cfg).eval()
        model.load_state_dict(sd, strict=True)
        g = torch.Generator().manual_seed(1234)
        x = torch.randn(B, cfg["in_channels"], T, generator=g)
        t = torch.tensor(ts, dtype=torch.float32 if any(isinstance(v, float) for v in ts) else torch.long)
        with torch.no_grad():
            y = model(x, timesteps=t)
        out[name + "/x"] = x.numpy()
        out[name + "/t"] = t.numpy()
        out[name + "/y"] = y.numpy()
        print(name, tuple(y.shape), float(y.abs().mean()))
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "unet_golden.npz"), **out)


if __name__ == "__main__":
    main()
