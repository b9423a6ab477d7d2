"""Measure the arithmetic error of the fused conv kernels (fp32 SIMT, f16x3, bf16) against an fp64 reference.
Run on the GPU box:  python tools/conv_precision.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "synthetic-sleep-eeg-signal-generation-using-latent-diffusion-models_b200"))
import torch
from test_gpu_conv import _run, _ref
import eegldm

lib = eegldm.lib()
dev = torch.device("cuda", 0)
for positive in (False, True):
    for Cin in (32, 128, 512, 1024):
        g = torch.Generator().manual_seed(Cin)
        B, T, Cout, k = 2, 128, 128, 3
        x = torch.rand(B, T, Cin, generator=g) if positive else torch.randn(B, T, Cin, generator=g)
        w = (torch.rand(Cout, Cin, k, generator=g) if positive else torch.randn(Cout, Cin, k, generator=g)) / (Cin * k) ** 0.5
        ref = _ref(x.double(), w.double(), None, None, None, False, 0, None).double()
        xd = x.double().transpose(1, 2)
        ref = torch.nn.functional.conv1d(xd, w.double(), padding=1).transpose(1, 2)
        scale = ref.abs().mean().item()
        row = [f"pos={int(positive)} K={3*Cin:5d}"]
        for math in ("fp32", "f16x3", "bf16"):
            y = _run(lib, dev, x, w, None, None, None, False, 0, None, math).double()
            d = (y - ref)
            row.append(f"{math}: rms {d.pow(2).mean().sqrt().item()/scale:.2e} bias {d.mean().item()/scale:+.2e} max {d.abs().max().item()/scale:.2e}")
        print(" | ".join(row), flush=True)
