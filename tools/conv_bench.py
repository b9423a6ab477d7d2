"""Timing of single tcgen05 convolution launches over tile shapes and the two timing experiments
(no operand copies / no MMAs) that separate tensor-pipe time from operand-staging time.
    python tools/conv_bench.py [--batch 1024]"""
import argparse, ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "synthetic-sleep-eeg-signal-generation-using-latent-diffusion-models_b200"))
import torch
import eegldm
from eegldm import _lib

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=1024)
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--shapes", default="all")
a = ap.parse_args()
torch.cuda.set_device(0)
torch.zeros(1, device="cuda")
L = eegldm.lib()
# (T, Cin, Cout, k, residual): the UNet's layer shapes (config_ldm.yaml)
SHAPES = [(768, 128, 128, 3, 1), (384, 256, 256, 3, 1), (192, 512, 512, 3, 1), (192, 512, 1536, 1, 0), (192, 512, 512, 1, 1),
          (192, 1024, 512, 3, 0), (384, 768, 256, 3, 0)]
MODES = [("cg1 cl2", 0, 2), ("cg1 cl1", 0, 1), ("pair", 1, 2)]
print(f"{'shape':28} {'mode':8} {'bn':>4} {'ms':>8} {'TF':>7} {'noload':>8} {'TF':>7} {'nomma':>8} {'no-L2pf':>8}")
for (T, ci, co, k, res) in SHAPES:
    fl = 2.0 * ci * co * k * T * a.batch
    for name, pair, cl in MODES:
        for bn256 in (1, 10 ** 6):
            if bn256 == 1 and co % 256:
                continue
            _lib.check(L.eegldm_set_conv_cluster(cl))
            _lib.check(L.eegldm_set_conv_tuning(pair, bn256, 1))
            ms = []
            for dbg in (0, 1, 2, 4):
                m = C.c_float()
                _lib.check(L.eegldm_bench_conv(a.batch, T, ci, co, k, res, 1, dbg, a.reps, C.byref(m), None))
                ms.append(m.value)
            print(f"T{T} {ci}->{co} k{k} r{res}".ljust(28), f"{name:8} {256 if bn256 == 1 else 128:4d} {ms[0]:8.3f} {fl/ms[0]/1e9:7.1f} "
                  f"{ms[1]:8.3f} {fl/ms[1]/1e9:7.1f} {ms[2]:8.3f} {ms[3]:8.3f}", flush=True)
