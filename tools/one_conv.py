"""One tcgen05 convolution shape, a few launches (for `ncu --set full --import-source on -k regex:conv_tc_kernel`).
    python tools/one_conv.py T Cin Cout k residual [--batch 1024] [--fuse 13]"""
import argparse, ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "synthetic-sleep-eeg-signal-generation-using-latent-diffusion-models_b200"))
import torch
import eegldm
from eegldm import _lib

ap = argparse.ArgumentParser()
ap.add_argument("shape", type=int, nargs=5)
ap.add_argument("--batch", type=int, default=1024)
ap.add_argument("--fuse", type=int, default=15)
ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--debug", type=int, default=0)
ap.add_argument("--pair", type=int, default=1)
a = ap.parse_args()
torch.cuda.set_device(0)
torch.zeros(1, device="cuda")
L = eegldm.lib()
_lib.check(L.eegldm_set_conv_tuning(a.pair, 1, a.fuse))
T, ci, co, k, res = a.shape
m = C.c_float()
_lib.check(L.eegldm_bench_conv(a.batch, T, ci, co, k, res, 1, a.debug, a.reps, C.byref(m), None))
print(f"T{T} {ci}->{co} k{k} r{res}: {m.value:.3f} ms, {2.0 * ci * co * k * T * a.batch / m.value / 1e9:.1f} TFLOP/s")
