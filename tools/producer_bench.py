"""Where does the fused activation producer's time go?  Times single direct-mode conv launches with parts of the producer
switched off (eegldm_bench_conv debug bits: 8 no row loads, 16 no affine/SiLU, 32 no fp16 split, 64 no shared-memory stores,
128 raw segment = no scale/shift/SiLU at all) next to the pre-pass form of the same layer.
    python tools/producer_bench.py [--batch 1024]"""
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
a = ap.parse_args()
torch.cuda.set_device(0)
torch.zeros(1, device="cuda")
L = eegldm.lib()
SHAPES = [(768, 128, 128, 3, 1), (192, 512, 512, 3, 1), (192, 512, 512, 1, 1), (192, 1024, 512, 1, 0), (768, 384, 128, 1, 0)]
VARIANTS = [("pre-pass", 1, 0), ("direct", 5, 0), ("-loads", 5, 8), ("-math", 5, 16), ("-math-split", 5, 48), ("-stores", 5, 64),
            ("-loads-math-split", 5, 56), ("raw", 5, 128), ("raw-loads", 5, 136)]
print("shape".ljust(26), " ".join(f"{n:>17}" for n, _, _ in VARIANTS))
for (T, ci, co, k, res) in SHAPES:
    out = []
    for name, fuse, dbg in VARIANTS:
        _lib.check(L.eegldm_set_conv_tuning(0, 1, fuse))
        m = C.c_float()
        _lib.check(L.eegldm_bench_conv(a.batch, T, ci, co, k, res, 1, dbg, a.reps, C.byref(m), None))
        out.append(m.value)
    print(f"T{T} {ci}->{co} k{k} r{res}".ljust(26), " ".join(f"{v:17.3f}" for v in out), flush=True)
