"""Where the tcgen05 conv kernel's time goes, per warp role (eegldm_bench_conv_timeline): for each of the UNet's layer shapes,
the average per-CTA cycle counters -- how long the MMA warp waited for a free accumulator (epilogue-bound), for an activation
stage (producer-bound), for a weight stage (L2 / shared-memory-write bound), and how long epilogue / producers were busy.
    python tools/conv_timeline.py [--batch 1024] [--fuse 13]"""
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
ap.add_argument("--fuse", type=int, default=15)
ap.add_argument("--bn256", type=int, default=1)
ap.add_argument("--cluster", type=int, default=2)
ap.add_argument("--pair", type=int, default=1, help="CTA-pair mask (1: 256-wide launches, 2: 128-wide launches)")
ap.add_argument("--only", type=int, default=-1, help="run one row of SHAPES")
ap.add_argument("--debug", type=int, default=0, help="conv debug bits (64: one cluster alone on the device)")
a = ap.parse_args()
torch.cuda.set_device(0)
torch.zeros(1, device="cuda")
L = eegldm.lib()
_lib.check(L.eegldm_set_conv_cluster(a.cluster))
_lib.check(L.eegldm_set_conv_tuning(a.pair, a.bn256, a.fuse))
# (T, Cin, Cout, k, residual): the UNet's layer shapes (config_ldm.yaml)
SHAPES = [(768, 128, 128, 3, 0), (768, 128, 128, 3, 1), (768, 256, 128, 3, 0), (384, 256, 256, 3, 1), (384, 768, 256, 3, 0),
          (192, 512, 512, 3, 0), (192, 512, 512, 3, 1), (192, 1024, 512, 3, 0), (192, 512, 1536, 1, 0), (192, 512, 1536, 1, 2), (192, 512, 512, 1, 1)]
hdr = ["ms", "TF", "tiles", "total", "mma:acc", "mma:A", "mma:B", "mma:issue", "epi:wait", "epi:busy", "prod:wait", "prod:busy", "load:B"]
print(f"{'shape':26}" + "".join(f"{h:>10}" for h in hdr))
for (T, ci, co, k, res) in (SHAPES if a.only < 0 else [SHAPES[a.only]]):
    fl = 2.0 * ci * co * k * T * a.batch
    m = C.c_float()
    tl = (C.c_double * 16)()
    _lib.check(L.eegldm_bench_conv_timeline(a.batch, T, ci, co, k, res, 1, a.debug, a.reps, C.byref(m), tl, None))
    tiles = max(tl[9], 1.0)
    per = lambda i: tl[i] / tiles          # cycles per tile
    issue = per(0) - per(1) - per(2) - per(3)
    vals = [m.value, fl / m.value / 1e9, tl[9], per(0), per(1), per(2), per(3), issue, per(4), per(5), per(6), per(7), per(8)]
    print(f"T{T} {ci}->{co} k{k} r{res}".ljust(26) + "".join(f"{v:10.2f}" if i < 3 else f"{v:10.0f}" for i, v in enumerate(vals)), flush=True)
print("cycles per tile; mma:issue = total - the MMA warp's three waits (includes the MMAs' own back-pressure)")
