"""Per-launch device time of one UNet forward (eager launches, CUDA events on the launch stream).
    python tools/layer_profile.py [--batch 1024] [--math f16x3]"""
import argparse, ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "synthetic-sleep-eeg-signal-generation-using-latent-diffusion-models_b200"))
import torch
import eegldm
from eegldm import _lib
from oracle import unet as ou

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=1024)
ap.add_argument("--math", default="f16x3")
ap.add_argument("--cluster", type=int, default=2)
ap.add_argument("--warmup", type=int, default=2, help="untimed forwards before the profiled one (0 under ncu)")
ap.add_argument("--pair", type=int, default=1)
ap.add_argument("--bn256", type=int, default=1)
ap.add_argument("--gnfuse", type=int, default=15)
ap.add_argument("--length", type=int, default=768, help="input length (3072: the raw-signal DM variant, attention at T = 768)")
a = ap.parse_args()
dev = torch.device("cuda", 0)
_lib.check(eegldm.lib().eegldm_set_conv_cluster(a.cluster))
_lib.check(eegldm.lib().eegldm_set_conv_tuning(a.pair, a.bn256, a.gnfuse))
cfg = ou.full_cfg()
m = eegldm.UNetModel(**cfg, math=a.math)
m.load_state_dict(ou.make_unet_state_dict(cfg, 0))
m = m.to(dev).eval()
x = torch.randn(a.batch, 1, a.length, device=dev)
t = torch.tensor([500])
for _ in range(a.warmup):
    m(x, timesteps=t)
torch.cuda.synchronize()
L = eegldm.lib()
L.eegldm_profile_enable(1)
m(x, timesteps=t)
torch.cuda.synchronize()
names = ["conv", "gn", "attn", "other", "split", "cnarrow"]
i = 0
tot = {}
print(f"{'#':>4} {'kind':6} {'ms':>8} {'GFLOP':>9} {'TFLOP/s':>8} {'GB/s':>8}")
while True:
    k, ms, fl, by = C.c_int(), C.c_double(), C.c_double(), C.c_double()
    if L.eegldm_profile_record(i, C.byref(k), C.byref(ms), C.byref(fl), C.byref(by)) != 0:
        break
    n = names[k.value]
    tot[n] = tot.get(n, 0) + ms.value
    if n in ("conv", "attn") or ms.value > 0.3:
        print(f"{i:4d} {n:6} {ms.value:8.3f} {fl.value/1e9:9.1f} {fl.value/ms.value/1e9 if ms.value else 0:8.1f} {by.value/ms.value/1e6 if ms.value else 0:8.0f}")
    i += 1
L.eegldm_profile_enable(0)
print("totals ms:", {k: round(v, 2) for k, v in tot.items()}, "sum", round(sum(tot.values()), 2))
