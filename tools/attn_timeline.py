"""Where the tcgen05 attention kernel's time goes (eegldm_bench_attention): per-CTA cycle averages of the S phase (operand loads +
QK^T), the softmax, and the PV phase with its epilogue.    python tools/attn_timeline.py [--batch 1024]"""
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
print(f"{'shape':22}{'ms':>8}{'TF':>9}{'CTAs':>8}{'S phase':>10}{'softmax':>10}{'PV+epi':>10}{'total':>10}")
for (T, H, ch) in [(192, 1, 512), (128, 1, 512), (256, 1, 512), (192, 4, 128), (768, 1, 512)]:
    m = C.c_float()
    tl = (C.c_double * 8)()
    _lib.check(L.eegldm_bench_attention(a.batch if T <= 256 else max(a.batch // 8, 1), T, H, ch, a.reps, C.byref(m), tl, None))
    fl = 4.0 * (a.batch if T <= 256 else max(a.batch // 8, 1)) * T * T * H * ch
    print(f"T{T} H{H} ch{ch}".ljust(22) + f"{m.value:8.3f}{fl / m.value / 1e9:9.1f}{tl[7]:8.0f}{tl[1]:10.0f}{tl[2]:10.0f}{tl[3]:10.0f}{tl[4]:10.0f}", flush=True)
