"""One launch of the tcgen05 attention kernel at the headline shape (for `ncu --set full -k regex:attn_tc_kernel`).    python tools/attn_one.py"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "synthetic-sleep-eeg-signal-generation-using-latent-diffusion-models_b200"))
import torch, eegldm
from eegldm import _lib
torch.zeros(1, device="cuda")
m = C.c_float()
_lib.check(eegldm.lib().eegldm_bench_attention(1024, 192, 1, 512, 2, C.byref(m), None, None))
print(m.value)
