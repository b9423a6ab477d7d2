"""Per-step wall time of the device-resident and the end-to-end sampling call, interleaved (diagnostic)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "synthetic-sleep-eeg-signal-generation-using-latent-diffusion-models_b200"))
import torch
import eegldm
from oracle import aekl as oa, unet as ou
from oracle.sample import SAMPLER_DEFAULTS

dev = torch.device("cuda", 0)
ucfg, acfg = ou.full_cfg(), oa.full_cfg()
unet = eegldm.UNetModel(**ucfg, math="f16x3"); unet.load_state_dict(ou.make_unet_state_dict(ucfg, 0)); unet = unet.to(dev).eval()
aekl = eegldm.AutoencoderKL(**acfg); aekl.load_state_dict(oa.make_aekl_state_dict(acfg, 42)); aekl = aekl.to(dev).eval()
sched = eegldm.DDIMScheduler(**SAMPLER_DEFAULTS); sched.set_timesteps(50)
B = 1024
nh = torch.randn(B, 1, 768).pin_memory(); nd = nh.to(dev)
oh = torch.empty(B, 1, 3072).pin_memory()
def t(fn):
    torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); return (time.perf_counter() - t0) * 1e3
dev_step = lambda: eegldm.ddim_sample(unet, sched, nd, 50, aekl)
e2e_step = lambda: eegldm.ddim_sample_host(unet, sched, nh, 50, aekl, out_host=oh, device=dev)
for i in range(2): dev_step()
for i in range(6):
    print(f"round {i}: device {t(dev_step):8.1f} ms   e2e {t(e2e_step):8.1f} ms", flush=True)
