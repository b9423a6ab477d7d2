"""Large-batch robustness of the fused sampling call on one GPU: finite outputs, and the last 130 rows of a B-row batch equal the
same rows sampled as a small batch (windows are independent; exercises 64-bit offsets and the arena at B up to 12288)."""
import os, sys, time
ROOT = "/root/repo"
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "synthetic-sleep-eeg-signal-generation-using-latent-diffusion-models_b200"))
import torch, eegldm
from eegldm import synthetic
dev = torch.device("cuda", 0)
unet = eegldm.UNetModel(**synthetic.LDM_UNET_CFG, math="f16x3"); unet.load_state_dict(synthetic.seeded_state_dict(unet, 0)); unet = unet.to(dev).eval()
aekl = eegldm.AutoencoderKL(**synthetic.AEKL_224_CFG); aekl.load_state_dict(synthetic.seeded_state_dict(aekl, 42)); aekl = aekl.to(dev).eval()
sched = eegldm.DDIMScheduler(**synthetic.DDIM_CFG); sched.set_timesteps(3)
for B in (4096, 8192, 12288):
    noise = torch.randn(B, 1, 768, generator=torch.Generator().manual_seed(0)).to(dev)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    y = eegldm.ddim_sample(unet, sched, noise, 3, aekl)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    ref = eegldm.ddim_sample(unet, sched, noise[B - 130:], 3, aekl)      # the last rows again, as a small batch
    print(f"B={B}: {dt:.2f} s, finite {bool(torch.isfinite(y).all())}, last rows == small-batch rows: {bool(torch.equal(y[B - 130:], ref))}, "
          f"peak mem {torch.cuda.max_memory_allocated()/2**30:.1f} GiB torch-side", flush=True)
    del y, noise
