"""Generate golden loss / gradient vectors of the latent-diffusion training step from the REFERENCE's own denoiser.

Run in the build container (needs /root/reference):   python tests/golden/make_golden_ldm_train.py
Imports ``/root/reference/src/models/unet.py`` unchanged, loads the seeded weights of ``oracle.unet.make_unet_state_dict``
(strict), and runs the batch body of ``train_epoch_ldm`` (src/training/training.py:420-443) in fp32 with torch autograd through
the reference module: add_noise, model(noisy, timesteps), MSE against noise / velocity, backward.  Stored per case: z0, noise,
timesteps, the loss, and the gradient of every parameter (small cases) or a digest of it -- sum, L2 norm, the first 32 and the 32
largest-magnitude entries with their indices -- for the full config_ldm.yaml network (30.5 M parameters).
"""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, "/root/reference/src")

from models.unet import UNetModel  # noqa: E402  (the reference)
from oracle import unet as ou  # noqa: E402
from oracle.schedulers import DDPMScheduler  # noqa: E402

from train_cases import FULL_GRADS, TRAIN_CASES  # noqa: E402


def digest(g: torch.Tensor):
    f = g.flatten().double()
    top = torch.topk(f.abs(), min(32, f.numel())).indices.sort().values
    return dict(sum=np.float64(f.sum()), norm=np.float64(f.norm()), head=f[:32].float().numpy(), top_idx=top.numpy(), top=f[top].float().numpy())


def main():
    torch.set_num_threads(8)
    out = {}
    for name, (over, B, T, pred, schedule, (b0, b1)) in TRAIN_CASES.items():
        cfg = ou.full_cfg(**over)
        sd = ou.make_unet_state_dict(cfg, seed=0)
        model = UNetModel(**cfg).train()
        model.load_state_dict(sd, strict=True)
        sched = DDPMScheduler(1000, b0, b1, schedule, pred)
        g = torch.Generator().manual_seed(4321)
        z0 = torch.randn(B, cfg["in_channels"], T, generator=g)
        noise = torch.randn(B, cfg["in_channels"], T, generator=g)
        t = torch.randint(0, 1000, (B,), generator=g)
        noisy = sched.add_noise(z0, noise, t)
        pred_out = model(x=noisy, timesteps=t)
        target = sched.get_velocity(z0, noise, t) if pred == "v_prediction" else noise
        loss = F.mse_loss(pred_out.float(), target.float())
        loss.backward()
        out[name + "/z0"] = z0.numpy(); out[name + "/noise"] = noise.numpy(); out[name + "/t"] = t.numpy()
        out[name + "/loss"] = np.float32(loss.item())
        for k, p in model.named_parameters():
            if name in FULL_GRADS:
                out[f"{name}/grad/{k}"] = p.grad.numpy()
            else:
                for dk, dv in digest(p.grad).items():
                    out[f"{name}/digest/{k}/{dk}"] = dv
        print(name, float(loss), sum(p.numel() for p in model.parameters()))
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "ldm_train_golden.npz"), **out)


if __name__ == "__main__":
    main()
