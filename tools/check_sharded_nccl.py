"""Config 4 correctness on real GPUs (run under torchrun, one rank per GPU): the batch-sharded DDIM sampling gathered with
one NCCL all-gather equals the same batch sampled on ONE GPU, row for row and bit for bit (windows are independent; every
rank runs the same kernels on its rows).
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/check_sharded_nccl.py [--batch 37]"""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "synthetic-sleep-eeg-signal-generation-using-latent-diffusion-models_b200"))
import torch
import torch.distributed as dist
import eegldm
from eegldm import synthetic

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=37)     # ragged over 2, 4 and 8 ranks
ap.add_argument("--steps", type=int, default=10)
a = ap.parse_args()
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
dist.init_process_group("nccl", device_id=dev)
unet = eegldm.UNetModel(**synthetic.LDM_UNET_CFG, math="f16x3")
unet.load_state_dict(synthetic.seeded_state_dict(unet, 0))
unet = unet.to(dev).eval()
aekl = eegldm.AutoencoderKL(**synthetic.AEKL_224_CFG)
aekl.load_state_dict(synthetic.seeded_state_dict(aekl, 42))
aekl = aekl.to(dev).eval()
sched = eegldm.DDIMScheduler(**synthetic.DDIM_CFG)
sched.set_timesteps(a.steps)
noise = torch.randn(a.batch, 1, 768, generator=torch.Generator().manual_seed(0)).to(dev)
fn = lambda n: eegldm.ddim_sample(unet, sched, n, a.steps, aekl, crop=36)
for B in (a.batch, a.batch - a.batch % world):       # ragged (all_gather of padded shards) and even (all_gather_into_tensor)
    out = eegldm.sample_sharded(fn, noise[:B])
    if rank == 0:
        full = fn(noise[:B])
        same = out.shape == full.shape and bool(torch.equal(out, full))
        print(f"world {world} batch {B}: gathered {tuple(out.shape)} == single-GPU {tuple(full.shape)}: {same}; "
              f"finite {bool(torch.isfinite(out).all())}; max |diff| {float((out - full).abs().max()):.3g}", flush=True)
        assert same
dist.barrier()
dist.destroy_process_group()
