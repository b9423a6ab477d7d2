"""N>1 host logic on CPU: world_size-2 gloo run of the batch-axis sharding + single all-gather used by
multi-GPU sampling.  The per-rank sampler is a stand-in (the oracle's DDIM on a
tiny UNet) because the CUDA path needs a GPU; what is tested is that sharded == unsharded row for row."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import PKG, ROOT


def _worker(rank, world, port, B, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, PKG)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(1)
        from eegldm.sampler import sample_sharded
        from oracle import sample as osamp
        from oracle import unet as ou
        cfg = ou.full_cfg(model_channels=32, channel_mult=[1, 2], attention_resolutions=[2], image_size=32)
        sd = ou.make_unet_state_dict(cfg, 0)
        noise = torch.randn(B, 1, 32, generator=torch.Generator().manual_seed(0))
        fn = lambda n: osamp.ddim_sample(cfg, sd, n, n_steps=4)
        out = sample_sharded(fn, noise)
        if rank == 0:
            full = fn(noise)
            q.put((out.shape == full.shape, float((out - full).abs().max())))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("B", [4, 5])
def test_sharded_sampling_equals_unsharded_gloo(B):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + B
    procs = [ctx.Process(target=_worker, args=(r, 2, port, B, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok, err = q.get(timeout=180)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert ok
    assert err < 1e-5   # per-sample independence: identical up to batch-size dependent CPU conv round-off
