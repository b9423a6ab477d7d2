"""The reference's sampling loop (``src/sample_trials.py:153-169``) as one fused call, plus the
batch-axis sharding used for multi-GPU sampling.

``ddim_sample``       device tensors in/out -- n_steps CUDA-graph launches + the decoder
``ddim_sample_host``  host (pinned) buffers in/out -- what bench.py's ``e2e`` leg times
``sample_sharded``    one process per GPU: rank r denoises rows [r*B/R, (r+1)*B/R) and a single
                      all-gather collects the decoded windows; no collective inside the 50 steps.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._module import check_cuda_f32
from .schedulers import DDIMScheduler


def _out_shape(unet, aekl, B, T):
    if aekl is None:
        return (B, unet.in_channels, T)
    return (B, aekl.out_channels, T * aekl._factor)


def _crop(y, crop):
    return y[:, :, crop:-crop] if crop else y   # sample_trials.py:169


def _resolve_steps(scheduler, num_inference_steps):
    """``scheduler.set_timesteps(n)`` is how the reference chooses the step count (sample_trials.py:144): honour it when the
    caller does not pass one, and refuse a silent mismatch."""
    set_n = getattr(scheduler, "num_inference_steps", None)
    if num_inference_steps is None:
        return int(set_n) if set_n else 50
    if set_n and int(set_n) != int(num_inference_steps):
        raise ValueError(f"num_inference_steps={num_inference_steps} but scheduler.set_timesteps({set_n}) was called")
    return int(num_inference_steps)


def _check_range(unet):
    if unet._math == "f16x3" and unet.range_overflow():
        raise _lib.EegldmError(-1, "f16x3: an activation left the fp16 operand range (|x| >= 65504 or NaN); the result is "
                                   "invalid -- use math='fp32' for this model")


@torch.no_grad()
def ddim_sample(unet, scheduler: DDIMScheduler, noise, num_inference_steps=None, aekl=None,
                scale_factor: float = 1.0, crop: int = 0, check_range: bool = True):
    """x_T = noise -> DDIM(num_inference_steps) -> decode(x_0 / scale_factor)[..., crop:-crop].
    ``num_inference_steps`` defaults to what ``scheduler.set_timesteps`` chose (50 if it was never called).
    ``check_range`` (f16x3 only) reads the operand-range flag after the last step -- one device synchronisation per call."""
    if scheduler.clip_sample:
        raise NotImplementedError("fused sampling implements clip_sample=False (sample_trials.py:142)")
    num_inference_steps = _resolve_steps(scheduler, num_inference_steps)
    noise = check_cuda_f32(noise, "noise")
    B, z, T = noise.shape
    out = torch.empty(_out_shape(unet, aekl, B, T), device=noise.device, dtype=torch.float32)
    with torch.cuda.device(noise.device):
        unet._sync_weights()
        if aekl is not None:
            aekl._sync_weights()
        _lib.check(_lib.lib().eegldm_ddim_sample(
            unet._h, aekl._h if aekl is not None else None, C.byref(scheduler._cfg), C.c_void_p(noise.data_ptr()),
            float(scale_factor), int(num_inference_steps), C.c_void_p(out.data_ptr()), int(B), int(T),
            C.c_void_p(_lib.current_stream_ptr(noise.device))))
        if check_range:
            _check_range(unet)
    return _crop(out, crop)


@torch.no_grad()
def ddim_sample_host(unet, scheduler: DDIMScheduler, noise_host, num_inference_steps=None, aekl=None,
                     scale_factor: float = 1.0, out_host=None, device=None):
    """Same, with HOST tensors (ideally pinned): H2D, sampling, D2H, synchronised on return."""
    if scheduler.clip_sample:
        raise NotImplementedError("fused sampling implements clip_sample=False (sample_trials.py:142)")
    num_inference_steps = _resolve_steps(scheduler, num_inference_steps)
    if noise_host.is_cuda or noise_host.dtype != torch.float32 or not noise_host.is_contiguous():
        raise ValueError("noise_host must be a contiguous fp32 CPU tensor")
    B, z, T = noise_host.shape
    shape = _out_shape(unet, aekl, B, T)
    if out_host is None:
        out_host = torch.empty(shape, dtype=torch.float32, pin_memory=True)
    elif (out_host.is_cuda or out_host.dtype != torch.float32 or not out_host.is_contiguous()
          or tuple(out_host.shape) != tuple(shape)):
        raise ValueError(f"out_host must be a contiguous fp32 CPU tensor of shape {tuple(shape)}")
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    with torch.cuda.device(device):
        unet._sync_weights()
        if aekl is not None:
            aekl._sync_weights()
        _lib.check(_lib.lib().eegldm_ddim_sample_host(
            unet._h, aekl._h if aekl is not None else None, C.byref(scheduler._cfg), C.c_void_p(noise_host.data_ptr()),
            float(scale_factor), int(num_inference_steps), C.c_void_p(out_host.data_ptr()), int(B), int(T),
            C.c_void_p(_lib.current_stream_ptr(device))))
    return out_host


def shard_range(n: int, rank: int, world: int):
    """Contiguous rows of rank ``rank`` when ``n`` windows are split over ``world`` ranks
    (the first ``n % world`` ranks take one extra row)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def sample_sharded(sample_fn, noise, group=None):
    """Shard ``noise`` [B, z, T] over the ranks of ``group`` along the batch axis, run
    ``sample_fn(local_noise) -> local_windows`` on each rank, and all-gather the windows (one
    collective, after the last denoise step).  Works with NCCL (device tensors) and gloo (CPU)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return sample_fn(noise)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    B = noise.shape[0]
    lo, hi = shard_range(B, rank, world)
    local = sample_fn(noise[lo:hi])
    counts = [shard_range(B, r, world) for r in range(world)]
    if B % world == 0:
        out = torch.empty((B,) + tuple(local.shape[1:]), device=local.device, dtype=local.dtype)
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
        return out
    maxn = max(h - l for l, h in counts)
    pad = torch.zeros((maxn,) + tuple(local.shape[1:]), device=local.device, dtype=local.dtype)
    pad[: hi - lo] = local
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    return torch.cat([b[: h - l] for b, (l, h) in zip(bufs, counts)], dim=0)
