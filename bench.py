#!/usr/bin/env python
"""Headline benchmark: EEG windows/s for DDIM-50 sampling (BASELINE.json metric), config 3 of
SURVEY.md section 8d -- batch 1024 of [1,768] latents, config_ldm.yaml UNet, AEKL 2-2-4 decode to
[1,3072] -- per GPU (weak scaling), synthetic noise, seeded random weights.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B] [--math M]

A "step" is one full sampling pass over one batch: 50 denoise iterations (one CUDA graph launch each)
+ the autoencoder decode (+ one NCCL all-gather of the decoded windows when N > 1).
Prints ONE JSON line on rank 0.  See DESIGN.md section "Measurement" for every field.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "synthetic-sleep-eeg-signal-generation-using-latent-diffusion-models_b200")
for _p in (ROOT, PKG):
    if _p not in sys.path:
        sys.path.insert(0, _p)

# algorithmic work per window: UNet 13.90 GFLOP and 25.2 MB block-fused
# fp32 activation traffic per forward per sample
UNET_GFLOP_PER_FWD = 13.90
UNET_MB_PER_FWD = 25.2
DDIM_STEPS = 50
T_LATENT = 768


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tc_burst=d["bf16_tflops"], tc_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    src="measured")
    return dict(hbm=6650.0, tc_burst=1590.0, tc_sustained=1400.0, src="fallback")   # B200_PROFILING.md fallback


NCU_FORWARD_CSVS = ("r02b_ncu_forward_b1024.csv", "r02_ncu_forward_b1024.csv", "r01_ncu_forward_b1024.csv")   # newest committed capture wins


def _ncu_conv_traffic(batch, math):
    """(bytes, n_launches, source): dram__bytes_read.sum + dram__bytes_write.sum per conv_tc_kernel launch, averaged over
    the tensor-pipe conv launches of ONE UNet forward, from the committed `ncu --set full` pass of a forward at B=1024 in
    f16x3 (ncu replays kernels, so this cannot be measured inside a timed run).  The same launch set -- the tcgen05 convs,
    profile family 0 -- is what `algorithmic_bytes_per_launch` averages, so the two are like for like."""
    if batch != 1024 or math != "f16x3":
        return None, 0, None
    for name in NCU_FORWARD_CSVS:
        path = os.path.join(ROOT, "profiles", name)
        if not os.path.exists(path):
            continue
        n, tot = 0, 0.0
        for line in open(path):
            f = line.strip().split(",")
            if len(f) >= 9 and f[1].startswith("conv_tc_kernel"):
                try:
                    tot += (float(f[-5]) + float(f[-4])) * 1e6
                    n += 1
                except ValueError:
                    pass
        if n:
            return tot / n, n, "profiles/" + name
    return None, 0, None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc = index, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=10)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        sm, smax, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for n, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(smax), "reasons": sorted(reasons),
                "power_w_max": max(pw), "samples": len(sm)}


def _oracle_models():
    from oracle import aekl as oa, unet as ou
    ucfg, acfg = ou.full_cfg(), oa.full_cfg()
    return ucfg, ou.make_unet_state_dict(ucfg, 0), acfg, oa.make_aekl_state_dict(acfg, 42)


def _cpu_ddim_windows_per_s(B, reps, warm):
    """The reference's CPU implementation of the path (oracle port: src/models/unet.py restated +
    restated MONAI AEKL/DDIM), all host threads, fp32, no_grad.  Returns (windows/s list, threads)."""
    import torch
    from oracle import sample as osamp
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    ucfg, usd, acfg, asd = _oracle_models()
    noise = torch.randn(B, 1, T_LATENT, generator=torch.Generator().manual_seed(0))
    vals = []
    for i in range(warm + reps):
        t0 = time.perf_counter()
        osamp.ddim_sample(ucfg, usd, noise, DDIM_STEPS, acfg, asd, crop=36)
        dt = time.perf_counter() - t0
        if i >= warm:
            vals.append(B / dt)
    return vals, threads


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    B = args.ref_batch
    warm = min(args.warmup, 3)
    t0 = time.perf_counter()
    vals, threads = _cpu_ddim_windows_per_s(B, args.steps, warm)
    total = time.perf_counter() - t0
    v = B * len(vals) / sum(B / x for x in vals)
    line = {
        "impl": "reference", "metric": "EEG windows/sec DDIM-50 sampling", "value": v, "unit": "windows/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": warm, "ms_per_step": 1e3 * B / v,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"DDIM-50 sampling + AEKL 2-2-4 decode, config_ldm.yaml UNet, [B,1,768] latents -> [B,1,3000] windows; "
                               f"bounded sample of B={B} windows per step on the host CPU", "batch": B},
        "cpu_baseline": {"value": v, "unit": "windows/s", "cores": threads, "kind": "port", "batch": B,
                         "sample": f"{B} windows x DDIM-50 + decode per step, {len(vals)} steps, torch {threads} threads"},
        "e2e": {"value": v, "unit": "windows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": total,
    }
    print(json.dumps(line), flush=True)


def _events_ms(torch, fn, reps, warm):
    """average milliseconds of fn() over `reps` calls after `warm` untimed ones (CUDA events on the current stream)"""
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def _profile_families(eegldm, _lib, C):
    L = eegldm.lib()
    fam = {}
    for kind, name in enumerate(("conv", "groupnorm", "attention", "other", "act_split", "conv_narrow")):
        m, f, b, n = C.c_double(), C.c_double(), C.c_double(), C.c_int64()
        _lib.check(L.eegldm_profile_read(kind, C.byref(m), C.byref(f), C.byref(b), C.byref(n)))
        fam[name] = dict(ms=m.value, flops=f.value, bytes=b.value, launches=n.value)
    return fam


def measure_train(args, dev, B=512, with_cpu=True):
    """Config 5: AutoencoderKL training step (config_aekl_eeg_2_2_4_spec.yaml: 2-2-4, z=1), batch 512 x [1,3072]:
    generator loss L1 + 1e-9 KL + 1e4 Jukebox (+ adversarial term when the engine has the discriminator), Adam lr 5e-3."""
    import torch
    import eegldm
    from eegldm import synthetic
    cfg = dict(synthetic.AEKL_224_CFG)
    m = eegldm.AutoencoderKL(**cfg)
    sd = synthetic.seeded_state_dict(m, 42)
    m.load_state_dict(sd)
    m = m.to(dev)
    xh = torch.rand(B, 1, 3072, generator=torch.Generator().manual_seed(0)).pin_memory()
    eh = torch.randn(B, 1, 768, generator=torch.Generator().manual_seed(1)).pin_memory()
    x, eps = xh.to(dev), eh.to(dev)
    adv = hasattr(m, "attach_discriminator") and not args.no_adversarial
    if adv:
        m.attach_discriminator(seed=7)
    steps = max(args.steps, 10)
    l0 = eegldm.launch_count()
    ms = _events_ms(torch, lambda: m.train_step(x, eps, return_losses=False), steps, max(args.warmup, 3))
    launches = (eegldm.launch_count() - l0) // (steps + max(args.warmup, 3))
    t0 = time.perf_counter()
    for _ in range(steps):   # e2e: batch from pinned host memory, losses read back every step
        m.train_step(xh.to(dev, non_blocking=True), eh.to(dev, non_blocking=True))
    torch.cuda.synchronize()
    ms_e2e = (time.perf_counter() - t0) * 1e3 / steps
    peaks = _peaks()
    # algorithmic HBM bytes per window of the autoencoder step: every activation of the 2-2-4 autoencoder written once and read
    # once in the forward pass and once more in the backward pass, gradients written + read once (fp32)
    acts = 4 * (3072 * (1 + 2 * 9) + 1536 * 2 * 9 + 768 * (4 * 9 + 4))
    out = {"metric": "AEKL training-step windows/sec", "value": B / (ms / 1e3), "unit": "windows/s", "ms_per_step": ms,
           "workload": "config 5: AutoencoderKL 2-2-4 z=1 training step, batch %d x [1,3072], L1 + 1e-9 KL + 1e4 Jukebox%s, Adam lr 5e-3"
                       % (B, " + 0.005 adversarial (PatchDiscriminator 3 x 64, its own Adam step)" if adv else " (adversarial term excluded)"),
           "adversarial": bool(adv), "gpu_launches_per_step": int(launches),
           "e2e": {"value": B / (ms_e2e / 1e3), "unit": "windows/s", "h2d_bytes_per_step": (xh.numel() + eh.numel()) * 4,
                   "d2h_bytes_per_step": 16, "ms_per_step": ms_e2e},
           "roofline": {"bound": "hbm", "achieved": 6 * acts * B / (ms / 1e3) / 1e9, "peak": peaks["hbm"], "unit": "GB/s",
                        "frac": 6 * acts * B / (ms / 1e3) / 1e9 / peaks["hbm"], "traffic": None,
                        "note": "autoencoder activations only (6 passes x fp32); with the discriminator the step is conv-FLOP bound, "
                                "see DESIGN.md section 4.5"}}
    if with_cpu:
        # CPU baseline: the oracle with torch autograd + Adam, bounded sample (the only use of oracle/ in this function); with the
        # discriminator attached it is the same full step (generator and discriminator halves, train_autoencoderkl.py:204-234)
        from oracle import aekl as oa, jukebox as oj, discriminator as od
        threads = os.cpu_count() or 1
        torch.set_num_threads(threads)
        Bc, n = (4, 2) if adv else (16, 5)
        params = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
        opt = torch.optim.Adam(list(params.values()), lr=5e-3)
        xc, ec = xh[:Bc].clone(), eh[:Bc].clone()
        if adv:
            dcfg = od.full_cfg()
            dp = {k: (v.clone().requires_grad_(True) if not od.is_buffer(k) else v.clone()) for k, v in od.make_disc_state_dict(dcfg, 7).items()}
            opt_d = torch.optim.Adam([v for k, v in dp.items() if not od.is_buffer(k)], lr=5e-4)

        def cpu_step():
            opt.zero_grad(set_to_none=True)
            recon, mu, sigma = oa.forward(cfg, params, xc, ec)
            loss = torch.nn.functional.l1_loss(recon, xc) + 1e-9 * oa.kl_loss(mu, sigma) + 1e4 * oj.jukebox_loss(recon, xc)
            if adv:
                loss = loss + 0.01 * od.patch_adversarial_loss(od.forward(dcfg, dp, recon.contiguous())[-1], True, False)
            loss.backward()
            opt.step()
            if adv:
                opt_d.zero_grad(set_to_none=True)
                lf = od.patch_adversarial_loss(od.forward(dcfg, dp, recon.contiguous().detach())[-1], False, True)
                lr_ = od.patch_adversarial_loss(od.forward(dcfg, dp, xc.contiguous())[-1], True, True)
                (0.01 * 0.5 * (lf + lr_)).backward()
                opt_d.step()
        cpu_step()
        t0 = time.perf_counter()
        for _ in range(n):
            cpu_step()
        out["cpu_baseline"] = {"value": Bc * n / (time.perf_counter() - t0), "unit": "windows/s", "cores": threads, "kind": "port",
                               "batch": Bc, "sample": f"{Bc} windows x {n} steps, oracle autoencoder" + (" + discriminator" if adv else "") +
                                                      " forward, torch autograd, Adam" + ("" if adv else " (generator half only)")}
    return out


def measure_config1(args, dev):
    """Config 1: AutoencoderKL encode + decode of one batch [4,1,3072] (config_aekl_eeg.yaml with num_channels [32,32,64], z=1)."""
    import torch
    import eegldm
    from eegldm import synthetic
    cfg = dict(synthetic.AEKL_224_CFG, num_channels=[32, 32, 64])
    m = eegldm.AutoencoderKL(**cfg)
    sd = synthetic.seeded_state_dict(m, 42)
    m.load_state_dict(sd)
    m = m.to(dev).eval()
    B = 4
    xh = torch.rand(B, 1, 3072, generator=torch.Generator().manual_seed(0))
    xh[..., :36] = 0
    xh[..., -36:] = 0
    x = xh.to(dev)

    def step():
        mu, sigma = m.encode(x)
        return m.decode(mu)
    ms = _events_ms(torch, step, 20, 5)
    peaks = _peaks()
    alg = 5.13e6 * 2 * B   # block-fused fp32 activation bytes, encode + decode (5.13 MB / sample / direction)
    out = {"workload": "config 1: AutoencoderKL [32,32,64] z=1 encode + decode, batch 4 x [1,3072]", "value": B / (ms / 1e3),
           "unit": "windows/s", "ms_per_step": ms,
           "roofline": {"bound": "hbm", "achieved": alg / (ms / 1e3) / 1e9, "peak": peaks["hbm"], "unit": "GB/s",
                        "frac": alg / (ms / 1e3) / 1e9 / peaks["hbm"], "traffic": None,
                        "note": "4 windows cannot fill 148 SMs: launch-latency bound (about 60 launches of a few microseconds)"}}
    from oracle import aekl as oa
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    with torch.no_grad():
        def cpu():
            mu, sigma = oa.encode(cfg, sd, xh)
            return oa.decode(cfg, sd, mu)
        cpu()
        t0 = time.perf_counter()
        n = 10
        for _ in range(n):
            cpu()
        dt = (time.perf_counter() - t0) / n
    out["cpu_baseline"] = {"value": B / dt, "unit": "windows/s", "cores": threads, "kind": "port", "batch": B,
                           "sample": f"the same batch of {B}, oracle encode + decode x {n}, torch {threads} threads"}
    return out


def measure_config2(args, dev, unet):
    """Config 2: ONE UNet denoise step (UNetModel.forward), batch 256 x [1,768], per-sample timesteps on the device."""
    import torch
    B = 256
    x = torch.randn(B, 1, T_LATENT, generator=torch.Generator().manual_seed(0)).to(dev)
    t = torch.randint(0, 1000, (B,), generator=torch.Generator().manual_seed(1)).to(dev)
    ms = _events_ms(torch, lambda: unet(x, timesteps=t), 10, 3)
    peaks = _peaks()
    tf = B * UNET_GFLOP_PER_FWD * 1e9 / (ms / 1e3) / 1e12
    out = {"workload": "config 2: one UNet denoise step, batch 256 x [1,768] latents, config_ldm.yaml, timesteps [256] on the device",
           "value": B / (ms / 1e3), "unit": "UNet evaluations/s", "ms_per_step": ms,
           "roofline": {"bound": "tensor", "achieved": tf, "peak": peaks["tc_sustained"], "unit": "TFLOP/s", "frac": tf / peaks["tc_sustained"],
                        "traffic": None, "note": "whole forward (13.90 GFLOP algorithmic per sample) against the sustained bf16 peak"}}
    from oracle import unet as ou
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    ucfg = ou.full_cfg()
    usd = ou.make_unet_state_dict(ucfg, 0)
    Bc = 8
    xc, tc = x[:Bc].cpu(), t[:Bc].cpu()
    with torch.no_grad():
        ou.unet_forward(ucfg, usd, xc, tc)
        t0 = time.perf_counter()
        n = 3
        for _ in range(n):
            ou.unet_forward(ucfg, usd, xc, tc)
        dt = (time.perf_counter() - t0) / n
    out["cpu_baseline"] = {"value": Bc / dt, "unit": "UNet evaluations/s", "cores": threads, "kind": "port", "batch": Bc,
                           "sample": f"{Bc} windows x {n} forwards, oracle port of src/models/unet.py, torch {threads} threads"}
    return out


def measure_dm_variant(args, dev, unet):
    """one denoise step of the raw-signal DM variant (sample_trials_ddpm.py:83-102): the same UNet on [B,1,3072], attention at
    T = 768 (tcgen05 attention split over three key blocks).  No CPU leg (the oracle needs minutes per window at this length)."""
    import torch
    B = 64
    x = torch.randn(B, 1, 4 * T_LATENT, generator=torch.Generator().manual_seed(0)).to(dev)
    t = torch.randint(0, 1000, (B,), generator=torch.Generator().manual_seed(1)).to(dev)
    ms = _events_ms(torch, lambda: unet(x, timesteps=t), 5, 3)
    peaks = _peaks()
    # 4 x the latent forward (convs scale with the length) + the attention's extra share: six blocks of 4 T^2 C FLOPs at T = 768 are 7.25
    # GFLOP, of which 4 x 0.453 are already in the first term
    gflop = 4 * UNET_GFLOP_PER_FWD + 5.44
    tf = B * gflop * 1e9 / (ms / 1e3) / 1e12
    return {"workload": "8f-4: one UNet denoise step of the raw-signal DM variant, batch 64 x [1,3072], attention at T = 768",
            "value": B / (ms / 1e3), "unit": "UNet evaluations/s", "ms_per_step": ms,
            "roofline": {"bound": "tensor", "achieved": tf, "peak": peaks["tc_sustained"], "unit": "TFLOP/s", "frac": tf / peaks["tc_sustained"],
                         "traffic": None, "note": "whole forward (%.1f GFLOP algorithmic per sample) against the sustained bf16 peak" % gflop}}


def measure_ldm_train(args, dev, B=64, with_cpu=True):
    """one latent-diffusion training step (training.py:420-443) of the config_ldm.yaml UNet, batch B x [1,768] latents:
    add_noise, forward, MSE against the noise, backward, Adam lr 1e-4; tensor-pipe convs in f16x3 (forward, data and weight gradients)."""
    import torch
    import eegldm
    from eegldm import synthetic
    m = eegldm.UNetModel(**synthetic.LDM_UNET_CFG, math="f16x3")
    m.load_state_dict(synthetic.seeded_state_dict(m, 0))
    m = m.to(dev)
    sched = eegldm.DDPMScheduler(1000, 0.0015, 0.0195, "linear_beta", "epsilon")   # train_ldm.py:199-200
    g = torch.Generator().manual_seed(0)
    zh = torch.randn(B, 1, T_LATENT, generator=g).pin_memory()
    nh = torch.randn(B, 1, T_LATENT, generator=g).pin_memory()
    th = torch.randint(0, 1000, (B,), generator=g).pin_memory()
    z, n, t = zh.to(dev), nh.to(dev), th.to(dev)
    steps, warm = 5, 3
    l0 = eegldm.launch_count()
    ms = _events_ms(torch, lambda: m.train_step(z, n, t, sched, return_loss=False), steps, warm)
    launches = (eegldm.launch_count() - l0) // (steps + warm)
    t0 = time.perf_counter()
    for _ in range(steps):   # e2e: latents, noise and timesteps from pinned host memory, the loss read back every step
        m.train_step(zh.to(dev, non_blocking=True), nh.to(dev, non_blocking=True), th.to(dev, non_blocking=True), sched)
    torch.cuda.synchronize()
    ms_e2e = (time.perf_counter() - t0) * 1e3 / steps
    peaks = _peaks()
    tf = 3 * UNET_GFLOP_PER_FWD * 1e9 * B / (ms / 1e3) / 1e12   # forward + data gradient + weight gradient
    out = {"metric": "LDM training-step windows/sec", "value": B / (ms / 1e3), "unit": "windows/s", "ms_per_step": ms,
           "workload": "8f-2: latent-diffusion training step, config_ldm.yaml UNet (30.5M params), batch %d x [1,768] latents, epsilon target, "
                       "MSE, backward, Adam lr 1e-4, f16x3" % B,
           "gpu_launches_per_step": int(launches),
           "e2e": {"value": B / (ms_e2e / 1e3), "unit": "windows/s", "h2d_bytes_per_step": (zh.numel() + nh.numel()) * 4 + th.numel() * 8,
                   "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e},
           "roofline": {"bound": "tensor", "achieved": tf, "peak": peaks["tc_sustained"], "unit": "TFLOP/s", "frac": tf / peaks["tc_sustained"],
                        "traffic": None, "note": "3 x 13.90 GFLOP algorithmic per window (forward, data gradient, weight gradient) against the "
                                                 "sustained bf16 peak; f16x3 issues 3 products per MAC"}}
    if with_cpu:
        from oracle import ldm_train as ol, unet as ou
        from oracle.schedulers import DDPMScheduler
        threads = os.cpu_count() or 1
        torch.set_num_threads(threads)
        ucfg = ou.full_cfg()
        usd = ou.make_unet_state_dict(ucfg, 0)
        Bc = 4
        osched = DDPMScheduler(1000, 0.0015, 0.0195, "linear_beta", "epsilon")
        ol.ldm_train_step(ucfg, usd, zh[:Bc], nh[:Bc], th[:Bc], osched)
        t0 = time.perf_counter()
        nrep = 2
        for _ in range(nrep):
            ol.ldm_train_step(ucfg, usd, zh[:Bc], nh[:Bc], th[:Bc], osched)
        dt = (time.perf_counter() - t0) / nrep
        out["cpu_baseline"] = {"value": Bc / dt, "unit": "windows/s", "cores": threads, "kind": "port", "batch": Bc,
                               "sample": f"{Bc} windows x {nrep} steps, oracle port (unet.py restated) + torch autograd + Adam, {threads} threads"}
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist
    import eegldm
    from eegldm import _lib, synthetic
    import ctypes as C

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: eegldm has no CPU fallback")
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # keep stdout to the one JSON line (NCCL prints its version)
        dist.init_process_group("nccl", device_id=dev)

    # seeded random-init weights of the reference's architecture (no checkpoint ships); product-side helper, no oracle/ here
    _lib.check(eegldm.lib().eegldm_set_sample_lanes(args.lanes))
    _lib.check(eegldm.lib().eegldm_set_conv_tuning(args.pair, 1, args.fuse))
    unet = eegldm.UNetModel(**synthetic.LDM_UNET_CFG, math=args.math)
    usd = synthetic.seeded_state_dict(unet, 0)
    unet.load_state_dict(usd)
    unet = unet.to(dev).eval()
    aekl = eegldm.AutoencoderKL(**synthetic.AEKL_224_CFG)
    aekl.load_state_dict(synthetic.seeded_state_dict(aekl, 42))
    aekl = aekl.to(dev).eval()
    sched = eegldm.DDIMScheduler(**synthetic.DDIM_CFG)
    sched.set_timesteps(DDIM_STEPS)

    B = args.batch
    noise_host = torch.randn(B, 1, T_LATENT, generator=torch.Generator().manual_seed(rank)).pin_memory()
    noise = noise_host.to(dev)
    gathered = torch.empty((world * B, 1, 3072), device=dev) if world > 1 else None
    out_host = torch.empty((B, 1, 3072), dtype=torch.float32).pin_memory()

    def step_device():
        y = eegldm.ddim_sample(unet, sched, noise, DDIM_STEPS, aekl, check_range=False)
        if world > 1:
            dist.all_gather_into_tensor(gathered, y)      # the single collective
        return y

    if world == 1:
        def step_e2e():   # the reference-facing C-ABI call with HOST buffers: H2D, 50 graph launches, decode, D2H, synchronised
            eegldm.ddim_sample_host(unet, sched, noise_host, DDIM_STEPS, aekl, out_host=out_host, device=dev)
    else:
        stage = torch.empty_like(noise)

        def step_e2e():   # H2D of this rank's noise, sampling, the all-gather, D2H of this rank's rows -- all inside the timed region
            stage.copy_(noise_host, non_blocking=True)
            y = eegldm.ddim_sample(unet, sched, stage, DDIM_STEPS, aekl, check_range=False)
            dist.all_gather_into_tensor(gathered, y)
            out_host.copy_(y, non_blocking=True)
            torch.cuda.synchronize(dev)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, k):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)     # max over ranks
        return float(ms.item())

    for _ in range(args.warmup):
        y = step_device()
    if args.warmup and not bool(torch.isfinite(y).all()):
        raise SystemExit("bench: non-finite windows from the synthetic weights")
    if unet.range_overflow():
        raise SystemExit("bench: an activation left the f16x3 operand range (the result would be invalid)")
    shard_check = None
    if world > 1:
        # multi-GPU correctness, checked by the bench itself: the gathered tensor holds every rank's rows bit for bit, in rank
        # order (each rank compares its own rows, and rank 0's rows as every rank received them are compared across ranks)
        y = step_device()
        torch.cuda.synchronize(dev)
        ok_local = bool(torch.equal(gathered[rank * B:(rank + 1) * B], y))
        digest = gathered.double().sum(dim=(1, 2)).view(world, B).sum(dim=1)        # one number per rank's block
        digests = [torch.empty_like(digest) for _ in range(world)]
        dist.all_gather(digests, digest)
        ok_same = all(bool(torch.equal(d, digests[0])) for d in digests)
        flags = torch.tensor([int(ok_local), int(ok_same)], device=dev)
        dist.all_reduce(flags, op=dist.ReduceOp.MIN)
        shard_check = {"gathered_rows_equal_local": bool(flags[0].item()), "all_ranks_hold_same_gathered": bool(flags[1].item()),
                       "distinct_noise_per_rank": True}
        if not (flags[0].item() and flags[1].item()):
            raise SystemExit("bench: all-gathered windows differ from the ranks' local results")
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    # end-to-end leg first (host buffers, H2D + D2H inside the timed region), then the device-resident leg: the board runs at
    # its power cap and sheds ~8 % of clock as it heats up over the first minute, so the leg timed later reads lower
    step_e2e()                                            # warm the host path (allocator pools)
    ms_e2e = timed(step_e2e, args.steps)
    e2e = world * B * args.steps / (ms_e2e / 1e3)
    l0 = eegldm.launch_count()
    ms = timed(step_device, args.steps)
    launches = eegldm.launch_count() - l0
    clk = clocks.stop() if rank == 0 else None
    value = world * B * args.steps / (ms / 1e3)

    # ---- roofline leg: per-kernel-family device time measured live with CUDA events on the launch stream
    roof, prof = None, None
    if rank == 0:
        L = eegldm.lib()
        pb = min(B, args.profile_batch)
        psched = eegldm.DDIMScheduler(**synthetic.DDIM_CFG)
        psteps = 2
        psched.set_timesteps(psteps)
        L.eegldm_set_graphs(0)
        L.eegldm_profile_enable(1)
        eegldm.ddim_sample(unet, psched, noise[:pb], psteps, aekl)
        torch.cuda.synchronize(dev)
        fam = _profile_families(eegldm, _lib, C)
        L.eegldm_profile_enable(0)
        L.eegldm_set_graphs(1)
        tot = sum(v["ms"] for v in fam.values()) or 1.0
        conv = fam["conv"]
        peaks = _peaks()
        peak_tc = peaks["tc_sustained"]
        ach = conv["flops"] / (conv["ms"] / 1e3) / 1e12 if conv["ms"] else 0.0
        traffic, traffic_n, traffic_src = _ncu_conv_traffic(pb, args.math)
        mma = 3 if args.math == "f16x3" else 1
        roof = {"bound": "tensor", "kernel": "conv_tc_kernel (tcgen05 implicit-GEMM conv, %s): %d launches per UNet forward" % (
                    args.math, conv["launches"] // psteps),
                "achieved": ach, "peak": peak_tc, "unit": "TFLOP/s", "frac": ach / peak_tc,
                "peak_source": peaks["src"] + " bf16 sustained (MEASURED_PEAKS.json)",
                "traffic": traffic, "traffic_source": traffic_src, "traffic_launches_averaged": traffic_n,
                "algorithmic_bytes_per_launch": conv["bytes"] / max(conv["launches"], 1),
                "traffic_over_algorithmic": (traffic / (conv["bytes"] / max(conv["launches"], 1))) if traffic else None,
                "algorithmic_flops_per_launch": conv["flops"] / max(conv["launches"], 1),
                "avg_launch_ms": conv["ms"] / max(conv["launches"], 1), "share_of_step": conv["ms"] / tot,
                # f16x3 issues 3 fp16 MMAs per algorithmic MAC (hi*hi, hi*lo, lo*hi): the fraction of the 16-bit tensor peak the
                # kernel actually sustains is 3x `frac`; 1/3 is the ceiling of `frac` for an fp32-parity path on this pipe
                "products_per_mac": mma, "tensor_pipe_frac": ach * mma / peak_tc, "profile_batch": pb,
                "launch_set": "profile family 'conv' = the tcgen05 conv launches only (narrow fp32 convs: family 'conv_narrow'); "
                              "flops, bytes, time and the ncu traffic are all averaged over this one set",
                # whole-step view against both ceilings (13.90 GFLOP, 25.2 MB per forward per sample)
                # cross-check against the timed (graph-replayed) region: the profiled leg launches eagerly with events around every
                # kernel, so its clock state is not the timed legs'; `profile_over_timed_step` = profiled kernel time per denoise step /
                # timed time per denoise step, and `achieved_scaled_to_timed_step` is `achieved` at the timed region's speed
                "profile_over_timed_step": (tot / psteps) / (ms / args.steps / DDIM_STEPS),
                "achieved_scaled_to_timed_step": ach * (tot / psteps) / (ms / args.steps / DDIM_STEPS),
                "step_tensor_frac": value / world * DDIM_STEPS * UNET_GFLOP_PER_FWD * 1e9 / (peak_tc * 1e12),
                "step_hbm_frac": value / world * DDIM_STEPS * UNET_MB_PER_FWD * 1e6 / (peaks["hbm"] * 1e9)}
        prof = {k: {"ms": round(v["ms"], 3), "share": round(v["ms"] / tot, 4), "launches": v["launches"],
                    "tflops": (v["flops"] / (v["ms"] / 1e3) / 1e12 if v["ms"] else 0.0),
                    "gbs": (v["bytes"] / (v["ms"] / 1e3) / 1e9 if v["ms"] else 0.0)} for k, v in fam.items()}

    cpu, configs, fast = None, None, None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        vals, threads = _cpu_ddim_windows_per_s(args.ref_batch, 1, 0)
        cpu = {"value": vals[0], "unit": "windows/s", "cores": threads, "kind": "port", "batch": args.ref_batch,
               "sample": f"{args.ref_batch} windows x DDIM-50 + decode, oracle port (reference unet.py restated), torch {threads} threads"}
    if rank == 0 and world == 1 and not args.no_configs:
        # the other BASELINE.json configurations, measured in the same process (headline numbers stay config 3 / 4 above)
        configs = {"config2_unet_step_b256": measure_config2(args, dev, unet),
                   "config1_aekl_encode_decode_b4": measure_config1(args, dev),
                   "config5_aekl_train_step_b512": measure_train(args, dev),
                   "ldm_train_step_b64": measure_ldm_train(args, dev),
                   "dm_variant_unet_step_b64": measure_dm_variant(args, dev, unet)}
        # fast mode: ONE bf16 product per MAC -- reported apart, never as parity: its error against the parity mode is below
        y16 = eegldm.ddim_sample(unet, sched, noise, DDIM_STEPS, aekl)
        ub = eegldm.UNetModel(**synthetic.LDM_UNET_CFG, math="bf16")
        ub.load_state_dict(usd)
        ub = ub.to(dev).eval()
        ms_b = _events_ms(torch, lambda: eegldm.ddim_sample(ub, sched, noise, DDIM_STEPS, aekl), 2, 1)
        yb = eegldm.ddim_sample(ub, sched, noise, DDIM_STEPS, aekl)
        err = (yb - y16).abs()
        fast = {"math": "bf16 (single product, not a parity mode)", "value": B / (ms_b / 1e3), "unit": "windows/s", "ms_per_step": ms_b,
                "max_abs_error_vs_f16x3": float(err.max()), "mean_abs_error_vs_f16x3": float(err.mean()),
                "output_abs_mean": float(y16.abs().mean()), "passes_parity_tolerance": bool(torch.allclose(yb, y16, rtol=1e-3, atol=1e-4))}
        del ub

    if rank == 0:
        line = {
            "metric": "EEG windows/sec DDIM-50 sampling", "value": value, "unit": "windows/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "config %d: DDIM-50 sampling, batch %d per GPU of [1,768] latents, config_ldm.yaml UNet "
                                   "(30.5M params), AEKL 2-2-4 decode -> [B,1,3072]" % (3 if world == 1 else 4, B),
                       "math": args.math, "batch_per_gpu": B, "ddim_steps": DDIM_STEPS, "graph_lanes": args.lanes, "fuse_epilogues": args.fuse,
                       "parallelism": f"batch-shard x{world}, one all-gather of decoded windows" if world > 1 else "single GPU",
                       "l2": "activations per launch (>= 400 MB at B=1024) exceed the 126 MB L2; no explicit flush"},
            "e2e": {"value": e2e, "unit": "windows/s", "h2d_bytes_per_step": world * noise_host.numel() * 4,
                    "d2h_bytes_per_step": world * out_host.numel() * 4, "ms_per_step": ms_e2e / args.steps,
                    "path": "eegldm_ddim_sample_host (C ABI, pinned host buffers)" if world == 1 else
                            "per rank: H2D noise, eegldm_ddim_sample, NCCL all-gather, D2H of the rank's windows"},
            "gpu_launches": int(launches), "clocks": clk, "roofline": roof, "kernel_profile": prof, "cpu_baseline": cpu,
        }
        if shard_check is not None:
            line["shard_check"] = shard_check
        if configs is not None:
            line["configs"] = configs
        if fast is not None:
            line["fast_mode"] = fast
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_train(args):
    """--workload train: config 5 alone, as its own bench line; --workload ldm_train: the latent-diffusion training step."""
    import torch
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    if args.workload == "ldm_train":
        r = measure_ldm_train(args, dev, args.batch if args.batch != 1024 else 64, with_cpu=not args.no_cpu_baseline)
        r["adversarial"] = False
    else:
        r = measure_train(args, dev, args.batch if args.batch != 1024 else 512)
    line = {"metric": r["metric"], "value": r["value"], "unit": r["unit"], "n_gpus": 1, "steps": max(args.steps, 10),
            "warmup": max(args.warmup, 3), "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": {"workload": r["workload"], "adversarial": r["adversarial"]},
            "e2e": r["e2e"], "gpu_launches": r["gpu_launches_per_step"] * max(args.steps, 10), "roofline": r["roofline"],
            "cpu_baseline": r.get("cpu_baseline")}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=1024, help="windows per GPU per step (config 3: 1024)")
    ap.add_argument("--math", default=os.environ.get("EEGLDM_BENCH_MATH", "f16x3"), help="fp32 | f16x3 (both parity-green)")
    ap.add_argument("--lanes", type=int, default=1, help="independent batch halves inside the denoise-step graph (1 or 2)")
    ap.add_argument("--fuse", type=int, default=15, help="eegldm_set_conv_tuning fuse_epilogues bit mask (1 GroupNorm statistics, "
                    "2 qkv operand images, 4 in-kernel activation producer, 8 attention -> proj_out operand image, 16 attention splits fp32 q,k,v itself)")
    ap.add_argument("--pair", type=int, default=1, help="eegldm_set_conv_tuning CTA-pair mask (bit 0: 256-wide conv launches, bit 1: 128-wide)")
    ap.add_argument("--ref-batch", type=int, default=8, help="windows per CPU-baseline step (bounded sample)")
    ap.add_argument("--profile-batch", type=int, default=1024)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the configs / fast_mode objects (configs 1, 2, 5 and the bf16 line)")
    ap.add_argument("--no-adversarial", action="store_true", help="config 5 without the PatchDiscriminator term")
    ap.add_argument("--workload", default="sample", choices=["sample", "train", "ldm_train"],
                    help="sample = config 3/4 (headline); train = config 5 (AEKL training step); ldm_train = UNet training step (8f-2)")
    args = ap.parse_args()
    if args.workload in ("train", "ldm_train"):
        run_train(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
