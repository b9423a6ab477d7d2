"""The reference's own training loop body, src/train_autoencoderkl.py:203-234, run UNCHANGED against the drop-in modules
(eegldm.AutoencoderKL / PatchDiscriminator / JukeboxLoss / PatchAdversarialLoss across the autograd boundary, torch.optim.Adam on
the modules' nn.Parameters) and compared with the same loop over the CPU oracle (fp64 ground truth, fp32 reference accuracy)."""
import types

import pytest
import torch
from torch.nn import L1Loss

from oracle import aekl as oa
from oracle import discriminator as od

from test_gpu_adversarial import _oracle_full_step

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("nc,dover,B,L", [([2, 2, 4], {}, 3, 3072), ([4, 4], dict(num_channels=8, num_layers_d=2), 4, 512)])
def test_reference_loop_body_verbatim(built_lib, cuda_device, nc, dover, B, L):
    import eegldm
    acfg = oa.full_cfg(num_channels=nc, attention_levels=[False] * len(nc))
    asd = oa.make_aekl_state_dict(acfg, 42)
    dcfg = od.full_cfg(**dover)
    dsd = od.make_disc_state_dict(dcfg, 7, weight_std=0.1)
    # (seed 4 puts ONE LeakyReLU pre-activation of the small discriminator within fp32 rounding of its kink: the fp32 and fp64 gradients
    # then differ by that element's whole contribution, 0.6 % of a bias gradient -- verified by perturbing the input by 1e-3)
    x = torch.rand(B, 1, L, generator=torch.Generator().manual_seed(B + 10))
    kl_weight, spectral_weight, adv_weight = 1e-6, 1e-3, 0.5
    device = cuda_device
    # AutoencoderKL.forward draws eps = randn_like(z_sigma) on the device: reproduce the draw for the oracle
    torch.cuda.manual_seed(1234)
    eps = torch.randn((B, 1, L // 2 ** (len(nc) - 1)), device=device, dtype=torch.float32).cpu()
    kw = dict(kl_w=kl_weight, spec_w=spectral_weight, adv_w=adv_weight, lr_g=5e-3, lr_d=5e-4)
    ref, gg, gnew, dg, dnew = _oracle_full_step(acfg, asd, dcfg, dsd, x, eps, **kw)
    to64 = lambda d_: {k: (v.double() if v.is_floating_point() else v) for k, v in d_.items()}
    ref64, gg64, gnew64, dg64, dnew64 = _oracle_full_step(acfg, to64(asd), dcfg, to64(dsd), x.double(), eps.double(), **kw)

    model = eegldm.AutoencoderKL(**acfg)
    model.load_state_dict(asd)
    model = model.to(device).train()
    discriminator = eegldm.PatchDiscriminator(**dcfg)
    discriminator.load_state_dict(dsd)
    discriminator = discriminator.to(device).train()
    l1_loss = L1Loss()
    adv_loss = eegldm.PatchAdversarialLoss(criterion="least_squares")
    jukebox_loss = eegldm.JukeboxLoss(spatial_dims=1, reduction="sum")
    optimizer_g = torch.optim.Adam(params=model.parameters(), lr=kw["lr_g"])
    optimizer_d = torch.optim.Adam(params=discriminator.parameters(), lr=kw["lr_d"])
    args = types.SimpleNamespace(spe="spectral")
    batch = {"eeg": x}
    torch.cuda.manual_seed(1234)

    # ---- src/train_autoencoderkl.py:201-234, verbatim -------------------------------------------------------------------------
    eeg_data = batch['eeg'].to(device)

    optimizer_g.zero_grad(set_to_none=True)
    reconstruction, z_mu, z_sigma = model(eeg_data)

    recons_loss = l1_loss(reconstruction.float(), eeg_data.float())

    recons_spectral = jukebox_loss(reconstruction.float(), eeg_data.float())

    kl_loss = 0.5 * torch.sum(z_mu.pow(2) + z_sigma.pow(2) - torch.log(z_sigma.pow(2)) - 1, dim=[1])
    kl_loss = torch.sum(kl_loss) / kl_loss.shape[0]

    logits_fake = discriminator(reconstruction.contiguous().float())[-1]
    generator_loss = adv_loss(logits_fake, target_is_real=True, for_discriminator=False)
    if args.spe == "spectral":
        loss_g = recons_loss + kl_weight * kl_loss + adv_weight * generator_loss  + recons_spectral * spectral_weight
    else:
        loss_g = recons_loss + kl_weight * kl_loss + adv_weight * generator_loss 
    loss_g.backward()
    g_grads = {k: p.grad.detach().cpu().clone() for k, p in model.named_parameters()}   # (test probe, not in the reference)
    optimizer_g.step()

    # Discriminator part
    optimizer_d.zero_grad(set_to_none=True)

    logits_fake = discriminator(reconstruction.contiguous().detach())[-1]
    loss_d_fake = adv_loss(logits_fake, target_is_real=False, for_discriminator=True)
    logits_real = discriminator(eeg_data.contiguous().detach())[-1]
    loss_d_real = adv_loss(logits_real, target_is_real=True, for_discriminator=True)
    discriminator_loss = (loss_d_fake + loss_d_real) * 0.5

    loss_d = adv_weight * discriminator_loss

    loss_d.backward()
    optimizer_d.step()
    # ---------------------------------------------------------------------------------------------------------------------------

    got = dict(l1=recons_loss.item(), kl=kl_loss.item(), spectral=recons_spectral.item(), total=loss_g.item(),
               generator=generator_loss.item(), discriminator=discriminator_loss.item())
    for k in ref:
        assert got[k] == pytest.approx(ref[k], rel=2e-4, abs=1e-7), k

    def as_good_as_fp32(name, v, ref32, r64, floor):
        scale = max(float(r64.abs().max()), 1e-30)
        err = float((v.double() - r64).abs().max()) / scale
        err_ref = float((ref32.double() - r64).abs().max()) / scale
        assert err <= max(floor, 3.0 * err_ref), (name, err, err_ref)

    for k, v in gg.items():
        as_good_as_fp32(k, g_grads[k], v, gg64[k], 2e-3)
    for k, p in discriminator.named_parameters():
        as_good_as_fp32(k, p.grad.cpu(), dg[k], dg64[k], 2e-3)
    for k, v in gnew.items():
        as_good_as_fp32(k, model.state_dict()[k].cpu(), v, gnew64[k], 2e-3)
    new = discriminator.state_dict()
    for k, v in dnew.items():
        if k.endswith("num_batches_tracked"):
            assert int(new[k]) == 3
        else:
            as_good_as_fp32(k, new[k].cpu().float(), v.float(), dnew64[k], 2e-3)
    # the updated weights reach the inference path of both modules
    with torch.no_grad():
        r2 = model.reconstruct(eeg_data)
        assert bool(torch.isfinite(r2).all())
        mu_ref, _ = oa.encode(acfg, {k: v for k, v in gnew.items()}, x)
        torch.testing.assert_close(r2.cpu(), oa.decode(acfg, gnew, mu_ref), rtol=5e-3, atol=5e-3)


@pytest.mark.parametrize("case", ["small_eps", "small_v_convresample"])
def test_reference_ldm_loop_body_verbatim(built_lib, cuda_device, case):
    """src/training/training.py:418-443 (train_epoch_ldm's loop body) unchanged: Stage1Wrapper over the drop-in autoencoder, the
    drop-in scheduler and denoiser under autocast, GradScaler, torch.optim.Adam(lr 1e-4) -- against the oracle's fp32 step on the
    same latents / noise / timesteps (GradScaler's loss scale cancels in scaler.step)."""
    import eegldm
    import torch.nn as nn
    import torch.nn.functional as F
    from collections import OrderedDict
    from torch.cuda.amp import GradScaler, autocast
    from oracle import ldm_train as ol, unet as ou
    from oracle.schedulers import DDPMScheduler as ODDPM
    from train_cases import TRAIN_CASES
    over, B, T, pred, schedule, (b0, b1) = TRAIN_CASES[case]
    ucfg = ou.full_cfg(**over)
    usd = ou.make_unet_state_dict(ucfg, seed=0)
    acfg = oa.full_cfg(num_channels=[4, 4], attention_levels=[False, False], latent_channels=ucfg["in_channels"])
    asd = oa.make_aekl_state_dict(acfg, 42)
    device = cuda_device

    class Stage1Wrapper(nn.Module):   # src/training/training.py:15-26
        def __init__(self, model: nn.Module) -> None:
            super().__init__()
            self.model = model

        def forward(self, x: torch.Tensor) -> torch.Tensor:
            z_mu, z_sigma = self.model.encode(x)
            z = self.model.sampling(z_mu, z_sigma)
            return z

    aekl = eegldm.AutoencoderKL(**acfg)
    aekl.load_state_dict(asd)
    stage1 = Stage1Wrapper(aekl.to(device).eval())
    model = eegldm.UNetModel(**ucfg)
    model.load_state_dict(usd)
    model = model.to(device)
    scheduler = eegldm.DDPMScheduler(num_train_timesteps=1000, beta_schedule=schedule, beta_start=b0, beta_end=b1, prediction_type=pred)
    optimizer = torch.optim.Adam(model.parameters(), lr=1e-4)
    scaler = GradScaler(init_scale=1024.0)
    scale_factor = 0.7
    x = {"eeg": torch.rand(B, 1, 2 * T, generator=torch.Generator().manual_seed(3))}

    # the random draws of the loop body, reproduced for the oracle: timesteps, the sampling noise of stage1, the diffusion noise
    torch.manual_seed(77)
    torch.cuda.manual_seed(77)
    t_ref = torch.randint(0, scheduler.num_train_timesteps, (B,), device=device).long()
    with torch.no_grad():
        e_ref = stage1(x["eeg"].to(device)) * scale_factor
    n_ref = torch.randn_like(e_ref)
    osched = ODDPM(1000, b0, b1, schedule, pred)
    loss_o, grads_o, new_o = ol.ldm_train_step(ucfg, usd, e_ref.cpu(), n_ref.cpu(), t_ref.cpu(), osched, lr=1e-4)
    torch.manual_seed(77)
    torch.cuda.manual_seed(77)

    # ---- src/training/training.py:415-443, verbatim ---------------------------------------------------------------------------
    model.train()

    images = x['eeg'].to(device)
    timesteps = torch.randint(0, scheduler.num_train_timesteps, (images.shape[0],), device=device).long()

    optimizer.zero_grad(set_to_none=True)
    with autocast(enabled=True):
        with torch.no_grad():
            ##### Replace
            e = stage1(images) * scale_factor

        noise = torch.randn_like(e).to(device)
        noisy_e = scheduler.add_noise(original_samples=e, noise=noise, timesteps=timesteps)
        noise_pred = model(x=noisy_e, timesteps=timesteps)

        if scheduler.prediction_type == "v_prediction":
            # Use v-prediction parameterization
            target = scheduler.get_velocity(e, noise, timesteps)
        elif scheduler.prediction_type == "epsilon":
            target = noise
        loss = F.mse_loss(noise_pred.float(), target.float())

    losses = OrderedDict(loss=loss)

    scaler.scale(losses["loss"]).backward()
    scaler.step(optimizer)
    scaler.update()
    # ---------------------------------------------------------------------------------------------------------------------------

    assert torch.equal(timesteps, t_ref) and torch.equal(noise, n_ref)
    assert abs(loss.item() - loss_o) <= 1e-4 * abs(loss_o)
    gmax = max(float(v.abs().max()) for v in grads_o.values())
    for k, p in model.named_parameters():   # p.grad was unscaled in place by scaler.step
        r = grads_o[k]
        torch.testing.assert_close(p.grad.cpu(), r, rtol=2e-3, atol=1e-4 * float(r.abs().max()) + 1e-5 * gmax, msg=lambda m: f"{k}: {m}")
        solid = r.abs() > 1e-4 * gmax
        torch.testing.assert_close(p.detach().cpu()[solid], new_o[k][solid], rtol=0, atol=2e-6, msg=lambda m: f"{k} (after Adam): {m}")
