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
