"""PatchDiscriminator + PatchAdversarialLoss inside the autoencoder training step
(src/train_autoencoderkl.py:204-234; config/config_aekl_eeg.yaml:30-40), CUDA path through the C ABI vs torch autograd over the
CPU oracle (oracle/discriminator.py, oracle/aekl.py, oracle/jukebox.py).  Tolerances as in test_gpu_train.py."""
import pytest
import torch

from oracle import aekl as oa
from oracle import discriminator as od
from oracle import jukebox as oj

pytestmark = pytest.mark.gpu


def _close(a, b, rtol, floor):
    scale = float(b.abs().max())
    torch.testing.assert_close(a, b, rtol=rtol, atol=floor * max(scale, 1e-30))


def _disc(cfg, sd, dev, math="f16x3"):
    import eegldm
    d = eegldm.PatchDiscriminator(**cfg, math=math)
    d.load_state_dict(sd)
    return d.to(dev)


@pytest.mark.parametrize("math", ["fp32", "f16x3"])
@pytest.mark.parametrize("over,B,L", [({}, 3, 3072), (dict(num_channels=8, num_layers_d=2), 5, 512), (dict(num_channels=16, num_layers_d=4), 2, 1024)])
def test_discriminator_forward_matches_oracle(built_lib, cuda_device, over, B, L, math):
    """The reference configuration takes the tcgen05 path in f16x3 (64 -> 128 -> 256 -> 512 convs; stride 2 as row pairs);
    the narrow configurations fall back to the fp32 kernels layer by layer."""
    cfg = od.full_cfg(**over)
    sd = od.make_disc_state_dict(cfg, 7, weight_std=0.1)
    x = torch.rand(B, 1, L, generator=torch.Generator().manual_seed(0))
    d = _disc(cfg, sd, cuda_device, math)
    run = {}
    ref = od.forward(cfg, sd, x, training=True, update_running=run)[-1]
    y = d(x.to(cuda_device))[-1]
    assert y.shape == ref.shape
    _close(y.cpu(), ref, 1e-3, 1e-4)
    d.sync_trained()                                   # running statistics after ONE training-mode forward
    new = d.state_dict()
    for k, v in run.items():
        _close(new[k].cpu(), v, 1e-4, 1e-5)
    assert int(new["0.adn.N.num_batches_tracked"]) == 1
    # eval(): running statistics
    sd2 = dict(sd)
    sd2.update(run)
    ref_eval = od.forward(cfg, sd2, x, training=False)[-1]
    y_eval = d.eval()(x.to(cuda_device))[-1]
    _close(y_eval.cpu(), ref_eval, 1e-3, 1e-4)


def _oracle_full_step(acfg, asd, dcfg, dsd, x, eps, kl_w, spec_w, adv_w, lr_g, lr_d):
    """train_autoencoderkl.py:204-234 with torch autograd."""
    gp = {k: v.clone().requires_grad_(True) for k, v in asd.items()}
    dp = {k: (v.clone().requires_grad_(True) if not od.is_buffer(k) else v.clone()) for k, v in dsd.items()}
    opt_g = torch.optim.Adam(list(gp.values()), lr=lr_g)
    opt_d = torch.optim.Adam([v for k, v in dp.items() if not od.is_buffer(k)], lr=lr_d)
    run = {}
    recon, mu, sigma = oa.forward(acfg, gp, x, eps)
    l1 = torch.nn.functional.l1_loss(recon, x)
    spec = oj.jukebox_loss(recon, x, reduction="sum")
    kl = oa.kl_loss(mu, sigma)
    gen = od.patch_adversarial_loss(od.forward(dcfg, dp, recon.contiguous(), True, run)[-1], True, False)
    dp.update(run)
    loss_g = l1 + kl_w * kl + adv_w * gen + spec_w * spec
    loss_g.backward()
    g_grads = {k: p.grad.clone() for k, p in gp.items()}
    opt_g.step()
    opt_d.zero_grad(set_to_none=True)
    d_fake = od.patch_adversarial_loss(od.forward(dcfg, dp, recon.contiguous().detach(), True, run)[-1], False, True)
    dp.update(run)
    d_real = od.patch_adversarial_loss(od.forward(dcfg, dp, x.contiguous().detach(), True, run)[-1], True, True)
    dp.update(run)
    disc = 0.5 * (d_fake + d_real)
    (adv_w * disc).backward()
    d_grads = {k: p.grad.clone() for k, p in dp.items() if not od.is_buffer(k)}
    opt_d.step()
    losses = dict(l1=float(l1), kl=float(kl), spectral=float(spec), total=float(loss_g), generator=float(gen), discriminator=float(disc))
    return losses, g_grads, {k: p.detach() for k, p in gp.items()}, d_grads, {k: p.detach() for k, p in dp.items()}


@pytest.mark.parametrize("math", ["fp32", "f16x3"])
@pytest.mark.parametrize("nc,dover,B,L", [([2, 2, 4], {}, 3, 3072), ([4, 4], dict(num_channels=8, num_layers_d=2), 4, 512)])
def test_full_training_step_matches_autograd(built_lib, cuda_device, nc, dover, B, L, math):
    import eegldm
    acfg = oa.full_cfg(num_channels=nc, attention_levels=[False] * len(nc))
    asd = oa.make_aekl_state_dict(acfg, 42)
    dcfg = od.full_cfg(**dover)
    dsd = od.make_disc_state_dict(dcfg, 7, weight_std=0.1)
    g = torch.Generator().manual_seed(B)
    x = torch.rand(B, 1, L, generator=g)
    eps = torch.randn(B, 1, L // 2 ** (len(nc) - 1), generator=g)
    kw = dict(kl_w=1e-6, spec_w=1e-3, adv_w=0.5, lr_g=5e-3, lr_d=5e-4)
    ref, gg, gnew, dg, dnew = _oracle_full_step(acfg, asd, dcfg, dsd, x, eps, **kw)
    # The step is badly conditioned in fp32 (BatchNorm over a tiny batch, LeakyReLU kinks, the gradient reaches the encoder
    # through four normalised stages): the fp32 oracle itself is ~1 % (relative to each tensor's largest entry) away from the
    # same oracle in fp64.  Ground truth = fp64; the CUDA path must be as close to it as the fp32 reference is.
    to64 = lambda d_: {k: (v.double() if v.is_floating_point() else v) for k, v in d_.items()}
    _, gg64, gnew64, dg64, dnew64 = _oracle_full_step(acfg, to64(asd), dcfg, to64(dsd), x.double(), eps.double(), **kw)

    def as_good_as_fp32(name, got, ref32, ref64, floor):
        scale = max(float(ref64.abs().max()), 1e-30)
        err = float((got.double() - ref64).abs().max()) / scale
        err_ref = float((ref32.double() - ref64).abs().max()) / scale
        assert err <= max(floor, 3.0 * err_ref), (name, err, err_ref)

    m = eegldm.AutoencoderKL(**acfg)
    m.load_state_dict(asd)
    m = m.to(cuda_device)
    d = _disc(dcfg, dsd, cuda_device, math)
    out = m.train_step(x.to(cuda_device), eps.to(cuda_device), kl_weight=kw["kl_w"], spectral_weight=kw["spec_w"], lr=kw["lr_g"],
                       discriminator=d, adv_weight=kw["adv_w"], lr_d=kw["lr_d"])
    for k in ref:
        assert out[k] == pytest.approx(ref[k], rel=2e-4, abs=1e-7), k
    grads = m.grad_dict()
    for k, v in gg.items():
        as_good_as_fp32(k, grads[k], v, gg64[k], 2e-3)     # generator gradients carry the adversarial term through D
    dgr = d.grad_dict()
    assert set(dgr) == set(dg)
    for k, v in dg.items():
        as_good_as_fp32(k, dgr[k], v, dg64[k], 2e-3)
    m.sync_trained()
    d.sync_trained()
    for k, v in gnew.items():
        as_good_as_fp32(k, m.state_dict()[k].cpu(), v, gnew64[k], 2e-3)
    new = d.state_dict()
    for k, v in dnew.items():
        if k.endswith("num_batches_tracked"):
            assert int(new[k]) == 3                # three training-mode forwards per step (fake, fake again, real)
        else:
            as_good_as_fp32(k, new[k].cpu().float(), v.float(), dnew64[k], 2e-3)
    # a second step continues from the updated state (Adam moments, running statistics) without error and lowers nothing to NaN
    out2 = m.train_step(x.to(cuda_device), eps.to(cuda_device), kl_weight=kw["kl_w"], spectral_weight=kw["spec_w"], lr=kw["lr_g"],
                        discriminator=d, adv_weight=kw["adv_w"], lr_d=kw["lr_d"])
    assert all(v == v for v in out2.values())


def test_patch_adversarial_loss_module(built_lib, cuda_device):
    import eegldm
    loss = eegldm.PatchAdversarialLoss(criterion="least_squares")
    z = torch.randn(4, 1, 384, generator=torch.Generator().manual_seed(0))
    for real, for_d in ((True, False), (False, True), (True, True), (False, False)):
        got = loss(z.to(cuda_device), target_is_real=real, for_discriminator=for_d)
        torch.testing.assert_close(got.cpu(), od.patch_adversarial_loss(z, real, for_d))
    with pytest.raises(eegldm.EegldmError):
        eegldm.PatchDiscriminator(spatial_dims=1, num_layers_d=3, num_channels=64, in_channels=1, out_channels=1)   # upstream default k=4
