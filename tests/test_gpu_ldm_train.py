"""the latent-diffusion training step (src/training/training.py:420-443) through the C ABI
(eegldm_unet_train_step) against
  * the CPU oracle with torch autograd (oracle/ldm_train.py) on the same seeded weights / latents / noise / timesteps: loss, every
    parameter's gradient, every parameter after one Adam step;
  * golden gradients produced by autograd through the REFERENCE's own UNetModel (tests/golden/ldm_train_golden.npz), including
    the full config_ldm.yaml network (gradient digests).
Tolerances: loss rtol 1e-4; gradients rtol 2e-3 with an absolute floor of 1e-4 x the tensor's largest entry + 1e-5 x the model's
largest gradient entry (fp32 atomics and split-K reductions reorder sums; f16x3 operands carry 2^-22 relative error)."""
import pytest
import torch

from oracle import ldm_train as ol

from test_oracle_ldm_train import _G, check_against_golden, golden_case
from train_cases import FULL_GRADS, TRAIN_CASES

pytestmark = pytest.mark.gpu


def _model(cfg, sd, dev, math):
    import eegldm
    m = eegldm.UNetModel(**cfg, math=math)
    m.load_state_dict(sd)
    return m.to(dev)


def _sched(name):
    import eegldm
    _, _, _, pred, schedule, (b0, b1) = TRAIN_CASES[name]
    return eegldm.DDPMScheduler(1000, b0, b1, schedule, pred)


def _close_all(got, ref, rtol, floor, gfloor):
    gmax = max(float(v.abs().max()) for v in ref.values())
    for k, r in ref.items():
        torch.testing.assert_close(got[k], r, rtol=rtol, atol=floor * float(r.abs().max()) + gfloor * gmax, msg=lambda m: f"{k}: {m}")


@pytest.mark.parametrize("math", ["fp32", "f16x3"])
@pytest.mark.parametrize("name", [n for n in TRAIN_CASES if n != "ldm_full"])
def test_training_step_matches_oracle_autograd(built_lib, cuda_device, name, math):
    torch.set_num_threads(8)
    cfg, sd, sched, z0, noise, t, loss_ref = golden_case(name)
    loss_o, grads_o, new_o = ol.ldm_train_step(cfg, sd, z0, noise, t, sched, lr=1e-4)
    m = _model(cfg, sd, cuda_device, math)
    loss = m.train_step(z0.to(cuda_device), noise.to(cuda_device), t.to(cuda_device), _sched(name), lr=1e-4)
    assert abs(loss - loss_o) <= 1e-4 * abs(loss_o)
    grads = m.grad_dict()
    _close_all(grads, grads_o, 2e-3, 1e-4, 1e-5)
    check_against_golden(name, grads, 2e-3, 1e-4)
    # Adam's first step is lr * sign(g) wherever |g| >> eps: compare where the oracle gradient is not rounding noise
    new = m._export(0)
    gmax = max(float(v.abs().max()) for v in grads_o.values())
    for k, r in new_o.items():
        solid = grads_o[k].abs() > 1e-4 * gmax
        torch.testing.assert_close(new[k][solid], r[solid], rtol=0, atol=2e-6, msg=lambda msg: f"{k}: {msg}")
        assert float((new[k] - sd[k]).abs().max()) <= 1.01e-4
    # the trained weights reach the inference path
    m.sync_trained()
    x = torch.randn(2, cfg["in_channels"], z0.shape[-1], generator=torch.Generator().manual_seed(5))
    from oracle import unet as ou
    y = m(x.to(cuda_device), timesteps=torch.tensor([321])).cpu()
    torch.testing.assert_close(y, ou.unet_forward(cfg, new_o, x, torch.tensor([321])), rtol=2e-3, atol=2e-3)


@pytest.mark.parametrize("math", ["f16x3"])
def test_full_network_gradients_match_reference_golden(built_lib, cuda_device, math):
    """config_ldm.yaml (30.5 M parameters, B = 2 x [1, 768]): tensor-pipe forward, data and weight gradients."""
    cfg, sd, sched, z0, noise, t, loss_ref = golden_case("ldm_full")
    m = _model(cfg, sd, cuda_device, math)
    loss = m.train_step(z0.to(cuda_device), noise.to(cuda_device), t.to(cuda_device), _sched("ldm_full"), lr=0.0)
    assert abs(loss - loss_ref) <= 1e-4 * abs(loss_ref)
    check_against_golden("ldm_full", m.grad_dict(), 2e-3, 1e-4)
    assert not m.range_overflow()


def test_second_step_uses_updated_weights(built_lib, cuda_device):
    """two steps on the device == two oracle steps (Adam moments, tensor-pipe weight images rebuilt after the update)"""
    name = "small_eps"
    cfg, sd, sched, z0, noise, t, _ = golden_case(name)
    import torch.nn.functional as F  # noqa: F401
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    opt = torch.optim.Adam(list(params.values()), lr=1e-3)
    losses_o = []
    for _ in range(2):
        opt.zero_grad()
        l = ol.ldm_loss(cfg, params, z0, noise, t, sched)
        l.backward()
        opt.step()
        losses_o.append(float(l))
    m = _model(cfg, sd, cuda_device, "f16x3")
    losses = [m.train_step(z0.to(cuda_device), noise.to(cuda_device), t.to(cuda_device), _sched(name), lr=1e-3) for _ in range(2)]
    assert abs(losses[0] - losses_o[0]) <= 1e-4 * abs(losses_o[0])
    assert abs(losses[1] - losses_o[1]) <= 2e-3 * abs(losses_o[1])
    assert losses[1] != losses[0]
