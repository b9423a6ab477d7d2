"""Property tests (hypothesis) of the host-side logic that the GPU path relies on: the batch-axis partition used by
multi-GPU sampling, and the two-scalar DDIM update that the UNet's output epilogue applies (x <- c1*x + c2*model_out)."""
import sys

import numpy as np
import torch
from hypothesis import given, settings, strategies as st

from conftest import PKG

if PKG not in sys.path:
    sys.path.insert(0, PKG)

from oracle.schedulers import DDIMScheduler  # noqa: E402


def _shard_range(n, rank, world):
    from eegldm.sampler import shard_range
    return shard_range(n, rank, world)


@settings(max_examples=200, deadline=None, derandomize=True)
@given(n=st.integers(0, 20000), world=st.integers(1, 16))
def test_shard_range_is_a_contiguous_balanced_partition(n, world):
    parts = [_shard_range(n, r, world) for r in range(world)]
    assert parts[0][0] == 0 and parts[-1][1] == n
    for (l0, h0), (l1, h1) in zip(parts, parts[1:]):
        assert h0 == l1 and l0 <= h0                       # contiguous, in rank order (the all-gather's row order)
    sizes = [h - l for l, h in parts]
    assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)   # balanced, the extra rows on the first ranks


@settings(max_examples=60, deadline=None, derandomize=True)
@given(n_steps=st.sampled_from([1, 2, 4, 10, 20, 25, 50, 100, 200, 250, 1000]),
       pred=st.sampled_from(["v_prediction", "epsilon"]),
       schedule=st.sampled_from(["scaled_linear_beta", "linear_beta"]),
       seed=st.integers(0, 2 ** 16))
def test_ddim_step_is_the_two_scalar_update_the_epilogue_applies(n_steps, pred, schedule, seed):
    """eegldm_sched_ddim_tables hands the kernel (c1, c2) per step; here the same closed form is checked against the
    restated scheduler's step() for every step of random schedules: eta = 0, clip_sample=False, set_alpha_to_one."""
    s = DDIMScheduler(num_train_timesteps=1000, beta_start=0.0015, beta_end=0.0205, schedule=schedule, prediction_type=pred,
                      clip_sample=False)
    s.set_timesteps(n_steps)
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(2, 1, 16, generator=g)
    out = torch.randn(2, 1, 16, generator=g)
    ac = s.alphas_cumprod.double()
    ratio = 1000 // n_steps
    for t in s.timesteps.tolist()[:: max(1, n_steps // 7)]:
        a_t = ac[t]
        prev = t - ratio
        a_p = ac[prev] if prev >= 0 else torch.tensor(1.0, dtype=torch.float64)
        if pred == "v_prediction":      # x0 = sqrt(a) x - sqrt(1-a) v ; eps = sqrt(a) v + sqrt(1-a) x
            c1 = a_p.sqrt() * a_t.sqrt() + (1 - a_p).sqrt() * (1 - a_t).sqrt()
            c2 = -a_p.sqrt() * (1 - a_t).sqrt() + (1 - a_p).sqrt() * a_t.sqrt()
        else:                           # x0 = (x - sqrt(1-a) eps) / sqrt(a)
            c1 = a_p.sqrt() / a_t.sqrt()
            c2 = (1 - a_p).sqrt() - a_p.sqrt() * (1 - a_t).sqrt() / a_t.sqrt()
        got, _ = s.step(out, t, x)
        want = (c1 * x.double() + c2 * out.double()).float()
        np.testing.assert_allclose(got.numpy(), want.numpy(), rtol=2e-4, atol=2e-5)
