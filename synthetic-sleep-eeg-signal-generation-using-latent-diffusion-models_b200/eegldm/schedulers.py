"""Drop-ins for ``generative.networks.schedulers.{DDIMScheduler, DDPMScheduler}`` as the reference uses
them (``src/sample_trials.py:136-145,163``; ``src/train_ldm.py:199-202``; ``src/training/training.py:420-437``).

The beta / alpha tables and the DDIM step coefficients come from the C ABI's host helpers
(``eegldm_sched_alphas_cumprod``, ``eegldm_sched_ddim_tables``) so that this class and the fused
``eegldm_ddim_sample`` loop share one definition.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib

_SCHEDULES = {"linear_beta": 0, "linear": 0, "scaled_linear_beta": 1, "scaled_linear": 1}
_PRED = {"epsilon": 0, "v_prediction": 1}


class _Scheduler:
    def __init__(self, num_train_timesteps=1000, beta_start=1e-4, beta_end=2e-2, schedule="linear_beta",
                 prediction_type="epsilon", beta_schedule=None, set_alpha_to_one=True, steps_offset=0):
        if beta_schedule is not None:  # pre-0.2.1 keyword still used by src/train_ldm.py:199
            schedule = beta_schedule
        if schedule not in _SCHEDULES:
            raise ValueError(f"unsupported schedule {schedule!r}")
        if prediction_type not in _PRED:
            raise ValueError(f"unsupported prediction_type {prediction_type!r}")
        self.num_train_timesteps = num_train_timesteps
        self.prediction_type = prediction_type
        self.schedule = schedule
        cfg = _lib.SchedCfg()
        cfg.num_train_timesteps = int(num_train_timesteps)
        cfg.beta_start, cfg.beta_end = float(beta_start), float(beta_end)
        cfg.schedule, cfg.prediction_type = _SCHEDULES[schedule], _PRED[prediction_type]
        cfg.set_alpha_to_one, cfg.steps_offset = int(bool(set_alpha_to_one)), int(steps_offset)
        self._cfg = cfg
        ac = np.empty(num_train_timesteps, dtype=np.float32)
        _lib.check(_lib.lib().eegldm_sched_alphas_cumprod(C.byref(cfg), ac.ctypes.data_as(C.POINTER(C.c_float))))
        self.alphas_cumprod = torch.from_numpy(ac)
        self.timesteps = torch.from_numpy(np.arange(0, num_train_timesteps)[::-1].astype(np.int64))
        self.one = torch.tensor(1.0)

    def to(self, device):  # Scheduler.to(), sample_trials.py:145
        self.alphas_cumprod = self.alphas_cumprod.to(device)
        self.timesteps = self.timesteps.to(device)
        return self

    def _coefs(self, timesteps, ref):
        ac = self.alphas_cumprod.to(ref.device)
        t = torch.as_tensor(timesteps, device=ref.device).long()
        a = ac[t] ** 0.5
        s = (1 - ac[t]) ** 0.5
        while a.dim() < ref.dim():
            a, s = a.unsqueeze(-1), s.unsqueeze(-1)
        return a, s

    def add_noise(self, original_samples, noise, timesteps):      # training.py:429
        a, s = self._coefs(timesteps, original_samples)
        return a * original_samples + s * noise

    def get_velocity(self, sample, noise, timesteps):             # training.py:432-434
        a, s = self._coefs(timesteps, sample)
        return a * noise - s * sample


class DDPMScheduler(_Scheduler):
    pass


class DDIMScheduler(_Scheduler):
    """Deterministic (eta = 0) DDIM, the only mode the reference uses."""

    def __init__(self, num_train_timesteps=1000, beta_start=1e-4, beta_end=2e-2, schedule="linear_beta",
                 prediction_type="epsilon", clip_sample=True, set_alpha_to_one=True, steps_offset=0, beta_schedule=None):
        super().__init__(num_train_timesteps, beta_start, beta_end, schedule, prediction_type, beta_schedule,
                         set_alpha_to_one, steps_offset)
        self.clip_sample = clip_sample
        self.num_inference_steps = None
        self._coef = None

    def set_timesteps(self, num_inference_steps: int, device=None):
        ts = np.empty(num_inference_steps, dtype=np.int64)
        coef = np.empty(2 * num_inference_steps, dtype=np.float32)
        _lib.check(_lib.lib().eegldm_sched_ddim_tables(
            C.byref(self._cfg), int(num_inference_steps), ts.ctypes.data_as(C.POINTER(C.c_int64)),
            coef.ctypes.data_as(C.POINTER(C.c_float))))
        self.num_inference_steps = num_inference_steps
        self.timesteps = torch.from_numpy(ts)
        if device is not None:
            self.timesteps = self.timesteps.to(device)
        self._coef = {int(t): (float(coef[2 * i]), float(coef[2 * i + 1])) for i, t in enumerate(ts)}

    def step_coefficients(self, timestep):
        """(c_x, c_m): x_prev = c_x * x + c_m * model_output (clip_sample=False)."""
        return self._coef[int(timestep)]

    def step(self, model_output, timestep, sample, eta: float = 0.0, generator=None):
        """-> (pred_prev_sample, pred_original_sample), sample_trials.py:163."""
        if eta != 0.0:
            raise NotImplementedError("the reference samples with eta = 0")
        t = int(timestep)
        prev = t - self.num_train_timesteps // self.num_inference_steps
        ac = self.alphas_cumprod
        a_t = float(ac[t])
        a_p = float(ac[prev]) if prev >= 0 else (1.0 if self._cfg.set_alpha_to_one else float(ac[0]))
        sa, sb = a_t ** 0.5, (1 - a_t) ** 0.5
        if self.prediction_type == "epsilon":
            x0 = (sample - sb * model_output) / sa
            eps = model_output
        else:
            x0 = sa * sample - sb * model_output
            eps = sa * model_output + sb * sample
        if self.clip_sample:
            x0 = torch.clamp(x0, -1, 1)
        return (a_p ** 0.5) * x0 + ((1 - a_p) ** 0.5) * eps, x0
