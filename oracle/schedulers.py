"""CPU restatement of the DDIM / DDPM schedulers (TEST INFRASTRUCTURE).

Upstream: ``monai-generative`` (PyPI, GitHub Project-MONAI/GenerativeModels),
``generative/networks/schedulers/{scheduler,ddim,ddpm}.py``; version unpinned in the
reference (``requirements.txt:12``), 0.2.x API implied by the call sites
``src/sample_trials.py:136-145,163`` (``schedule="scaled_linear_beta"``, 2-tuple
``step`` return).  The package is not installable here -> **parity unpinned**
against upstream; the beta / cumulative-alpha tables and ``add_noise`` are pinned against the
in-tree ancestor ``src/models/ldm.py`` (tests/golden/make_golden_sched.py), the DDIM step by the
closed-form known answers of SURVEY.md section 4 (tests/test_oracle_misc.py).

Call sites this follows: ``src/sample_trials.py:136-166`` (DDIM, v-prediction,
scaled-linear), ``src/train_ldm.py:199-202`` + ``src/training/training.py:420-437``
(DDPM add_noise / get_velocity, linear schedule, epsilon target).
"""
from __future__ import annotations

import numpy as np
import torch


def make_betas(schedule: str, num_train_timesteps: int, beta_start: float, beta_end: float) -> torch.Tensor:
    if schedule in ("linear_beta", "linear"):
        return torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=torch.float32)
    if schedule in ("scaled_linear_beta", "scaled_linear"):
        return torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
    raise ValueError(f"unsupported schedule {schedule!r}")


class _Base:
    def __init__(self, num_train_timesteps=1000, beta_start=1e-4, beta_end=2e-2, schedule="linear_beta",
                 prediction_type="epsilon"):
        self.num_train_timesteps = num_train_timesteps
        self.prediction_type = prediction_type
        self.betas = make_betas(schedule, num_train_timesteps, beta_start, beta_end)
        self.alphas = 1.0 - self.betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.one = torch.tensor(1.0)
        self.timesteps = torch.from_numpy(np.arange(0, num_train_timesteps)[::-1].astype(np.int64))

    def add_noise(self, original_samples, noise, timesteps):
        """x_t = sqrt(abar_t) x_0 + sqrt(1-abar_t) eps   (training.py:429)."""
        a = self.alphas_cumprod[timesteps] ** 0.5
        s = (1 - self.alphas_cumprod[timesteps]) ** 0.5
        while a.dim() < original_samples.dim():
            a, s = a.unsqueeze(-1), s.unsqueeze(-1)
        return a * original_samples + s * noise

    def get_velocity(self, sample, noise, timesteps):
        """v = sqrt(abar_t) eps - sqrt(1-abar_t) x_0   (training.py:432-434)."""
        a = self.alphas_cumprod[timesteps] ** 0.5
        s = (1 - self.alphas_cumprod[timesteps]) ** 0.5
        while a.dim() < sample.dim():
            a, s = a.unsqueeze(-1), s.unsqueeze(-1)
        return a * noise - s * sample


class DDPMScheduler(_Base):
    pass


class DDIMScheduler(_Base):
    """eta = 0 deterministic DDIM (the reference never passes eta)."""

    def __init__(self, num_train_timesteps=1000, beta_start=1e-4, beta_end=2e-2, schedule="linear_beta",
                 prediction_type="epsilon", clip_sample=True, set_alpha_to_one=True, steps_offset=0):
        super().__init__(num_train_timesteps, beta_start, beta_end, schedule, prediction_type)
        self.clip_sample = clip_sample
        self.steps_offset = steps_offset
        self.final_alpha_cumprod = torch.tensor(1.0) if set_alpha_to_one else self.alphas_cumprod[0]
        self.num_inference_steps = None

    def set_timesteps(self, num_inference_steps: int):
        if num_inference_steps > self.num_train_timesteps:
            raise ValueError("num_inference_steps > num_train_timesteps")
        self.num_inference_steps = num_inference_steps
        step_ratio = self.num_train_timesteps // num_inference_steps
        ts = (np.arange(0, num_inference_steps) * step_ratio).round()[::-1].copy().astype(np.int64)
        self.timesteps = torch.from_numpy(ts) + self.steps_offset

    def step(self, model_output, timestep: int, sample):
        timestep = int(timestep)
        prev = timestep - self.num_train_timesteps // self.num_inference_steps
        a_t = self.alphas_cumprod[timestep]
        a_prev = self.alphas_cumprod[prev] if prev >= 0 else self.final_alpha_cumprod
        b_t = 1 - a_t
        if self.prediction_type == "epsilon":
            x0 = (sample - b_t ** 0.5 * model_output) / a_t ** 0.5
            eps = model_output
        elif self.prediction_type == "sample":
            x0 = model_output
            eps = (sample - a_t ** 0.5 * x0) / b_t ** 0.5
        elif self.prediction_type == "v_prediction":
            x0 = (a_t ** 0.5) * sample - (b_t ** 0.5) * model_output
            eps = (a_t ** 0.5) * model_output + (b_t ** 0.5) * sample
        else:
            raise ValueError(self.prediction_type)
        if self.clip_sample:
            x0 = torch.clamp(x0, -1, 1)
        direction = (1 - a_prev) ** 0.5 * eps          # eta = 0 -> std_dev_t = 0
        return a_prev ** 0.5 * x0 + direction, x0

    def step_coefficients(self, timestep: int):
        """(c_x, c_m) with x_prev = c_x * x + c_m * model_output (clip_sample=False only)."""
        timestep = int(timestep)
        prev = timestep - self.num_train_timesteps // self.num_inference_steps
        a_t = float(self.alphas_cumprod[timestep])
        a_p = float(self.alphas_cumprod[prev]) if prev >= 0 else float(self.final_alpha_cumprod)
        sa, sb, sap, sbp = a_t ** 0.5, (1 - a_t) ** 0.5, a_p ** 0.5, (1 - a_p) ** 0.5
        if self.prediction_type == "v_prediction":
            return sap * sa + sbp * sb, -sap * sb + sbp * sa
        if self.prediction_type == "epsilon":
            return sap / sa, -sap * sb / sa + sbp
        raise ValueError(self.prediction_type)
