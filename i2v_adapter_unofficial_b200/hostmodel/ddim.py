"""DDIM scheduler with the configuration the reference pipeline uses
(``/root/reference/src/pipelines/pipeline_i2v_adapter.py:755-757``: SD1.5 scheduler config + ``clip_sample=False``,
``timestep_spacing="linspace"``).  Scheduler arithmetic is outside the hot path and stays on
PyTorch.
"""
from __future__ import annotations

import numpy as np
import torch


class _StepOutput:
    def __init__(self, prev_sample, pred_original_sample):
        self.prev_sample = prev_sample
        self.pred_original_sample = pred_original_sample


class DDIMScheduler:
    order = 1
    init_noise_sigma = 1.0

    def __init__(self, num_train_timesteps: int = 1000, beta_start: float = 0.00085, beta_end: float = 0.012,
                 beta_schedule: str = "scaled_linear", clip_sample: bool = False, set_alpha_to_one: bool = False,
                 steps_offset: int = 1, prediction_type: str = "epsilon", timestep_spacing: str = "linspace"):
        if beta_schedule != "scaled_linear" or prediction_type != "epsilon" or timestep_spacing != "linspace":
            raise NotImplementedError("only the SD1.5 configuration used by the reference pipeline is provided")
        self.num_train_timesteps = num_train_timesteps
        betas = torch.linspace(beta_start**0.5, beta_end**0.5, num_train_timesteps, dtype=torch.float32) ** 2
        self.alphas_cumprod = torch.cumprod(1.0 - betas, dim=0)
        self.final_alpha_cumprod = torch.tensor(1.0) if set_alpha_to_one else self.alphas_cumprod[0]
        self.clip_sample = clip_sample
        self.num_inference_steps = None
        self.timesteps = torch.from_numpy(np.arange(0, num_train_timesteps)[::-1].copy().astype(np.int64))

    def set_timesteps(self, num_inference_steps: int, device=None) -> None:
        self.num_inference_steps = num_inference_steps
        ts = np.linspace(0, self.num_train_timesteps - 1, num_inference_steps).round()[::-1].copy().astype(np.int64)
        self.timesteps = torch.from_numpy(ts).to(device)

    def scale_model_input(self, sample, timestep=None):
        return sample

    def add_noise(self, original_samples, noise, timesteps):
        ac = self.alphas_cumprod.to(device=original_samples.device, dtype=original_samples.dtype)
        timesteps = timesteps.to(original_samples.device)
        a = ac[timesteps] ** 0.5
        s = (1 - ac[timesteps]) ** 0.5
        while a.dim() < original_samples.dim():
            a, s = a.unsqueeze(-1), s.unsqueeze(-1)
        return a * original_samples + s * noise

    def step(self, model_output, timestep, sample, eta: float = 0.0, **_unused) -> _StepOutput:
        if eta != 0.0:
            raise NotImplementedError("the reference pipeline samples with eta = 0")
        t = int(timestep)
        prev_t = t - self.num_train_timesteps // self.num_inference_steps
        a_t = self.alphas_cumprod[t]
        a_prev = self.alphas_cumprod[prev_t] if prev_t >= 0 else self.final_alpha_cumprod
        a_t, a_prev = float(a_t), float(a_prev)
        x0 = (sample - (1 - a_t) ** 0.5 * model_output) / a_t**0.5
        if self.clip_sample:
            x0 = x0.clamp(-1, 1)
        prev = a_prev**0.5 * x0 + (1 - a_prev) ** 0.5 * model_output
        return _StepOutput(prev, x0)
