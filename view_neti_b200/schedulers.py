"""Noise schedules used on the hot path (diffusers DDPMScheduler / DDIMScheduler as configured by SD-2.1):
scaled-linear betas 0.00085 -> 0.012 over 1000 steps.  Host side only holds the alphas_cumprod table; the per-step
arithmetic of the sampler runs in the fused CFG + DDIM kernel (vn_cfg_ddim_step).

    reference training/coach.py:182-183   noisy = scheduler.add_noise(latents, noise, timesteps)
    reference training/coach.py:201-205   target = noise | scheduler.get_velocity(latents, noise, timesteps)
    reference sd_pipeline_call.py:49,101  scheduler.set_timesteps(...); scheduler.step(...).prev_sample
"""
from __future__ import annotations

from types import SimpleNamespace

import torch


def alphas_cumprod(num_train_timesteps: int = 1000, beta_start: float = 0.00085, beta_end: float = 0.012) -> torch.Tensor:
    betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
    return torch.cumprod(1.0 - betas, dim=0)


class DDPMScheduler:
    def __init__(self, prediction_type: str = "v_prediction", num_train_timesteps: int = 1000):
        self.config = SimpleNamespace(prediction_type=prediction_type, num_train_timesteps=num_train_timesteps)
        self.alphas_cumprod = alphas_cumprod(num_train_timesteps)

    def _coef(self, timesteps, like):
        acp = self.alphas_cumprod.to(device=like.device, dtype=like.dtype)[timesteps]
        shape = (-1,) + (1,) * (like.ndim - 1)
        return (acp ** 0.5).reshape(shape), ((1 - acp) ** 0.5).reshape(shape)

    def add_noise(self, original_samples, noise, timesteps):
        a, s = self._coef(timesteps, original_samples)
        return a * original_samples + s * noise

    def get_velocity(self, sample, noise, timesteps):
        a, s = self._coef(timesteps, sample)
        return a * noise - s * sample


class DDIMScheduler:
    """eta = 0 DDIM with SD's `leading` spacing, steps_offset = 1, set_alpha_to_one = False."""
    order = 1
    init_noise_sigma = 1.0

    def __init__(self, prediction_type: str = "v_prediction", num_train_timesteps: int = 1000, steps_offset: int = 1):
        self.config = SimpleNamespace(prediction_type=prediction_type, num_train_timesteps=num_train_timesteps,
                                      steps_offset=steps_offset)
        self.alphas_cumprod = alphas_cumprod(num_train_timesteps)
        self.final_alpha_cumprod = self.alphas_cumprod[0]
        self.timesteps = None
        self.num_inference_steps = None

    def set_timesteps(self, num_inference_steps: int, device=None):
        self.num_inference_steps = num_inference_steps
        ratio = self.config.num_train_timesteps // num_inference_steps
        ts = (torch.arange(0, num_inference_steps) * ratio).round().flip(0).to(torch.int64) + self.config.steps_offset
        self.timesteps = ts.to(device) if device is not None else ts

    def scale_model_input(self, sample, timestep=None):
        return sample

    def coefficients(self, t: int):
        """(alpha_cumprod_t, alpha_cumprod_prev) for timestep t."""
        prev = t - self.config.num_train_timesteps // self.num_inference_steps
        a_t = float(self.alphas_cumprod[t])
        a_prev = float(self.alphas_cumprod[prev]) if prev >= 0 else float(self.final_alpha_cumprod)
        return a_t, a_prev

    def step(self, model_output, timestep, sample, eta: float = 0.0, **_):
        """Host-side reference form of the update (the pipeline uses the fused CUDA kernel instead)."""
        if eta != 0.0:
            raise NotImplementedError("only eta = 0 (deterministic DDIM) is implemented")
        a_t, a_prev = self.coefficients(int(timestep))
        if self.config.prediction_type == "epsilon":
            x0 = (sample - (1 - a_t) ** 0.5 * model_output) / a_t ** 0.5
            eps = model_output
        else:
            x0 = a_t ** 0.5 * sample - (1 - a_t) ** 0.5 * model_output
            eps = a_t ** 0.5 * model_output + (1 - a_t) ** 0.5 * sample
        return SimpleNamespace(prev_sample=a_prev ** 0.5 * x0 + (1 - a_prev) ** 0.5 * eps)
