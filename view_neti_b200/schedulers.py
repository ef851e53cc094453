"""Noise schedules used on the hot path (diffusers DDPMScheduler / DDIMScheduler as configured by SD-2.1):
scaled-linear betas 0.00085 -> 0.012 over 1000 steps.  Host side only holds the alphas_cumprod table; the per-step
arithmetic of the sampler runs in the fused CFG + sampler-step kernels (vn_cfg_ddim_step, vn_cfg_dpmpp_step).

    reference training/coach.py:182-183   noisy = scheduler.add_noise(latents, noise, timesteps)
    reference training/coach.py:201-205   target = noise | scheduler.get_velocity(latents, noise, timesteps)
    reference sd_pipeline_call.py:49,101  scheduler.set_timesteps(...); scheduler.step(...).prev_sample
    reference training/validate.py:568, training/inference_dtu.py:304, scripts/inference.py
                                          pipeline.scheduler = DPMSolverMultistepScheduler.from_config(...)
"""
from __future__ import annotations

import math
from types import SimpleNamespace

import numpy as np

import torch


def alphas_cumprod(num_train_timesteps: int = 1000, beta_start: float = 0.00085, beta_end: float = 0.012) -> torch.Tensor:
    betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
    return torch.cumprod(1.0 - betas, dim=0)


class DDPMScheduler:
    def __init__(self, prediction_type: str = "v_prediction", num_train_timesteps: int = 1000):
        self.config = SimpleNamespace(prediction_type=prediction_type, num_train_timesteps=num_train_timesteps)
        self.alphas_cumprod = alphas_cumprod(num_train_timesteps)

    def _coef(self, timesteps, like):
        # the table is kept per (device, dtype): moving it from pageable host memory every call synchronises the stream
        # (the runtime drains the stream before a pageable copy), which cost the train loop two full stalls per step
        key = (like.device, like.dtype)
        cache = self.__dict__.setdefault("_acp_cache", {})
        if key not in cache:
            cache[key] = self.alphas_cumprod.to(device=like.device, dtype=like.dtype)
        acp = cache[key][timesteps]
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


class DPMSolverMultistepScheduler:
    """DPM-Solver++(2M), the sampler the reference's inference scripts put on the pipeline: diffusers 0.14 defaults
    (algorithm_type "dpmsolver++", solver_order 2, solver_type "midpoint", lower_order_final, no thresholding).
    Every update is linear in (sample, data prediction, previous data prediction), so the pipeline asks for five
    scalars per step (`kernel_coefficients`) and one fused kernel does guidance + conversion + update; `step` is the
    stateful host form with diffusers' call shape."""
    order = 1
    init_noise_sigma = 1.0

    def __init__(self, prediction_type: str = "v_prediction", num_train_timesteps: int = 1000, solver_order: int = 2,
                 lower_order_final: bool = True):
        if solver_order not in (1, 2):
            raise NotImplementedError("DPM-Solver++ orders 1 and 2 are implemented (the reference uses the default, 2)")
        self.config = SimpleNamespace(prediction_type=prediction_type, num_train_timesteps=num_train_timesteps,
                                      solver_order=solver_order, lower_order_final=lower_order_final,
                                      algorithm_type="dpmsolver++", solver_type="midpoint")
        self.alphas_cumprod = alphas_cumprod(num_train_timesteps)
        acp = self.alphas_cumprod.double()
        self.alpha_t, self.sigma_t = acp.sqrt(), (1 - acp).sqrt()
        self.lambda_t = self.alpha_t.log() - self.sigma_t.log()
        self.timesteps = None
        self.num_inference_steps = None
        self._x0_prev = None
        self._i = 0

    @classmethod
    def from_config(cls, config, **kw):
        """`DPMSolverMultistepScheduler.from_config(pipeline.scheduler.config)`: carries over what the solver uses."""
        get = (lambda k, d: config.get(k, d)) if isinstance(config, dict) else (lambda k, d: getattr(config, k, d))
        return cls(prediction_type=get("prediction_type", "v_prediction"),
                   num_train_timesteps=get("num_train_timesteps", 1000), **kw)

    def set_timesteps(self, num_inference_steps: int, device=None):
        self.num_inference_steps = num_inference_steps
        n = self.config.num_train_timesteps
        # numpy on purpose: diffusers builds this list with np.linspace(...).round() and ties (x.5) fall on the side
        # numpy's linspace arithmetic puts them (N = 30: 499, where torch.linspace gives 500)
        ts = torch.from_numpy(np.linspace(0, n - 1, num_inference_steps + 1).round()[::-1][:-1].copy().astype(np.int64))
        self.timesteps = ts.to(device) if device is not None else ts
        self._x0_prev, self._i = None, 0

    def scale_model_input(self, sample, timestep=None):
        return sample

    def kernel_coefficients(self, i: int):
        """Step i of the current schedule as (p, q, A, B0, B1):
            x0 = p * sample + q * model_output;   prev_sample = A * sample + B0 * x0 + B1 * x0_of_step_(i-1)."""
        ts = [int(t) for t in self.timesteps]
        t = ts[i]
        a_s, s_s, l_s = float(self.alpha_t[t]), float(self.sigma_t[t]), float(self.lambda_t[t])
        p, q = (1.0 / a_s, -s_s / a_s) if self.config.prediction_type == "epsilon" else (a_s, -s_s)
        nxt = ts[i + 1] if i + 1 < len(ts) else 0
        a_t, s_t, l_t = float(self.alpha_t[nxt]), float(self.sigma_t[nxt]), float(self.lambda_t[nxt])
        h = l_t - l_s
        c = -a_t * math.expm1(-h)
        first = (i == 0 or self.config.solver_order == 1
                 or (i == len(ts) - 1 and self.config.lower_order_final and len(ts) < 15))
        if first:
            return p, q, s_t / s_s, c, 0.0
        r0 = (l_s - float(self.lambda_t[ts[i - 1]])) / h
        return p, q, s_t / s_s, c * (1.0 + 0.5 / r0), -c * 0.5 / r0

    def step(self, model_output, timestep, sample, **_):
        """Host-side form with diffusers' call shape (the pipeline uses the fused CUDA kernel instead)."""
        i = self._i
        if int(self.timesteps[i]) != int(timestep):
            raise ValueError(f"step {i} of the schedule is t={int(self.timesteps[i])}, got t={int(timestep)}")
        p, q, A, B0, B1 = self.kernel_coefficients(i)
        x0 = p * sample + q * model_output
        prev = A * sample + B0 * x0 + (B1 * self._x0_prev if B1 != 0.0 else 0.0)
        self._x0_prev, self._i = x0, i + 1
        return SimpleNamespace(prev_sample=prev)
