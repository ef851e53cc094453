"""Coach — the reference's training driver (reference training/coach.py:36-834) reduced to the part that IS the hot
path, with the same step semantics and method names:

    coach.py:165-169  latents = vae.encode(pixel_values).latent_dist.sample().detach() * scaling_factor   [models/vae.py]
    coach.py:172-183  noise, timesteps ~ U[0, T), noisy_latents = scheduler.add_noise(latents, noise, timesteps)
    coach.py:186-194  _hs = self.get_text_conditioning(...)            -> context dict (XTI protocol)
    coach.py:197-198  model_pred = self.unet(noisy_latents, timesteps, _hs).sample      [CUDA library]
    coach.py:201-209  target = noise | scheduler.get_velocity(...)
    coach.py:211-214  loss = F.mse_loss(model_pred.float(), target.float()); backward   [CUDA dgrad-only backward]
    coach.py:216-218  optimizer.step(); lr_scheduler.step(); optimizer.zero_grad()

What produces the context dict (the NeTI mappers inside the CLIP text encoder, coach.py:276-311) is outside this
round's scope (SURVEY.md 8f "next" #1/#2); it is injected as `conditioning`, any nn.Module / callable returning the
dict.  Multi-GPU is batch-parallel: one process per GPU and ONE all-reduce of the flat trainable-parameter gradient
buffer per step (training/dist.py) instead of accelerate's DDP wrapper (coach.py:97-99).
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import Callable, Dict, Iterable, Optional

import torch
import torch.nn.functional as F

from ..schedulers import DDPMScheduler
from .dist import FlatGradAllReducer


class Coach:

    def __init__(self, cfg, unet, conditioning: Callable[..., Dict], noise_scheduler: Optional[DDPMScheduler] = None,
                 optimizer: Optional[torch.optim.Optimizer] = None, lr_scheduler=None, generator: Optional[torch.Generator] = None,
                 vae=None):
        self.cfg = cfg
        self.unet = unet
        self.vae = vae
        self.conditioning = conditioning
        self.noise_scheduler = noise_scheduler or DDPMScheduler()
        params = list(conditioning.parameters()) if isinstance(conditioning, torch.nn.Module) else []
        self.optimizer = optimizer or (torch.optim.AdamW(params, lr=getattr(getattr(cfg, "optim", cfg), "learning_rate", 1e-3))
                                       if params else None)
        self.lr_scheduler = lr_scheduler
        self.reducer = FlatGradAllReducer(params) if params else None
        if self.reducer is not None:
            self.reducer.broadcast_parameters_(0)       # what accelerate's DDP wrapper does at coach.py:97-99
        self.generator = generator
        self.global_step = 0

    # same name / argument meaning as reference coach.py:276-283
    def get_text_conditioning(self, input_ids=None, timesteps=None, input_ids_placeholder_object=None,
                              input_ids_placeholder_view=None, device=None, original_ti: bool = False) -> Dict:
        return self.conditioning(input_ids=input_ids, timesteps=timesteps,
                                 input_ids_placeholder_object=input_ids_placeholder_object,
                                 input_ids_placeholder_view=input_ids_placeholder_view, device=device, original_ti=original_ti)

    def encode_images(self, pixel_values: torch.Tensor) -> torch.Tensor:
        """coach.py:165-169: images in [-1, 1] -> scaled latents, frozen VAE, no autograd history."""
        if self.vae is None:
            raise ValueError("Coach: batch carries pixel_values but no vae was given")
        dist = self.vae.encode(pixel_values).latent_dist
        return dist.sample(self.generator).detach() * self.vae.config.scaling_factor

    def train_step(self, latents: Optional[torch.Tensor] = None, batch: Optional[Dict] = None) -> torch.Tensor:
        """One optimisation step.  `latents` given: the step starts at coach.py:172 (pre-encoded data); otherwise
        `batch["pixel_values"]` goes through the VAE first, as the reference does every step."""
        batch = batch or {}
        if latents is None:
            latents = self.encode_images(batch["pixel_values"])
        dev = latents.device
        noise = torch.randn(latents.shape, generator=self.generator, device=dev, dtype=latents.dtype)
        bsz = latents.shape[0]
        timesteps = torch.randint(0, self.noise_scheduler.config.num_train_timesteps, (bsz,), generator=self.generator,
                                  device=dev).long()
        noisy_latents = self.noise_scheduler.add_noise(latents, noise, timesteps)
        _hs = self.get_text_conditioning(input_ids=batch.get("input_ids"), timesteps=timesteps,
                                         input_ids_placeholder_object=batch.get("input_ids_placeholder_object"),
                                         input_ids_placeholder_view=batch.get("input_ids_placeholder_view"), device=dev)
        model_pred = self.unet(noisy_latents, timesteps, _hs).sample
        if self.noise_scheduler.config.prediction_type == "epsilon":
            target = noise
        elif self.noise_scheduler.config.prediction_type == "v_prediction":
            target = self.noise_scheduler.get_velocity(latents, noise, timesteps)
        else:
            raise ValueError(f"Unknown prediction type {self.noise_scheduler.config.prediction_type}")
        loss = F.mse_loss(model_pred.float(), target.float(), reduction="mean")
        loss.backward()
        if self.reducer is not None:
            self.reducer.allreduce_()
        if self.optimizer is not None:
            self.optimizer.step()
            if self.lr_scheduler is not None:
                self.lr_scheduler.step()
            self.optimizer.zero_grad()
        self.global_step += 1
        return loss.detach()

    def train(self, latent_batches: Iterable, max_train_steps: Optional[int] = None):
        """Batches are latent tensors, or dicts as the reference's dataloader yields (`pixel_values`, `input_ids`, ...)."""
        max_steps = max_train_steps or getattr(getattr(self.cfg, "optim", SimpleNamespace()), "max_train_steps", None)
        losses = []
        for b in latent_batches:
            losses.append(self.train_step(batch=b) if isinstance(b, dict) else self.train_step(b))
            if max_steps is not None and self.global_step >= max_steps:
                break
        return losses


class SyntheticConditioning(torch.nn.Module):
    """Stand-in for the NeTI mapper + CLIP path: a trainable table that emits the XTI context dict
    {"this_idx", "CONTEXT_TENSOR_i", "CONTEXT_TENSOR_BYPASS_i"} (one [B,77,D] pair per UNet layer).  Used by tests,
    smoke and the multi-GPU bench so that the step has real trainable parameters downstream of d_ctx."""

    def __init__(self, n_layers: int = 16, context_len: int = 77, dim: int = 1024, rank: int = 8, seed: int = 0):
        super().__init__()
        g = torch.Generator().manual_seed(seed)
        self.n_layers = n_layers
        self.base = torch.nn.Parameter(torch.randn(2, n_layers, context_len, rank, generator=g))
        self.proj = torch.nn.Parameter(torch.randn(rank, dim, generator=g) / rank ** 0.5)

    def forward(self, timesteps=None, device=None, **_) -> Dict:
        bsz = 1 if timesteps is None else timesteps.shape[0]
        ctx = (self.base @ self.proj)                                # [2, L, 77, D]
        out: Dict = {"this_idx": 0}
        for i in range(self.n_layers):
            out[f"CONTEXT_TENSOR_{i}"] = ctx[0, i].unsqueeze(0).expand(bsz, -1, -1)
            out[f"CONTEXT_TENSOR_BYPASS_{i}"] = ctx[1, i].unsqueeze(0).expand(bsz, -1, -1)
        return out
