"""Drop-in for the reference's inference entry point (reference sd_pipeline_call.py:9-133): the CFG denoise loop

    for i, t in timesteps:  u = unet(x, t, negative_prompt_embeds);  c = unet(x, t, prompt_embeds[i])
                            eps = u + s (c - u);  x = scheduler.step(eps, t, x).prev_sample          (:71-101)

B200-native shape of the same loop: the unconditional and the conditional pass run as ONE batched UNet forward
(batch 2n, the negative embedding bound to every layer of the first half, the per-timestep NeTI dict to the second
half) replayed from a CUDA graph, followed by ONE fused kernel for the guidance combine + sampler update: DPM-Solver++(2M),
which the reference's inference scripts install on the pipeline (validate.py:568, inference_dtu.py:304; vn_cfg_dpmpp_step),
or eta-0 DDIM (vn_cfg_ddim_step).  Everything stays on the device; `callback` is honoured.  The negative embedding comes from `pipeline.text_encoder`
when the pipeline has one, else from `pipeline.negative_prompt_embeds`; `output_type="latent"` returns latents, other
output types go through `pipeline.decode_latents` (:115 - models/vae.py when the pipeline holds our AutoencoderKL).
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import Any, Callable, Dict, List, Optional, Union

import torch

from . import ops
from ._abi import VNError
from .schedulers import DDIMScheduler, DPMSolverMultistepScheduler


class StableDiffusionPipelineOutput(SimpleNamespace):
    pass


class ViewNeTIPipeline:
    """Minimal holder with the attribute names sd_pipeline_call touches on a diffusers pipeline."""

    def __init__(self, unet, scheduler=None, text_encoder=None, tokenizer=None, vae=None, negative_prompt_embeds=None,
                 vae_scale_factor: int = 8):
        self.unet, self.scheduler = unet, scheduler or DDIMScheduler()
        self.text_encoder, self.tokenizer, self.vae = text_encoder, tokenizer, vae
        self.negative_prompt_embeds = negative_prompt_embeds
        self.vae_scale_factor = vae_scale_factor
        self._execution_device = unet.device

    def prepare_latents(self, batch, channels, height, width, dtype, device, generator, latents=None):
        shape = (batch, channels, height // self.vae_scale_factor, width // self.vae_scale_factor)
        if latents is None:
            latents = torch.randn(shape, generator=generator, device=generator.device if generator is not None else "cpu",
                                  dtype=torch.float32).to(device)
        return latents.to(device=device, dtype=torch.float32) * self.scheduler.init_noise_sigma

    def decode_latents(self, latents):
        """diffusers StableDiffusionPipeline.decode_latents (reference sd_pipeline_call.py:115): numpy NHWC in [0, 1]."""
        if self.vae is None:
            raise VNError("decode_latents: the pipeline has no vae (use output_type='latent')")
        from .models.vae import decode_latents
        return decode_latents(self.vae, latents)

    @staticmethod
    def numpy_to_pil(images):
        from PIL import Image
        if images.ndim == 3:
            images = images[None, ...]
        return [Image.fromarray((im * 255).round().astype("uint8")) for im in images]


def get_neg_prompt_input_ids(pipeline, negative_prompt: Optional[Union[str, List[str]]] = None):
    if negative_prompt is None:
        negative_prompt = ""
    uncond_tokens = [negative_prompt] if isinstance(negative_prompt, str) else negative_prompt
    return pipeline.tokenizer(uncond_tokens, padding="max_length", max_length=pipeline.tokenizer.model_max_length,
                              truncation=True, return_tensors="pt")


def _layer_contexts(unet, embed, n: int):
    """K / V context per cross-attention layer for one pass ([n,77,D] each), following the dict protocol."""
    ctxs = unet._contexts(embed)
    L = unet.cfg.num_cross_layers
    fix = lambda c: c if c.shape[0] == n else c.expand(n, -1, -1)   # noqa: E731
    return [fix(c) for c in ctxs[:L]], [fix(c) for c in ctxs[L:]]


@torch.no_grad()
def sd_pipeline_call(pipeline, prompt_embeds, height: Optional[int] = None, width: Optional[int] = None,
                     num_inference_steps: int = 50, guidance_scale: float = 7.5,
                     negative_prompt: Optional[Union[str, List[str]]] = None, num_images_per_prompt: Optional[int] = 1,
                     eta: float = 0.0, generator=None, latents: Optional[torch.Tensor] = None,
                     output_type: Optional[str] = "pil", return_dict: bool = True,
                     callback: Optional[Callable[[int, int, torch.Tensor], None]] = None, callback_steps: int = 1,
                     cross_attention_kwargs: Optional[Dict[str, Any]] = None):
    unet = pipeline.unet
    height = height or unet.config.sample_size * pipeline.vae_scale_factor
    width = width or unet.config.sample_size * pipeline.vae_scale_factor
    device = pipeline._execution_device
    n = num_images_per_prompt or 1
    if getattr(pipeline, "text_encoder", None) is not None and getattr(pipeline, "tokenizer", None) is not None:
        neg = get_neg_prompt_input_ids(pipeline, negative_prompt)
        negative_prompt_embeds, _ = pipeline.text_encoder(input_ids=neg.input_ids.to(device), attention_mask=None)
        negative_prompt_embeds = negative_prompt_embeds[0]
    elif getattr(pipeline, "negative_prompt_embeds", None) is not None:
        negative_prompt_embeds = pipeline.negative_prompt_embeds
    else:
        raise VNError("sd_pipeline_call needs pipeline.text_encoder + tokenizer or pipeline.negative_prompt_embeds")
    negative_prompt_embeds = negative_prompt_embeds.to(device).reshape(1, *negative_prompt_embeds.shape[-2:])
    if guidance_scale <= 1.0:
        raise VNError("guidance_scale <= 1 leaves noise_pred undefined in the reference loop (sd_pipeline_call.py:97-101)")
    sched = pipeline.scheduler
    dpm = isinstance(sched, DPMSolverMultistepScheduler)
    if not (dpm or isinstance(sched, DDIMScheduler)) or (eta != 0.0 and not dpm):      # diffusers' DPM-Solver ignores eta
        raise VNError("the fused denoise loop implements DPM-Solver++(2M) and the eta = 0 DDIM update; pass a "
                      "view_neti_b200 DPMSolverMultistepScheduler or DDIMScheduler")
    sched.set_timesteps(num_inference_steps, device="cpu")
    timesteps = [int(t) for t in sched.timesteps]
    if isinstance(prompt_embeds, list) and len(prompt_embeds) < len(timesteps):
        raise VNError(f"prompt_embeds has {len(prompt_embeds)} entries for {len(timesteps)} timesteps")
    latents = pipeline.prepare_latents(n, unet.in_channels, height, width, torch.float32, device, generator, latents)
    latents = latents.contiguous().clone()
    h, w = latents.shape[-2:]
    plan = unet.engine.plan(2 * n, h, w)          # uncond rows [0, n), cond rows [n, 2n)
    nk, nv = _layer_contexts(unet, negative_prompt_embeds, n)
    vpred = 0 if sched.config.prediction_type == "epsilon" else 1
    tbuf = torch.empty(2 * n, dtype=torch.int64, device=device)
    x0_prev = torch.zeros_like(latents) if dpm else None          # DPM-Solver++ carries one data prediction
    for i, t in enumerate(timesteps):
        embed = prompt_embeds[i] if isinstance(prompt_embeds, list) else prompt_embeds
        ck, cv = _layer_contexts(unet, embed, n)
        for l in range(unet.cfg.num_cross_layers):
            plan.ctx[0, l, :n].copy_(nk[l]); plan.ctx[0, l, n:].copy_(ck[l])
            plan.ctx[1, l, :n].copy_(nv[l]); plan.ctx[1, l, n:].copy_(cv[l])
        x = sched.scale_model_input(latents, t)
        plan.latents[:n].copy_(x); plan.latents[n:].copy_(x)
        plan.timesteps.copy_(tbuf.fill_(t))
        plan.run_forward()
        unet._generation += 1
        if dpm:
            ops.cfg_dpmpp_step(latents, plan.eps[:n], plan.eps[n:], x0_prev, guidance_scale, *sched.kernel_coefficients(i))
        else:
            a_t, a_prev = sched.coefficients(t)
            ops.cfg_ddim_step(latents, plan.eps[:n], plan.eps[n:], guidance_scale, a_t, a_prev, vpred)
        if callback is not None and i % callback_steps == 0:
            callback(i, t, latents)
    has_nsfw_concept = False
    if output_type == "latent":
        image, has_nsfw_concept = latents, None
    else:
        if not hasattr(pipeline, "decode_latents"):
            raise VNError("VAE decoding is outside the hot path: use output_type='latent' or give the pipeline a decode_latents")
        image = pipeline.decode_latents(latents)
        if output_type == "pil":
            image = pipeline.numpy_to_pil(image)
    if not return_dict:
        return image, has_nsfw_concept
    return StableDiffusionPipelineOutput(images=image, nsfw_content_detected=has_nsfw_concept)
