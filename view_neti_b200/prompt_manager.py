"""PromptManager - the reference's inference-time conditioning cache (reference prompt_manager.py:13-101) on the batched
conditioning path.

The reference computes, for one prompt, the XTI context dict of every inference timestep with len(timesteps) x 16 separate
text-encoder passes (800 for the 50-step schedule).  Here the timesteps are the batch dimension of `NeTIConditioning`
(which already stacks the 16 UNet layers), so a chunk of `chunk` timesteps is ONE [16*chunk, 77, C] encoder pass without
gradient bookkeeping; the result is the same List[Dict] (one dict per timestep, tensors repeated `num_images_per_prompt`
times) that `sd_pipeline_call(prompt_embeds=...)` consumes (sd_pipeline_call.py:78-94).
"""
from __future__ import annotations

from typing import Any, Dict, List, Optional, Sequence

import torch

from . import constants
from .models.neti_conditioning import NeTIConditioning


class PromptManager:
    """Precomputes, for one prompt, the per-timestep / per-UNet-layer context dicts of a whole sampling schedule."""

    def __init__(self, tokenizer, text_encoder: NeTIConditioning, timesteps: Sequence[int] = constants.SD_INFERENCE_TIMESTEPS,
                 unet_layers: List[str] = constants.UNET_LAYERS, placeholder_view_token_ids: List[int] = None,
                 placeholder_object_token_ids: List[int] = None, torch_dtype: torch.dtype = torch.float32, chunk: int = 10):
        self.tokenizer = tokenizer
        self.text_encoder = text_encoder
        self.timesteps = [int(t) for t in timesteps]
        self.unet_layers = unet_layers
        self.placeholder_view_token_ids = placeholder_view_token_ids
        self.placeholder_object_token_ids = placeholder_object_token_ids
        self.dtype = torch_dtype
        self.chunk = chunk
        assert len(unet_layers) == text_encoder.n_layers

    def _tokenize(self, text) -> torch.Tensor:
        if torch.is_tensor(text):                        # already token ids [1, 77] (no tokenizer files exist offline)
            return text.view(1, -1).long()
        return self.tokenizer(text, padding="max_length", max_length=self.tokenizer.model_max_length,
                              return_tensors="pt").input_ids

    @staticmethod
    def _placeholder(ids: torch.Tensor, placeholder_token_ids, text) -> torch.Tensor:
        """prompt_manager.py:62-71: which of the placeholder ids the prompt holds (-1: none)."""
        if not placeholder_token_ids:
            return torch.tensor([-1])
        locs = torch.isin(ids.cpu(), torch.tensor(list(placeholder_token_ids)))
        if locs.sum() == 0:
            return torch.tensor([-1])
        assert int(locs.sum()) == 1, f"should be exactly 1 placeholder token of a kind per prompt, for prompt [`{text}`]"
        return ids.cpu()[torch.where(locs)]

    @torch.no_grad()
    def embed_prompt(self, text, truncation_idx: Optional[int] = None, num_images_per_prompt: int = 1) -> List[Dict[str, Any]]:
        if truncation_idx is not None:
            raise NotImplementedError("nested-dropout truncation is off in the shipped configs and not implemented")
        ids = self._tokenize(text)
        ph_obj = self._placeholder(ids, self.placeholder_object_token_ids, text)
        ph_view = self._placeholder(ids, self.placeholder_view_token_ids, text)
        dev = self.text_encoder.token_embedding.device
        out: List[Dict[str, Any]] = []
        for c0 in range(0, len(self.timesteps), self.chunk):
            ts = torch.tensor(self.timesteps[c0:c0 + self.chunk], device=dev)
            n = ts.numel()
            hs = self.text_encoder(input_ids=ids.to(dev).repeat(n, 1), timesteps=ts,
                                   input_ids_placeholder_object=None if int(ph_obj[0]) == -1 else ph_obj.to(dev).repeat(n),
                                   input_ids_placeholder_view=ph_view.to(dev).repeat(n))
            for j in range(n):
                d: Dict[str, Any] = {"this_idx": 0}
                for k, v in hs.items():
                    if k != "this_idx":
                        d[k] = v[j:j + 1].to(self.dtype).repeat(num_images_per_prompt, 1, 1)
                out.append(d)
        return out
