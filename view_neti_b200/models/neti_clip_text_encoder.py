"""NeTICLIPTextModel — the object the reference calls `self.text_encoder` / `pipeline.text_encoder`
(reference models/neti_clip_text_encoder.py:15-225, models/net_clip_text_embedding.py:12-137), as a shell over the batched
CUDA conditioning path (models/neti_conditioning.py).

Call surface kept:
    text_encoder(batch=NeTIBatch(...))                   -> (BaseModelOutputWithPooling, BaseModelOutputWithPooling | None)
                                                            one UNet layer, as coach.py:289-305 / prompt_manager.py:78-99 call it
    text_encoder(input_ids=ids, attention_mask=None)     -> (out, None)      plain CLIP text model: the negative prompt of
                                                            sd_pipeline_call.py:36-41 (`embeds, _ = ...; embeds[0]`)
    text_encoder.text_model.embeddings.set_mapper(lookup, mapper_view) / .mapper_object_lookup / .mapper_view
    text_encoder.text_model.encoder / .final_layer_norm / .embeddings.position_embedding   (frozen: .requires_grad_(False))
    text_encoder.get_input_embeddings().weight / .resize_token_embeddings(n) / .dtype / .train() / .eval()
    text_encoder.gradient_checkpointing_enable()          no-op (the dgrad-only backward keeps only what it needs)

ALWAYS a 2-tuple, like the reference (:207-225).  The per-layer call exists for drop-in compatibility; the fast path is
`text_encoder.conditioning(...)` (all 16 layers in one stacked pass), which is what `Coach.get_text_conditioning` uses.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch

from .._abi import VNError
from ..utils.types import NeTIBatch
from .clip_encoder import CLIPEncoder
from .neti_conditioning import NeTIConditioning
from .neti_mapper import NeTIMapper


class BaseModelOutputWithPooling(tuple):
    """`out[0]`, `.last_hidden_state`, `.pooler_output`, `.hidden_states`, `.attentions` (transformers' record, as read at
    coach.py:297-303 and sd_pipeline_call.py:41)."""

    def __new__(cls, last_hidden_state, pooler_output):
        o = super().__new__(cls, (last_hidden_state, pooler_output))
        o.last_hidden_state, o.pooler_output = last_hidden_state, pooler_output
        o.hidden_states = o.attentions = None
        return o


class _Frozen:
    """Handle for a frozen sub-module of the text model (encoder, final_layer_norm, position_embedding): the reference only
    ever calls `.requires_grad_(False)` on them (coach.py:650-653)."""

    def __init__(self, name: str):
        self._name = name

    def requires_grad_(self, requires_grad: bool = True):
        if requires_grad:
            raise VNError(f"text_model.{self._name} is frozen on this path (coach.py:650-653): no weight gradients exist")
        return self


class _TokenEmbedding:
    def __init__(self, cond: NeTIConditioning):
        self._cond = cond

    @property
    def weight(self) -> torch.Tensor:
        return self._cond.token_embedding


class NeTICLIPTextEmbeddings:
    """`text_model.embeddings` (net_clip_text_embedding.py:12-32)."""

    def __init__(self, cond: NeTIConditioning):
        self._cond = cond
        self.token_embedding = _TokenEmbedding(cond)
        self.position_embedding = _Frozen("embeddings.position_embedding")

    def set_mapper(self, mapper_object_lookup: Optional[Dict[int, NeTIMapper]], mapper_view: Optional[NeTIMapper],
                   device="cuda") -> None:
        self._cond.set_mapper(mapper_object_lookup, mapper_view)

    @property
    def mapper_object_lookup(self) -> Dict[int, NeTIMapper]:
        return {int(k): v for k, v in self._cond.mapper_object_lookup.items()}

    @property
    def mapper_view(self) -> Optional[NeTIMapper]:
        return self._cond.mapper_view


class NeTICLIPTextTransformer:
    def __init__(self, cond: NeTIConditioning):
        self.embeddings = NeTICLIPTextEmbeddings(cond)
        self.encoder = _Frozen("encoder")
        self.final_layer_norm = _Frozen("final_layer_norm")


class NeTICLIPTextModel(torch.nn.Module):

    def __init__(self, conditioning: NeTIConditioning):
        super().__init__()
        self.conditioning = conditioning
        self.text_model = NeTICLIPTextTransformer(conditioning)

    @classmethod
    def from_parts(cls, token_embedding: torch.Tensor, position_embedding: torch.Tensor,
                   final_layer_norm: Tuple[torch.Tensor, torch.Tensor], encoder: CLIPEncoder) -> "NeTICLIPTextModel":
        """A text model without mappers yet (coach.py:617-621 loads it, :86 installs the mappers with set_mapper)."""
        return cls(NeTIConditioning(token_embedding, position_embedding, final_layer_norm, encoder, None, None))

    # ---- transformers surface the reference touches ----------------------------------------------------------------
    @property
    def dtype(self) -> torch.dtype:
        return self.conditioning.weight_dtype

    @property
    def device(self) -> torch.device:
        return self.conditioning.token_embedding.device

    def get_input_embeddings(self):
        return self.text_model.embeddings.token_embedding

    def resize_token_embeddings(self, new_num_tokens: int):
        """coach.py:362: grow the (frozen) token table for the placeholder tokens; new rows are zero until
        `_add_concept_token_to_tokenizer_static` fills them with the super-category embedding."""
        c = self.conditioning
        old = c.token_embedding
        if new_num_tokens != old.shape[0]:
            new = torch.zeros(new_num_tokens, old.shape[1], dtype=old.dtype, device=old.device)
            n = min(new_num_tokens, old.shape[0])
            new[:n] = old[:n]
            c.token_embedding = new
        return self.get_input_embeddings()

    def gradient_checkpointing_enable(self) -> None:
        pass

    # ---- forward ---------------------------------------------------------------------------------------------------
    def forward(self, input_ids: Optional[torch.Tensor] = None, attention_mask: Optional[torch.Tensor] = None,
                position_ids: Optional[torch.Tensor] = None, output_attentions: Optional[bool] = None,
                output_hidden_states: Optional[bool] = None, return_dict: Optional[bool] = None,
                batch: Optional[NeTIBatch] = None, layer_idx: Optional[int] = -1):
        if attention_mask is not None or position_ids is not None or output_attentions or output_hidden_states:
            raise VNError("NeTICLIPTextModel: attention_mask / position_ids / attention maps are not supported (the "
                          "reference passes none of them: coach.py:296, sd_pipeline_call.py:37-40)")
        c = self.conditioning
        dev = c.token_embedding.device
        if input_ids is not None:                                   # regular embedding logic (:83-88)
            ids = torch.as_tensor(input_ids, device=dev).view(-1, input_ids.shape[-1])
            n = ids.shape[0]
            zeros = torch.zeros(n, device=dev)
            last, bypass = c.encode(ids, zeros, zeros, None, None, use_mappers=False)
        elif batch is not None:                                     # NeTI logic (:93-101)
            ids = torch.as_tensor(batch.input_ids, device=dev).view(-1, batch.input_ids.shape[-1])
            if batch.truncation_idx is not None:
                raise NotImplementedError("nested-dropout truncation is off in the shipped configs and not implemented")
            last, bypass = c.encode(ids, torch.as_tensor(batch.timesteps, device=dev),
                                    torch.as_tensor(batch.unet_layers, device=dev),
                                    c._as_list(batch.input_ids_placeholder_object), c._as_list(batch.input_ids_placeholder_view))
        else:
            raise ValueError("You have to specify either batch or input_ids!")
        eot = ids.to(torch.int).argmax(dim=-1)                      # pooled output: features at the eot token (:188-203)
        rows = torch.arange(ids.shape[0], device=dev)
        out = BaseModelOutputWithPooling(last, last[rows, eot])
        if bypass is None:
            return out, None
        return out, BaseModelOutputWithPooling(bypass, bypass[rows, eot])
