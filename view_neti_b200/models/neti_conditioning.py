"""The conditioning path of one train step, batched: reference training/coach.py:276-311 runs the NeTI text encoder once
per UNet layer (16 passes); here the 16 passes are ONE [16*B, 77, C] pass.

    mapper outputs for all (timestep_b, layer) pairs         one fused launch per mapper   (models/neti_mapper.py)
    -> overwrite the placeholder rows of the token embeddings   net_clip_text_embedding.py:61-131
    -> + position embeddings -> CLIP encoder                    models/clip_encoder.py (vn_gemm / LayerNorm / GELU / attention)
    -> bypass injection at the placeholder rows                 neti_clip_text_encoder.py:118-178
    -> final LayerNorm of the plain and the bypass states       :180-182
    -> {"this_idx": 0, "CONTEXT_TENSOR_i", "CONTEXT_TENSOR_BYPASS_i"}   coach.py:287-305

The embedding scatter, the bypass arithmetic on 16*B rows and the final LayerNorm are a handful of small torch ops
(autograd carries the gradient between the CUDA encoder and the CUDA mappers); everything with FLOPs in it is the
library.  `NeTIConditioning` plugs into `Coach(cfg, unet, conditioning=...)` and is called exactly like
`Coach.get_text_conditioning`.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F

from .._abi import VNError
from ..constants import UNET_LAYERS
from .clip_encoder import CLIPEncoder
from .neti_mapper import NeTIMapper


class NeTIConditioning(torch.nn.Module):

    def __init__(self, token_embedding: torch.Tensor, position_embedding: torch.Tensor,
                 final_layer_norm: Tuple[torch.Tensor, torch.Tensor], encoder: CLIPEncoder,
                 mapper_object_lookup: Optional[Dict[int, NeTIMapper]], mapper_view: Optional[NeTIMapper],
                 n_layers: int = len(UNET_LAYERS), layer_norm_eps: float = 1e-5, weight_dtype: torch.dtype = torch.float32):
        super().__init__()
        dev = encoder.engine.dev
        # frozen text-model pieces (coach.py:649-652 freezes the text encoder except the mappers)
        self.register_buffer("token_embedding", token_embedding.to(dev, torch.float32), persistent=False)
        self.register_buffer("position_embedding", position_embedding.to(dev, torch.float32), persistent=False)
        self.register_buffer("final_ln_weight", final_layer_norm[0].to(dev, torch.float32), persistent=False)
        self.register_buffer("final_ln_bias", final_layer_norm[1].to(dev, torch.float32), persistent=False)
        self.encoder = encoder
        self.n_layers = n_layers
        self.eps = layer_norm_eps
        self.weight_dtype = weight_dtype
        # like net_clip_text_embedding.py:25-32 the object mappers live in a dict keyed by placeholder token id; registering
        # them in a ModuleDict makes them visible to .parameters() (the reference's plain dict hides them from DDP)
        self.mapper_object_lookup = torch.nn.ModuleDict({str(k): v for k, v in (mapper_object_lookup or {}).items()})
        self.mapper_view = mapper_view

    @staticmethod
    def _as_list(ids):
        """Placeholder ids as python ints with at most ONE device round trip (they usually arrive as host tensors / lists
        from the data loader; per-element reads of a CUDA tensor would cost one synchronisation each)."""
        if ids is None:
            return None
        return [int(v) for v in (ids.tolist() if torch.is_tensor(ids) else ids)]

    def _ids_on_device(self, ids) -> torch.Tensor:
        """Placeholder ids (Python ints) as a device tensor.  Cached by value: an H2D copy from pageable memory drains the
        stream first, i.e. every fresh `torch.tensor(list, device=cuda)` in the step would be a host stall."""
        key = tuple(ids)
        cache = self.__dict__.setdefault("_id_cache", {})
        t = cache.get(key)
        if t is None:
            if len(cache) > 4096:
                cache.clear()
            t = cache[key] = torch.tensor(list(key), device=self.token_embedding.device)
        return t

    @staticmethod
    def _positions(input_ids: torch.Tensor, placeholder: torch.Tensor):
        locs = input_ids == placeholder.unsqueeze(1)
        return locs.float().argmax(1), (locs.sum(1) == 1).all()

    @staticmethod
    def _inject_bypass(state, rows, pos, bypass, unconstrained: bool, alpha: float):
        existing = state[rows, pos]
        if not unconstrained:
            b = bypass / bypass.norm(dim=1, keepdim=True) * existing.norm(dim=1, keepdim=True)
            new = existing + alpha * b
        else:
            normalizing = state.norm(dim=-1).mean(-1).detach()
            new = bypass / bypass.norm(dim=1, keepdim=True) * normalizing.unsqueeze(1)
        out = state.clone()
        out[rows, pos] = new.to(state.dtype)
        return out

    def _check_placeholders_later(self, flag: torch.Tensor) -> None:
        """net_clip_text_embedding.py:99,130 assert that every prompt holds its placeholder exactly once.  Reading the flag back
        at once would stall the host until the device has drained (every step would start on an empty queue: ~2 ms of device
        idle per step, scripts/full_step_profile.py); it travels to pinned host memory asynchronously instead and is looked at
        as soon as it has landed - at the latest at the next call - where a bad prompt still raises."""
        self._raise_if_bad_placeholders(block=False)
        if flag.device.type != "cuda":
            if not bool(flag):
                raise VNError("every prompt must hold its placeholder token exactly once (net_clip_text_embedding.py:99,130)")
            return
        if getattr(self, "_ok_host", None) is None:
            self._ok_host = torch.ones(1, dtype=torch.bool).pin_memory()
            self._ok_event = torch.cuda.Event()
        self._raise_if_bad_placeholders(block=True)          # (at most one check in flight: the previous one is read first)
        self._ok_host.copy_(flag.reshape(1), non_blocking=True)
        self._ok_event.record()
        self._ok_pending = True

    def _raise_if_bad_placeholders(self, block: bool) -> None:
        if not getattr(self, "_ok_pending", False):
            return
        if not block and not self._ok_event.query():
            return
        self._ok_event.synchronize()
        self._ok_pending = False
        if not bool(self._ok_host[0]):
            raise VNError("every prompt must hold its placeholder token exactly once (net_clip_text_embedding.py:99,130)")

    def set_mapper(self, mapper_object_lookup: Optional[Dict[int, NeTIMapper]], mapper_view: Optional[NeTIMapper],
                   device=None) -> None:
        """net_clip_text_embedding.py:25-32 (`embeddings.set_mapper`): install / replace the mappers."""
        dev = self.token_embedding.device if device is None else device
        self.mapper_object_lookup = torch.nn.ModuleDict({str(k): v.to(dev) for k, v in (mapper_object_lookup or {}).items()})
        self.mapper_view = mapper_view.to(dev) if mapper_view is not None else None

    def encode(self, input_ids: torch.Tensor, timesteps: torch.Tensor, unet_layers: torch.Tensor, ph_o, ph_v,
               use_mappers: bool = True):
        """One stacked text-encoder pass over N = len(unet_layers) sequences: row r is prompt `input_ids[r]` at
        (timesteps[r], unet_layers[r]).  Returns the final-LayerNorm'ed plain states and bypass states ([N, L, C] each, the
        latter None without a bypass).  `use_mappers=False` is the plain CLIP text model (negative prompt,
        sd_pipeline_call.py:36-41)."""
        dev = self.token_embedding.device
        N, L = input_ids.shape
        C = self.token_embedding.shape[1]
        rows = torch.arange(N, device=dev)
        emb = self.token_embedding[input_ids]                                  # [N, L, C] (a fresh tensor: written below)
        obj = view = None
        ok = []
        t_rep, l_rep = timesteps.float(), unet_layers.float()
        if use_mappers and len(self.mapper_object_lookup) > 0 and ph_o is not None and ph_o[0] != -1:
            if any(v != ph_o[0] for v in ph_o):
                raise VNError("one object per batch (net_clip_text_embedding.py:67-68)")
            mapper = self.mapper_object_lookup[str(ph_o[0])]
            out = mapper(timestep=t_rep, unet_layer=l_rep, input_ids_placeholder_view=None, truncation_idx=None)
            pos_o, good = self._positions(input_ids, self._ids_on_device(ph_o))
            ok.append(good)
            emb[rows, pos_o] = out.word_embedding.to(emb.dtype)
            obj = (out, pos_o)
        if use_mappers and self.mapper_view is not None and ph_v is not None and not all(v == -1 for v in ph_v):   # :105-106
            out = self.mapper_view(timestep=t_rep, unet_layer=l_rep, input_ids_placeholder_view=ph_v, truncation_idx=None)
            pos_v, good = self._positions(input_ids, self._ids_on_device(ph_v))
            ok.append(good)
            emb[rows, pos_v] = out.word_embedding.to(emb.dtype)
            view = (out, pos_v)
        if ok:
            self._check_placeholders_later(torch.stack(ok).all())
        x = emb + self.position_embedding[:L]
        last = self.encoder(inputs_embeds=x)[0]
        with_bypass = None
        for item in (obj, view):                                               # object first, then view (:127-178)
            if item is not None and item[0].bypass_output is not None:
                out, pos = item
                base = last if with_bypass is None else with_bypass
                with_bypass = self._inject_bypass(base, rows, pos, out.bypass_output.to(last.dtype), out.bypass_unconstrained,
                                                  out.output_bypass_alpha)
        ln = lambda s: F.layer_norm(s, (C,), self.final_ln_weight, self.final_ln_bias, self.eps).to(self.weight_dtype)   # noqa: E731
        return ln(last), (ln(with_bypass) if with_bypass is not None else None)

    def forward(self, input_ids: torch.Tensor = None, timesteps: torch.Tensor = None,
                input_ids_placeholder_object=None, input_ids_placeholder_view=None, device=None,
                original_ti: bool = False, **_) -> Dict:
        """All 16 UNet layers of `Coach.get_text_conditioning` (coach.py:276-311) as ONE stacked pass (row = layer * B + b)."""
        dev = self.token_embedding.device
        input_ids = torch.as_tensor(input_ids, device=dev)
        timesteps = torch.as_tensor(timesteps, device=dev)
        B, L = input_ids.shape
        nl = 1 if original_ti else self.n_layers
        C = self.token_embedding.shape[1]
        ph_o = self._as_list(input_ids_placeholder_object)
        ph_v = self._as_list(input_ids_placeholder_view)
        last_n, bypass_n = self.encode(input_ids.repeat(nl, 1), timesteps.repeat(nl),
                                       torch.arange(nl, device=dev).repeat_interleave(B),
                                       ph_o * nl if ph_o is not None else None, ph_v * nl if ph_v is not None else None)
        if original_ti:
            return last_n[:B]                                                  # coach.py:307-309
        hs: Dict = {"this_idx": 0}
        # unbind, not 16 slices: its backward is ONE stack of the 16 context gradients instead of 16 zero-filled
        # full-size tensors that autograd would have to add up
        plain = last_n.view(nl, B, L, C).unbind(0)
        bypass = bypass_n.view(nl, B, L, C).unbind(0) if bypass_n is not None else None
        for i in range(nl):
            hs[f"CONTEXT_TENSOR_{i}"] = plain[i]
            if bypass is not None:
                hs[f"CONTEXT_TENSOR_BYPASS_{i}"] = bypass[i]
        return hs
