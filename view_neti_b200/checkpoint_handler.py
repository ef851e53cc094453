"""Mapper / learned-embedding checkpoints in the reference's on-disk layout (reference
checkpoint_handler.py:34-97 writes, :130-230 reads):

    mapper-steps-N_object.pt   {"cfg": <nested dict>, "mappers": {token_id: {"state_dict", "encoder", "placeholder_object_token"}}}
    mapper-steps-N_view.pt     {"cfg": ..., "mappers": {"dummy_key": {...}}}
    learned_embeds-steps-N.bin {placeholder token: embedding row}

Files written by the reference pickle its own classes (the positional-encoding object under "encoder", possibly others
inside "cfg"); `load_mapper` reads them WITHOUT the reference on sys.path: unknown classes from the reference's packages
are unpickled into inert stand-ins that only keep their attributes (tensors included).  The "encoder" entry is not needed
to rebuild an arch_view_net-15 mapper (checkpoint_handler.py:219 - "only used in arch_view <= 14"); its Fourier matrix is
compared with ours as a consistency check.  Files written here store {"w": tensor} there, which the reference's loader
ignores the same way.  Host-side code: nothing here touches the GPU.
"""
from __future__ import annotations

import os
import pickle
from pathlib import Path
from typing import Any, Dict, List, Optional, Tuple, Union

import torch

from ._abi import VNError
from .models.neti_mapper import NeTIMapper
from .utils.types import PESigmas

_REFERENCE_PACKAGES = ("models", "utils", "training", "checkpoint_handler", "constants", "prompt_manager")


class _ReferenceObject:
    """Stand-in for an instance of a class that lives in the reference's source tree."""

    def __setstate__(self, state):
        self.__dict__.update(state if isinstance(state, dict) else {"_state": state})


_SAFE_MODULE_ROOTS = ("torch", "collections", "numpy", "pathlib", "_codecs")
_SAFE_BUILTINS = {"set", "frozenset", "slice", "complex", "dict", "list", "tuple", "int", "float", "str", "bool", "bytes",
                  "bytearray", "range", "object"}


class _Unpickler(pickle.Unpickler):
    """Classes from the reference's own packages become inert attribute holders; everything else must come from a short
    allowlist (tensors, containers, numpy scalars, paths) - a checkpoint cannot name arbitrary callables."""

    def find_class(self, module: str, name: str):
        root = module.split(".")[0]
        if root in _REFERENCE_PACKAGES:
            return type(name, (_ReferenceObject,), {"__module__": "reference." + module})
        if root in _SAFE_MODULE_ROOTS or (module in ("builtins", "__builtin__") and name in _SAFE_BUILTINS):
            return super().find_class(module, name)
        raise pickle.UnpicklingError(f"checkpoint names {module}.{name}, which is not on the allowlist of this loader")


class _PickleModule:
    """`pickle_module` for torch.load: the standard pickle with the stand-in unpickler."""
    __name__ = "pickle"
    Unpickler = _Unpickler
    load = staticmethod(lambda f, **kw: _Unpickler(f, **kw).load())
    loads = staticmethod(pickle.loads)
    dump, dumps = staticmethod(pickle.dump), staticmethod(pickle.dumps)
    PicklingError, UnpicklingError = pickle.PicklingError, pickle.UnpicklingError
    HIGHEST_PROTOCOL, DEFAULT_PROTOCOL = pickle.HIGHEST_PROTOCOL, pickle.DEFAULT_PROTOCOL


def _tensor_attr(obj: Any, name: str) -> Optional[torch.Tensor]:
    """An attribute of an unpickled nn.Module stand-in (parameters / buffers sit in `_parameters` / `_buffers`)."""
    if isinstance(obj, dict):
        v = obj.get(name)
        return v if torch.is_tensor(v) else None
    for store in ("__dict__", "_parameters", "_buffers"):
        d = getattr(obj, store, None) if store != "__dict__" else getattr(obj, "__dict__", None)
        if isinstance(d, dict) and torch.is_tensor(d.get(name)):
            return d[name].detach()
    return None


class CheckpointHandler:
    """Same method names as the reference's class; constructed with what `save_model` needs."""

    def __init__(self, cfg: Optional[Dict] = None, placeholder_view_tokens: Optional[List[str]] = None,
                 placeholder_view_token_ids: Optional[List[int]] = None, placeholder_object_tokens: Optional[List[str]] = None,
                 placeholder_object_token_ids: Optional[List[int]] = None, save_root: Union[str, Path] = "."):
        self.cfg = self._encode_cfg(cfg)
        self.placeholder_view_tokens = list(placeholder_view_tokens or [])
        self.placeholder_view_token_ids = list(placeholder_view_token_ids or [])
        self.placeholder_object_tokens = list(placeholder_object_tokens or [])
        self.placeholder_object_token_ids = list(placeholder_object_token_ids or [])
        self.placeholder_tokens = self.placeholder_view_tokens + self.placeholder_object_tokens
        self.placeholder_token_ids = self.placeholder_view_token_ids + self.placeholder_object_token_ids
        self.save_root = Path(save_root)

    @staticmethod
    def _encode_cfg(cfg) -> Optional[Dict]:
        """The reference stores `pyrallis.encode(cfg)` (checkpoint_handler.py:69): a plain nested dict.  A RunConfig is
        encoded the same way; a dict is taken as is; None is allowed for handlers that only load."""
        if cfg is None:
            return None
        import dataclasses
        if dataclasses.is_dataclass(cfg):
            from .training.config import to_dict
            return to_dict(cfg)
        if not isinstance(cfg, dict):
            raise VNError("CheckpointHandler: cfg must be a RunConfig or the nested dict it encodes to")
        return cfg

    def _cfg_for_saving(self) -> Dict:
        cfg = self.cfg
        if not isinstance(cfg, dict) or not isinstance(cfg.get("model"), dict) or not isinstance(cfg.get("data"), dict):
            raise VNError("CheckpointHandler.save_mapper: cfg must be a nested dict with 'model' and 'data' sections (it is what "
                          "load_mapper rebuilds the mapper from, checkpoint_handler.py:143-198); construct the handler with the "
                          "RunConfig of the run")
        return cfg

    # ---- writing (checkpoint_handler.py:34-97) ---------------------------------------------------------
    def save_model(self, conditioning, embeds_save_name: str, mapper_save_name: str) -> None:
        """`conditioning`: a NeTIConditioning (token_embedding, mapper_object_lookup, mapper_view)."""
        self.save_learned_embeds(conditioning.token_embedding, embeds_save_name)
        lookup = {int(k): v for k, v in conditioning.mapper_object_lookup.items()} if conditioning.mapper_object_lookup else None
        self.save_mapper(lookup, conditioning.mapper_view, mapper_save_name)

    def save_learned_embeds(self, token_embedding: torch.Tensor, save_name: str) -> None:
        rows = token_embedding.detach()[self.placeholder_token_ids].cpu()
        torch.save({t: v for t, v in zip(self.placeholder_tokens, rows)}, self.save_root / save_name)

    @staticmethod
    def _entry(mapper: NeTIMapper, token: str) -> Dict:
        sd = {k: v.detach().cpu() for k, v in mapper.state_dict().items()}
        return {"state_dict": sd, "encoder": {"w": mapper.encoder_w.detach().cpu()}, "placeholder_object_token": token}

    def save_mapper(self, mapper_object_lookup: Optional[Dict[int, NeTIMapper]], mapper_view: Optional[NeTIMapper],
                    save_name: str) -> None:
        stem, suffix = Path(save_name).stem, Path(save_name).suffix
        cfg = self._cfg_for_saving()
        if mapper_object_lookup is not None:
            ckpt = {"cfg": cfg, "mappers": {k: self._entry(m, m.placeholder_object_token) for k, m in mapper_object_lookup.items()}}
            torch.save(ckpt, os.path.join(self.save_root, stem + "_object" + suffix))
        if mapper_view is not None:
            ckpt = {"cfg": cfg, "mappers": {"dummy_key": self._entry(mapper_view, "dummy")}}
            torch.save(ckpt, os.path.join(self.save_root, stem + "_view" + suffix))

    # ---- reading (checkpoint_handler.py:130-230) ------------------------------------------------------
    @staticmethod
    def load_mapper(mapper_path: Union[str, Path], embedding_type: str = "object",
                    placeholder_view_tokens: Optional[List[str]] = None, placeholder_view_token_ids: Optional[List[int]] = None,
                    placeholder_object_tokens: Optional[List[str]] = None, placeholder_object_token_ids: Optional[List[int]] = None,
                    device: Union[str, torch.device] = "cpu", cam_mins: Optional[torch.Tensor] = None,
                    cam_maxs: Optional[torch.Tensor] = None) -> Tuple[Dict, Union[NeTIMapper, Dict[int, NeTIMapper]]]:
        """Returns (cfg dict, view mapper) or (cfg dict, {placeholder token id: object mapper}) like the reference; mappers
        land on `device` (their forward needs CUDA, loading does not)."""
        ckpt = torch.load(mapper_path, map_location="cpu", weights_only=False, pickle_module=_PickleModule)
        cfg = ckpt["cfg"]
        model = cfg["model"] if isinstance(cfg, dict) else _ReferenceObject.__getattribute__(cfg, "__dict__")["model"]
        get = (lambda k, d=None: model.get(k, d)) if isinstance(model, dict) else (lambda k, d=None: getattr(model, k, d))
        if embedding_type == "view":
            if placeholder_view_tokens is None or placeholder_view_token_ids is None:
                raise VNError("view mappers need placeholder_view_tokens / placeholder_view_token_ids (checkpoint_handler.py:146)")
            output_bypass, target_norm = get("output_bypass_view", True), get("target_norm_view")
            unconstrained = get("bypass_unconstrained_view", False)
        elif embedding_type == "object":
            placeholder_view_tokens = placeholder_view_token_ids = None
            output_bypass, target_norm = get("output_bypass_object", True), get("target_norm_object")
            unconstrained = get("bypass_unconstrained_object", False)
            if target_norm is None and get("normalize_object_mapper_output", True):
                raise ValueError("need a target norm to pass to pretrained object mapper")      # :153-155
        else:
            raise ValueError(embedding_type)
        alpha = get("output_bypass_alpha_object", 0.2)             # (the reference reads the object key for both, :163-170)
        sig = get("pe_sigmas")
        sig = sig if isinstance(sig, dict) else dict(getattr(sig, "__dict__", {}))
        pe_sigmas = PESigmas(**{k: sig.get(k) for k in ("sigma_t", "sigma_l", "sigma_theta", "sigma_phi", "sigma_r", "sigma_dtu12")})
        out: Dict[Any, NeTIMapper] = {}
        for key, entry in ckpt["mappers"].items():
            token = entry["placeholder_object_token"]
            if embedding_type == "view":
                out_key: Any = "dummy"
            else:
                out_key = dict(zip(placeholder_object_tokens or [], placeholder_object_token_ids or [])).get(token, key)
            m = NeTIMapper(embedding_type=embedding_type, placeholder_view_tokens=placeholder_view_tokens,
                           placeholder_view_token_ids=placeholder_view_token_ids, output_dim=get("word_embedding_dim"),
                           arch_mlp_hidden_dims=get("arch_mlp_hidden_dims"), use_nested_dropout=get("use_nested_dropout"),
                           nested_dropout_prob=get("nested_dropout_prob", 0.5),
                           norm_scale=None if target_norm is None else torch.tensor(float(target_norm)),
                           use_positional_encoding=get("use_positional_encoding_object", 1),
                           num_pe_time_anchors=get("num_pe_time_anchors", 10), pe_sigmas=pe_sigmas,
                           arch_view_net=get("arch_view_net"), arch_view_mix_streams=get("arch_view_mix_streams", 0),
                           arch_view_disable_tl=get("arch_view_disable_tl"), original_ti=get("original_ti", False),
                           output_bypass=output_bypass, output_bypass_alpha=alpha, placeholder_object_token=token,
                           bypass_unconstrained=unconstrained, cam_mins=cam_mins, cam_maxs=cam_maxs)
            state = {k: v for k, v in entry["state_dict"].items() if k != "encoder.w"}     # a Parameter only on CPU runs
            m.load_state_dict(state, strict=True)
            w = entry["state_dict"].get("encoder.w")
            if w is None:
                w = _tensor_attr(entry.get("encoder"), "w")
            if w is not None and (w.shape != m.encoder_w.shape or not torch.allclose(w.float(), m.encoder_w.cpu(), atol=0, rtol=0)):
                m.encoder_w.copy_(w.float())            # a checkpoint trained with another Fourier matrix keeps its own
            out[out_key] = m.to(device).eval()
        return cfg, (out["dummy"] if embedding_type == "view" else out)

    @staticmethod
    def load_learned_embeds(learned_embeds_path: Union[str, Path]) -> Tuple[List[str], torch.Tensor]:
        """The part of load_learned_embed_in_clip (:232-270) that does not need a tokenizer: tokens and their rows."""
        d = torch.load(learned_embeds_path, map_location="cpu", weights_only=False, pickle_module=_PickleModule)
        return list(d.keys()), torch.stack([v.float() for v in d.values()])
