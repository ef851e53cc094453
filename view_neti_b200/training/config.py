"""RunConfig — the reference's training configuration (reference training/config.py:11-293) with the same section and
field names, so its yaml files (input_configs/*.yaml) and dotted command-line overrides (`--optim.max_train_steps 10`,
README.md:40-44) load unchanged.  pyrallis is not available offline; `load_config` is a small yaml + dotted-override
loader over plain dataclasses.

Field tables are data: (name, default) per section.  Validation follows config.py:142-178 (pe_sigmas dict -> PESigmas with
the experiment keys applied) and :268-293 (micro-batch <= 3, mode-3 requirements, unique object tokens, modes 4/5 need a
pretrained view mapper).
"""
from __future__ import annotations

import dataclasses
from dataclasses import dataclass, field, make_dataclass
from pathlib import Path
from typing import Any, Dict, List, Optional, Sequence

from ..utils.types import PESigmas

_MISSING = dataclasses.MISSING


def _section(name: str, table: Sequence, post=None):
    fields = []
    for key, default in table:
        if default is _MISSING:
            fields.append((key, Any))
        elif isinstance(default, (list, dict)):
            fields.append((key, Any, field(default_factory=lambda d=default: type(d)(d))))
        else:
            fields.append((key, Any, field(default=default)))
    ns = {"__post_init__": post} if post else {}
    cls = make_dataclass(name, fields, namespace=ns)
    cls.__module__ = __name__
    return cls


LogConfig = _section("LogConfig", [
    ("exp_name", ""), ("overwrite_ok", False), ("exp_dir", Path("./outputs")), ("save_steps", 1000),
    ("logging_dir", Path("logs")), ("report_to", "all"), ("checkpoints_total_limit", None), ("save_dataset_images", True),
])

DataConfig = _section("DataConfig", [
    ("train_data_dir", "synthetic"), ("train_data_subsets", None), ("placeholder_object_token", "<>"),
    ("super_category_object_token", "object"), ("super_category_view_token", "view"),
    ("placeholder_object_tokens", None), ("super_category_object_tokens", None), ("fixed_object_token_or_path", None),
    ("dataloader_num_workers", 8), ("repeats", 100), ("resolution", 512), ("dtu_preprocess_key", 1), ("center_crop", False),
    ("flip_p", 0.5), ("placeholder_view_tokens", None), ("caption_strategy", 0), ("camera_representation", "spherical"),
    ("dtu_lighting", 3), ("dtu_subset", -2), ("augmentation_key", 0),
])

_SIGMA_DTU12_BY_KEY = {1: 1.0, 2: 0.5, 3: 0.25, 4: 0.75, 5: 0.1}
_SIGMA_T_BY_KEY = {0: 0.03, 1: 0.06, 2: 0.2, 3: 0.5}
_SIGMA_L_BY_KEY = {0: 2.0, 1: 4.0}


def _model_post(self):
    """config.py:142-178.  Note the reference fills sigma_theta and sigma_r from the 'sigma_phi' key (:147-149)."""
    s = self.pe_sigmas
    if s is not None and not isinstance(s, PESigmas):
        phi = s.get("sigma_phi", 1.0)
        s = PESigmas(sigma_t=s["sigma_t"], sigma_l=s["sigma_l"], sigma_theta=phi, sigma_phi=phi, sigma_r=phi,
                     sigma_dtu12=s.get("sigma_dtu12", 2.0))
        if self.pe_sigma_exp_key in _SIGMA_DTU12_BY_KEY:
            s.sigma_dtu12 = _SIGMA_DTU12_BY_KEY[self.pe_sigma_exp_key]
        if self.pe_t_exp_key not in _SIGMA_T_BY_KEY or self.pe_l_exp_key not in _SIGMA_L_BY_KEY:
            raise ValueError("model.pe_t_exp_key must be 0..3 and model.pe_l_exp_key 0..1")
        s.sigma_t = _SIGMA_T_BY_KEY[self.pe_t_exp_key]
        s.sigma_l = _SIGMA_L_BY_KEY[self.pe_l_exp_key]
        self.pe_sigmas = s
    if self.pretrained_view_mapper is not None:
        self.pretrained_view_mapper = Path(self.pretrained_view_mapper)


ModelConfig = _section("ModelConfig", [
    ("pretrained_model_name_or_path", "CompVis/stable-diffusion-v1-4"), ("pretrained_view_mapper", None),
    ("pretrained_view_mapper_key", None), ("word_embedding_dim", 768), ("arch_mlp_hidden_dims", 128),
    ("use_nested_dropout", True), ("nested_dropout_prob", 0.5), ("normalize_object_mapper_output", True),
    ("normalize_view_mapper_output", False), ("target_norm_object", None), ("target_norm_view", None),
    ("use_positional_encoding_object", 1), ("use_positional_encoding_view", 1),
    ("pe_sigmas", {"sigma_t": 0.03, "sigma_l": 2.0, "sigma_theta": 1.0, "sigma_phi": 1.0, "sigma_r": 1.0, "sigma_dtu12": 2.0}),
    ("pe_sigma_exp_key", 0), ("pe_t_exp_key", 0), ("pe_l_exp_key", 0), ("pe_sigmas_view", {"sigma_phi": 1.0}),
    ("num_pe_time_anchors", 10), ("output_bypass_object", True), ("output_bypass_view", True), ("revision", None),
    ("mapper_checkpoint_path", None), ("arch_view_net", 0), ("arch_view_mix_streams", 0), ("arch_view_disable_tl", True),
    ("original_ti", False), ("bypass_unconstrained_object", False), ("bypass_unconstrained_view", False),
    ("output_bypass_alpha_view", 0.2), ("output_bypass_alpha_object", 0.2),
], post=_model_post)


def _eval_post(self):
    if self.validation_seeds is None:
        self.validation_seeds = list(range(self.num_validation_images))
    assert len(self.validation_seeds) == self.num_validation_images, \
        "Length of validation_seeds should equal num_validation_images"


EvalConfig = _section("EvalConfig", [
    ("validation_prompts", []), ("validation_view_tokens", None), ("num_validation_images", 3), ("validation_seeds", [0, 1, 2]),
    ("validation_steps", 250), ("num_denoising_steps", 30), ("dtu_upsample_key", 1), ("eval_placeholder_object_tokens", None),
], post=_eval_post)

OptimConfig = _section("OptimConfig", [
    ("max_train_steps", 1000), ("learning_rate", 1e-3), ("scale_lr", True), ("train_batch_size", 3),
    ("gradient_checkpointing", False), ("gradient_accumulation_steps", 3), ("seed", None), ("lr_scheduler", "constant"),
    ("lr_warmup_steps", 0), ("adam_beta1", 0.9), ("adam_beta2", 0.999), ("adam_weight_decay", 1e-2), ("adam_epsilon", 1e-8),
    ("mixed_precision", "no"), ("allow_tf32", False),
])

# training/pretrained_models.py:1-5 maps small integer keys to checkpoint paths of pretrained view mappers
lookup_pretrained_models: Dict[str, str] = {}


@dataclass
class RunConfig:
    """learnable_mode: 0 object only | 1 view only | 2 view + object | 3 view + several objects | 4 view (pretrained) +
    object | 5 view (pretrained, frozen) + object   (config.py:252-262)"""
    learnable_mode: int = 0
    debug: bool = False
    seed: int = 0
    log: Any = field(default_factory=LogConfig)
    data: Any = field(default_factory=DataConfig)
    model: Any = field(default_factory=ModelConfig)
    eval: Any = field(default_factory=EvalConfig)
    optim: Any = field(default_factory=OptimConfig)

    def __post_init__(self):
        if self.optim.train_batch_size > 3:                                      # config.py:269-271
            raise ValueError("batch size should be 3 and so should grad accumulation")
        if self.learnable_mode == 3:
            assert self.data.dataloader_num_workers == 0, "can't support multiple workers right now for learnable mode 3"
            assert self.data.super_category_object_tokens is not None
            ev = self.eval.eval_placeholder_object_tokens
            if ev is not None:
                assert all(d in self.data.placeholder_object_tokens for d in ev), \
                    "eval.eval_placeholder_tokens not in data.placeholder_object_tokens"
        toks = self.data.placeholder_object_tokens
        if toks is not None:
            assert len(toks) == len(set(toks)), "cfg.data.placeholder_object_tokens must be unique strings"
        if self.learnable_mode in (4, 5):
            m = self.model
            assert m.pretrained_view_mapper or m.pretrained_view_mapper_key
            if m.pretrained_view_mapper_key:
                m.pretrained_view_mapper = Path(lookup_pretrained_models[str(m.pretrained_view_mapper_key)])


_SECTIONS = {"log": LogConfig, "data": DataConfig, "model": ModelConfig, "eval": EvalConfig, "optim": OptimConfig}
_PATHS = {("log", "exp_dir"), ("log", "logging_dir")}


def _coerce(text: str):
    import yaml
    return yaml.safe_load(text)


def from_dict(d: Dict[str, Any]) -> RunConfig:
    d = dict(d or {})
    kw: Dict[str, Any] = {}
    for name, cls in _SECTIONS.items():
        sec = dict(d.pop(name, None) or {})
        known = {f.name for f in dataclasses.fields(cls)}
        unknown = set(sec) - known
        if unknown:
            raise ValueError(f"unknown field(s) in section '{name}': {sorted(unknown)}")
        for k in list(sec):
            if (name, k) in _PATHS and sec[k] is not None:
                sec[k] = Path(sec[k])
        kw[name] = cls(**sec)
    unknown = set(d) - {"learnable_mode", "debug", "seed"}
    if unknown:
        raise ValueError(f"unknown top-level field(s): {sorted(unknown)}")
    return RunConfig(**d, **kw)


def load_config(config_path: Optional[str] = None, overrides: Sequence[str] = ()) -> RunConfig:
    """yaml file + dotted overrides (`--section.field value` or `--section.field=value`, bare `--flag` = true)."""
    import yaml
    d: Dict[str, Any] = {}
    if config_path:
        with open(config_path) as f:
            d = yaml.safe_load(f) or {}
    toks = list(overrides)
    i = 0
    while i < len(toks):
        tok = toks[i]
        if not tok.startswith("--"):
            raise ValueError(f"expected --name, got '{tok}'")
        key, eq, val = tok[2:].partition("=")
        i += 1
        if not eq:
            if i < len(toks) and not toks[i].startswith("--"):
                val = toks[i]
                i += 1
            else:
                val = "true"                       # bare flag (README.md:43 `--log.overwrite_ok`)
        parts = key.split(".")
        node = d
        for p in parts[:-1]:
            node = node.setdefault(p, {})
        node[parts[-1]] = _coerce(val)
    return from_dict(d)


def to_dict(cfg: RunConfig) -> Dict[str, Any]:
    """Plain nested dict (what the reference stores under 'cfg' in a mapper checkpoint via pyrallis.encode)."""
    def enc(v):
        if dataclasses.is_dataclass(v):
            return {f.name: enc(getattr(v, f.name)) for f in dataclasses.fields(v)}
        if isinstance(v, Path):
            return str(v)
        if isinstance(v, (list, tuple)):
            return [enc(x) for x in v]
        if isinstance(v, dict):
            return {k: enc(x) for k, x in v.items()}
        return v
    return enc(cfg)


def parse_args(argv: Sequence[str]) -> RunConfig:
    """`python scripts/train.py --config_path input_configs/train.yaml --optim.max_train_steps 10` (pyrallis.wrap syntax)."""
    argv = list(argv)
    path = None
    rest: List[str] = []
    i = 0
    while i < len(argv):
        if argv[i] == "--config_path":
            path = argv[i + 1]
            i += 2
        elif argv[i].startswith("--config_path="):
            path = argv[i].split("=", 1)[1]
            i += 1
        else:
            rest.append(argv[i])
            i += 1
    return load_config(path, rest)
