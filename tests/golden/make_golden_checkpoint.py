"""Generates tests/golden/mapper_ckpt_{object,view}.pt: mapper checkpoints in the layout the REFERENCE writes
(/root/reference/checkpoint_handler.py:57-97, `mapper-steps-N_{object,view}.pt`), built from the reference's own objects:
its NeTIMapper instances (state_dict + the pickled `encoder` object, a reference class), its RunConfig dataclass tree as
the "cfg" entry (the reference stores pyrallis.encode(cfg), a plain nested dict; pyrallis is not installable here, the
equivalent dict is produced with dataclasses.asdict) and the same top-level keys.  The reference's CheckpointHandler itself
cannot be imported (pyrallis / accelerate), so the dict is assembled here exactly as its save_mapper does.

    python tests/golden/make_golden_checkpoint.py
"""
import dataclasses
import os
import sys
import types
from pathlib import Path

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.modules.setdefault("ipdb", types.ModuleType("ipdb"))
torch.Tensor.cuda = lambda self, *a, **k: self
torch.nn.Module.cuda = lambda self, *a, **k: self
sys.path.insert(0, "/root/reference")
from models.neti_mapper import NeTIMapper  # noqa: E402
from training.config import DataConfig, RunConfig  # noqa: E402
from utils.types import PESigmas  # noqa: E402


def plain(o):
    """pyrallis.encode turns Paths into strings and leaves plain containers: mimic that on the asdict tree."""
    if isinstance(o, dict):
        return {k: plain(v) for k, v in o.items()}
    if isinstance(o, (list, tuple)):
        return [plain(v) for v in o]
    if isinstance(o, Path):
        return str(o)
    return o


def main():
    cfg = RunConfig(data=DataConfig(train_data_dir=Path("data/dtu/scan114")))
    cfg.model.word_embedding_dim = 128
    cfg.model.arch_mlp_hidden_dims = 64
    cfg.model.use_nested_dropout = False
    cfg.model.arch_view_net = 15
    cfg.model.arch_view_disable_tl = False
    cfg.model.target_norm_object = 0.3714
    cfg.model.target_norm_view = 0.4102
    cfg.model.bypass_unconstrained_object = True
    cfg.model.bypass_unconstrained_view = True
    cfg.model.pe_sigmas = PESigmas(sigma_t=0.03, sigma_l=2.0, sigma_theta=0.5, sigma_phi=0.5, sigma_r=0.5, sigma_dtu12=0.5)
    cfg_dict = plain(dataclasses.asdict(cfg))
    common = dict(output_dim=128, arch_mlp_hidden_dims=64, arch_view_net=15, arch_view_disable_tl=False,
                  use_nested_dropout=False, pe_sigmas=cfg.model.pe_sigmas, output_bypass=True, bypass_unconstrained=True,
                  output_bypass_alpha=0.2)
    torch.manual_seed(41)
    objs = {}
    for tok in ("<statue>", "<teapot>"):
        objs[tok] = NeTIMapper(embedding_type="object", norm_scale=torch.tensor(0.3714), placeholder_object_token=tok, **common)
    tokens = ["<view_0_10_1p2>", "<view_10_40_1p2>", "<view_20_70_1p2>"]
    mv = NeTIMapper(embedding_type="view", norm_scale=torch.tensor(0.4102), placeholder_view_tokens=list(tokens),
                    placeholder_view_token_ids=[49410, 49411, 49412], **common)
    # checkpoint_handler.py:63-78 (keys of the lookup are the placeholder token ids)
    obj_ckpt = {"cfg": cfg_dict, "mappers": {}}
    for i, (tok, m) in enumerate(objs.items()):
        obj_ckpt["mappers"][49408 + i] = {"state_dict": m.state_dict(), "encoder": m.encoder, "placeholder_object_token": tok}
    # :80-97
    view_ckpt = {"cfg": cfg_dict, "mappers": {"dummy_key": {"state_dict": mv.state_dict(), "encoder": mv.encoder,
                                                            "placeholder_object_token": "dummy"}}}
    out = os.path.join(ROOT, "tests", "golden")
    torch.save(obj_ckpt, os.path.join(out, "mapper_ckpt_object.pt"))
    torch.save(view_ckpt, os.path.join(out, "mapper_ckpt_view.pt"))
    print({k: os.path.getsize(os.path.join(out, k)) for k in ("mapper_ckpt_object.pt", "mapper_ckpt_view.pt")},
          type(mv.encoder).__module__, type(mv.encoder).__name__, list(mv.state_dict().keys())[:3])


if __name__ == "__main__":
    main()
