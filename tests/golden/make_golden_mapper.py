"""Generates tests/golden/neti_mapper.pt by running the REFERENCE's own NeTIMapper
(/root/reference/models/neti_mapper.py, imported unmodified) on the CPU of the build container.

The reference hard-codes `.cuda()` (models/positional_encoding.py:171,186) and imports `ipdb`; here `ipdb` is an empty
stub module and `torch.Tensor.cuda` / `nn.Module.cuda` are patched to identity for the duration of this script, so the
reference's arithmetic runs unchanged on CPU tensors.  Configuration = the paper's model (arch_view_net 15,
neti_mapper.py:601-608): Fourier features of (t, l[, theta, phi]) -> 2 x (Linear 64 + LayerNorm + LeakyReLU) ->
Linear(64, 2*dim) -> split word / bypass -> word = normalize(word) * norm_scale.

    python tests/golden/make_golden_mapper.py
"""
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.modules.setdefault("ipdb", types.ModuleType("ipdb"))
torch.Tensor.cuda = lambda self, *a, **k: self
torch.nn.Module.cuda = lambda self, *a, **k: self
sys.path.insert(0, "/root/reference")
from models.neti_mapper import NeTIMapper  # noqa: E402
from utils.types import PESigmas  # noqa: E402


def run(mapper, t, l, ids, gen):
    mapper.train()
    out = mapper(t, l, ids)
    gw = torch.randn(out.word_embedding.shape, generator=gen)
    gb = torch.randn(out.bypass_output.shape, generator=gen)
    loss = (out.word_embedding * gw).sum() + (out.bypass_output * gb).sum()
    params = {k: v for k, v in mapper.named_parameters() if k != "encoder.w"}
    grads = torch.autograd.grad(loss, list(params.values()))
    return {"word": out.word_embedding.detach(), "bypass": out.bypass_output.detach(), "gw": gw, "gb": gb,
            "grads": dict(zip(params.keys(), [g.detach() for g in grads])),
            "bypass_unconstrained": out.bypass_unconstrained, "output_bypass_alpha": out.output_bypass_alpha}


def main():
    gen = torch.Generator().manual_seed(7)
    sig = PESigmas(sigma_t=0.03, sigma_l=2.0, sigma_theta=0.5, sigma_phi=0.5, sigma_r=0.5, sigma_dtu12=0.5)
    common = dict(output_dim=256, arch_mlp_hidden_dims=64, arch_view_net=15,   # train.yaml:23,34
                  arch_view_disable_tl=False, use_nested_dropout=False,
                  pe_sigmas=sig, output_bypass=True, bypass_unconstrained=True, output_bypass_alpha=0.2)
    t = torch.tensor([0., 17., 500., 999., 250., 731.])
    l = torch.tensor([0., 15., 7., 3., 11., 6.])
    out = {"t": t, "l": l}
    # object mapper
    torch.manual_seed(123)
    mo = NeTIMapper(embedding_type="object", norm_scale=torch.tensor(0.3714), placeholder_object_token="<statue>", **common)
    out["object"] = {"state": {k: v.detach().clone() for k, v in mo.state_dict().items()}, "w": mo.encoder.w.detach().clone(),
                     "norm_scale": 0.3714, **run(mo, t, l, None, gen)}
    # view mapper, (theta, phi) degrees of freedom
    tokens = ["<view_0_10_1p2>", "<view_10_40_1p2>", "<view_20_70_1p2>", "<view_35_100_1p2>"]
    ids = [49408, 49409, 49410, 49411]
    torch.manual_seed(321)
    mv = NeTIMapper(embedding_type="view", norm_scale=torch.tensor(0.4102), placeholder_view_tokens=list(tokens),
                    placeholder_view_token_ids=list(ids), **common)
    vid = torch.tensor([49408, 49411, 49409, 49410, 49410, 49408])
    out["view"] = {"state": {k: v.detach().clone() for k, v in mv.state_dict().items()}, "w": mv.encoder.w.detach().clone(),
                   "norm_scale": 0.4102, "tokens": tokens, "ids": ids, "input_ids": vid, **run(mv, t, l, vid, gen)}
    path = os.path.join(ROOT, "tests", "golden", "neti_mapper.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path), "bytes;", {k: tuple(v.shape) for k, v in out["view"]["state"].items()})
    print("deg_freedom", mv.deg_freedom, "w", tuple(mv.encoder.w.shape), "theta range", mv.theta_min, mv.theta_max)


if __name__ == "__main__":
    main()
