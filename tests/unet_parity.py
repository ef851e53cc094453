"""End-to-end parity of the CUDA path (through the drop-in UNet API) against the fp32 CPU oracle:
eps prediction, fp32 MSE loss and the gradients of all 32 XTI context tensors (coach.py:197-214)."""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn.functional as F

from oracle.unet_sd21 import ResnetBlock2D, Transformer2DModel, UNetOracle, train_step_oracle
from view_neti_b200.sd21 import UNetConfig, init_state_dict


def make_inputs(cfg: UNetConfig, nb: int, h: int, w: int, seed: int = 1, bypass: bool = True):
    g = torch.Generator(device="cpu").manual_seed(seed)
    lat = torch.randn(nb, cfg.in_channels, h, w, generator=g)
    tgt = torch.randn(nb, cfg.out_channels, h, w, generator=g)
    t = torch.randint(0, 1000, (nb,), generator=g)
    ctx: Dict = {"this_idx": 0}
    for i in range(cfg.num_cross_layers):
        ctx[f"CONTEXT_TENSOR_{i}"] = torch.randn(nb, cfg.context_len, cfg.cross_attention_dim, generator=g)
        if bypass:
            ctx[f"CONTEXT_TENSOR_BYPASS_{i}"] = torch.randn(nb, cfg.context_len, cfg.cross_attention_dim, generator=g)
    return lat, t, tgt, ctx


def ctx_to(ctx: Dict, device, requires_grad=True) -> Dict:
    out = {}
    for k, v in ctx.items():
        out[k] = v.detach().clone().to(device).requires_grad_(requires_grad) if torch.is_tensor(v) else v
    return out


def rel(a, b) -> float:
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    if not torch.isfinite(a).all():
        return float("inf")
    return float((a - b).norm() / (b.norm() + 1e-20))


def oracle_step(cfg, sd, lat, t, tgt, ctx, trace=False):
    unet = UNetOracle(cfg)
    unet.load_state_dict(sd)
    acts = {}
    if trace:
        for name, m in unet.named_modules():
            if isinstance(m, (ResnetBlock2D, Transformer2DModel)):
                def hook(mod, inp, out, name=name):
                    acts[name] = out
                    if out.requires_grad:
                        out.retain_grad()
                m.register_forward_hook(hook)
    c = ctx_to(ctx, "cpu")
    eps, loss, grads = train_step_oracle(unet, lat, t, tgt, c)
    keys = [k for k, v in c.items() if torch.is_tensor(v)]
    return eps, loss, dict(zip(keys, grads)), acts


def run_parity(cfg: UNetConfig, nb: int, h: int, w: int, seed: int = 1, bypass: bool = True, trace: bool = False,
               model=None):
    from view_neti_b200.unet import UNet2DConditionModel
    sd = init_state_dict(cfg, 0)
    lat, t, tgt, ctx = make_inputs(cfg, nb, h, w, seed, bypass)
    eps_o, loss_o, grads_o, acts = oracle_step(cfg, sd, lat, t, tgt, ctx, trace)
    if model is None:
        model = UNet2DConditionModel(sd, cfg, "cuda")
    c = ctx_to(ctx, "cuda")
    pred = model(lat.cuda(), t.cuda(), c).sample
    loss = F.mse_loss(pred.float(), tgt.cuda().float(), reduction="mean")
    loss.backward()
    res = {
        "eps_mse": float(((pred.detach().float().cpu() - eps_o) ** 2).mean()),
        "eps_rel": rel(pred, eps_o),
        "loss_rel": abs(float(loss) - float(loss_o)) / abs(float(loss_o)),
        "this_idx": c["this_idx"],
    }
    worst, num, den = 0.0, 0.0, 0.0
    per = {}
    for k, go in grads_o.items():
        gg = c[k].grad
        assert gg is not None, f"no gradient for {k}"
        e = rel(gg, go)
        per[k] = e
        worst = max(worst, e)
        num += float((gg.float().cpu() - go).norm() ** 2)
        den += float(go.norm() ** 2)
    res["grad_worst_rel"] = worst
    res["grad_flat_rel"] = (num / den) ** 0.5
    res["per_grad"] = per
    if trace:
        plan = model.engine.plan(nb, h, w)
        tf, tb = {}, {}
        for name, out in acts.items():
            mine = plan.trace_f[name]
            o = out.detach().permute(0, 2, 3, 1).reshape(mine.shape)
            tf[name] = rel(mine, o)
            if out.grad is not None and name in plan.trace_b:
                tb[name] = rel(plan.trace_b[name], out.grad.permute(0, 2, 3, 1).reshape(plan.trace_b[name].shape))
        res["trace_f"], res["trace_b"] = tf, tb
    return res
