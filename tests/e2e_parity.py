"""Loss -> mapper-parameter gradient parity of the COMPLETE differentiable path of one train step
(reference training/coach.py:186-214 + 276-311):

    NeTI mappers -> placeholder rows of the token embeddings -> 23-layer CLIP text encoder (16 stacked passes) -> bypass
    injection -> final LayerNorm -> XTI context dict -> frozen SD-2.1 UNet -> fp32 MSE -> backward into the mappers

CUDA path (view_neti_b200) against the fp32 CPU oracles (oracle/neti_mapper.py, neti_conditioning.py, clip_encoder.py,
unet_sd21.py) on identical (latents, timestep, target, prompt) inputs and identical weights.  north_star contract:
eps-MSE < 1e-3, relative L2 error of the flat mapper gradient < 1e-2.
"""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn.functional as F

from oracle.clip_encoder import encoder_forward
from oracle.clip_encoder import init_state_dict as clip_init
from oracle.neti_conditioning import conditioning_forward
from oracle.neti_mapper import encode_inputs, mapper_forward
from oracle.unet_sd21 import UNetOracle
from view_neti_b200.sd21 import UNetConfig, init_state_dict

OBJ_ID = 49408
VIEW_TOKENS = ["<view_0_10_1p2>", "<view_10_40_1p2>", "<view_20_70_1p2>", "<view_35_100_1p2>"]
VIEW_IDS = [49409, 49410, 49411, 49412]
NS_OBJ, NS_VIEW = 0.3714, 0.4102


def _rel(a, b) -> float:
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    if not torch.isfinite(a).all():
        return float("inf")
    return float((a - b).norm() / (b.norm() + 1e-20))


def _cuda_side(cfg, sd_unet, sd_clip, clip_shape, tok, pos, fln, mo, mv, unet, lat, tgt, t, ids, ph_o, view):
    from view_neti_b200.models.clip_encoder import CLIPEncoder, ClipEncoderConfig
    from view_neti_b200.models.neti_conditioning import NeTIConditioning
    from view_neti_b200.unet import UNet2DConditionModel
    heads, layers, inter = clip_shape
    enc = CLIPEncoder(sd_clip, ClipEncoderConfig(hidden_size=cfg.cross_attention_dim, num_attention_heads=heads,
                                                 num_hidden_layers=layers, intermediate_size=inter), "cuda")
    cond = NeTIConditioning(tok, pos, fln, enc, {OBJ_ID: mo.cuda()}, mv.cuda())
    if unet is None:
        unet = UNet2DConditionModel(sd_unet, cfg, "cuda")
    hs = cond(input_ids=ids.cuda(), timesteps=t.cuda(), input_ids_placeholder_object=ph_o, input_ids_placeholder_view=view)
    pred = unet(lat.cuda(), t.cuda(), hs).sample
    loss = F.mse_loss(pred.float(), tgt.cuda().float(), reduction="mean")
    loss.backward()
    flat = torch.cat([p.grad.reshape(-1).float().cpu() for m in (mo, mv) for _, p in m.named_parameters()])
    return flat, pred, loss, hs


def run_e2e(cfg: UNetConfig, clip_layers: int, clip_heads: int, clip_inter: int, nb: int, h: int, w: int, seed: int = 1,
            unet=None, bypass_unconstrained: bool = True) -> Dict:
    from view_neti_b200.models.neti_mapper import NeTIMapper
    from view_neti_b200.utils.types import PESigmas
    C, L = cfg.cross_attention_dim, cfg.context_len
    g = torch.Generator().manual_seed(seed)
    # ---- weights shared by both sides ----
    sd_unet = init_state_dict(cfg, 0)
    sd_clip = clip_init(C, clip_heads, clip_layers, clip_inter, seed=7)
    tok = torch.randn(OBJ_ID + 8, C, generator=g) * 0.02
    pos = torch.randn(L, C, generator=g) * 0.01
    fln = (1 + 0.1 * torch.randn(C, generator=g), 0.05 * torch.randn(C, generator=g))
    sig = PESigmas(sigma_t=0.03, sigma_l=2.0, sigma_theta=0.5, sigma_phi=0.5, sigma_r=0.5, sigma_dtu12=0.5)
    kw = dict(output_dim=C, arch_mlp_hidden_dims=64, arch_view_net=15, arch_view_disable_tl=False, use_nested_dropout=False,
              pe_sigmas=sig, output_bypass=True, bypass_unconstrained=bypass_unconstrained, output_bypass_alpha=0.2)
    with torch.random.fork_rng(devices=[]):
        torch.manual_seed(seed + 100)
        mo = NeTIMapper(embedding_type="object", norm_scale=torch.tensor(NS_OBJ), placeholder_object_token="<statue>", **kw)
        mv = NeTIMapper(embedding_type="view", norm_scale=torch.tensor(NS_VIEW), placeholder_view_tokens=list(VIEW_TOKENS),
                        placeholder_view_token_ids=list(VIEW_IDS), **kw)
    st_o = {k: v.detach().clone() for k, v in mo.state_dict().items()}
    st_v = {k: v.detach().clone() for k, v in mv.state_dict().items()}
    w_o, w_v = mo.encoder_w.clone(), mv.encoder_w.clone()
    # ---- inputs ----
    lat = torch.randn(nb, cfg.in_channels, h, w, generator=g)
    tgt = torch.randn(nb, cfg.out_channels, h, w, generator=g)
    t = torch.randint(0, 1000, (nb,), generator=g)
    ids = torch.randint(1000, 40000, (nb, L), generator=g)
    view = torch.tensor([VIEW_IDS[(i + 1) % len(VIEW_IDS)] for i in range(nb)])
    ids[:, 0], ids[:, 5], ids[:, 3] = 49406, OBJ_ID, view
    ph_o = torch.full((nb,), OBJ_ID)

    # ---- CUDA path (skipped on a CPU-only box: tests/test_oracle_cpu.py dry-runs the oracle half) ----
    names = [n for n, _ in mo.named_parameters()]
    flat_gpu = pred = loss = hs = None
    if torch.cuda.is_available():
        flat_gpu, pred, loss, hs = _cuda_side(cfg, sd_unet, sd_clip, (clip_heads, clip_layers, clip_inter), tok, pos, fln, mo, mv,
                                              unet, lat, tgt, t, ids, ph_o, view)

    # ---- oracle ----
    nl = 16
    so = {k: v.clone().requires_grad_(True) for k, v in st_o.items()}
    sv = {k: v.clone().requires_grad_(True) for k, v in st_v.items()}
    t_rep = t.float().repeat(nl)
    l_rep = torch.arange(nl).repeat_interleave(nb).float()
    table = {i: [float(s.replace("p", ".")) for s in tk[6:-1].split("_")] for tk, i in zip(VIEW_TOKENS, VIEW_IDS)}
    allp = torch.tensor(list(table.values()))
    vp = torch.tensor([table[int(i)][:2] for i in view]).repeat(nl, 1)
    wo, bo = mapper_forward(so, w_o, encode_inputs(t_rep, l_rep), NS_OBJ)
    wv, bv = mapper_forward(sv, w_v, encode_inputs(t_rep, l_rep, vp, allp.min(0).values.tolist(), allp.max(0).values.tolist()),
                            NS_VIEW)
    obj_out = torch.stack([wo.view(nl, nb, C), bo.view(nl, nb, C)], dim=1)                 # [16, 2, B, C]
    view_out = torch.stack([wv.view(nl, nb, C), bv.view(nl, nb, C)], dim=1)
    hs_o = conditioning_forward(ids, ph_o, view, tok, pos, fln, lambda x: encoder_forward(sd_clip, x, clip_heads, clip_layers),
                                obj_out, view_out, bypass_unconstrained, 0.2)
    oracle = UNetOracle(cfg)
    oracle.load_state_dict(sd_unet)
    eps_o = oracle(lat, t, hs_o).sample
    loss_o = F.mse_loss(eps_o.float(), tgt.float(), reduction="mean")
    leaves = [so[n] for n in names] + [sv[n] for n in names]
    grads = torch.autograd.grad(loss_o, leaves)
    flat_o = torch.cat([x.reshape(-1) for x in grads])

    if flat_gpu is None:
        return {"oracle_loss": float(loss_o.detach()), "mapper_grad_norm": float(flat_o.norm()), "n": flat_o.numel()}
    res = {"eps_mse": float(((pred.detach().float().cpu() - eps_o.detach()) ** 2).mean()),
           "loss_rel": abs(float(loss) - float(loss_o)) / abs(float(loss_o)),
           "mapper_grad_flat_rel": _rel(flat_gpu, flat_o),
           "mapper_grad_norm": float(flat_o.norm()),
           "ctx_rel": max(_rel(hs[k], hs_o[k]) for k in hs if k != "this_idx")}
    n_o = sum(x.numel() for x in grads[:len(names)])
    res["object_grad_rel"] = _rel(flat_gpu[:n_o], flat_o[:n_o])
    res["view_grad_rel"] = _rel(flat_gpu[n_o:], flat_o[n_o:])
    off, per = 0, {}
    for which, gs in (("object", grads[:len(names)]), ("view", grads[len(names):])):
        for n, x in zip(names, gs):
            per[f"{which}.{n}"] = _rel(flat_gpu[off:off + x.numel()], x.reshape(-1))
            off += x.numel()
    res["per_param"] = per
    return res
