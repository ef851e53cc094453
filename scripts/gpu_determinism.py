"""Run-to-run reproducibility of the forward / backward on one plan (prints relative differences)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from tests.unet_parity import make_inputs, rel
from view_neti_b200.sd21 import SD21, TINY, init_state_dict
from view_neti_b200.unet import UNet2DConditionModel

cfg, L = (SD21, 64) if (len(sys.argv) > 1 and sys.argv[1] == "sd") else (TINY, 16)
model = UNet2DConditionModel(init_state_dict(cfg, 0), cfg, "cuda")
plan = model.engine.plan(1, L, L)
lat, t, tgt, ctx = make_inputs(cfg, 1, L, L, seed=3)
plan.latents.copy_(lat); plan.timesteps.copy_(t)
for i in range(16):
    plan.ctx[0, i].copy_(ctx[f"CONTEXT_TENSOR_{i}"]); plan.ctx[1, i].copy_(ctx[f"CONTEXT_TENSOR_BYPASS_{i}"])
e1 = plan.forward().clone()
tr1 = {k: v.clone() for k, v in plan.trace_f.items()}
e2 = plan.forward().clone()
print("fwd eps run-to-run rel:", rel(e2, e1))
worst = [(rel(plan.trace_f[k], tr1[k]), k) for k in tr1]
print("first block that differs:", next(((k, r) for r, k in worst if r > 0), None), " max:", max(worst))
g = torch.randn_like(e1) * 1e-3
plan.d_eps.copy_(g); d1 = plan.backward().clone()
b1 = {k: v.clone() for k, v in plan.trace_b.items()}
plan.d_eps.copy_(g); d2 = plan.backward().clone()
print("bwd d_ctx run-to-run rel:", rel(d2, d1))
for k in list(b1)[:6] + list(b1)[-4:]:
    print("   dout", k, rel(plan.trace_b[k], b1[k]))
plan.d_eps.copy_(2 * g); d3 = plan.backward().clone()
print("bwd linearity rel:", rel(d3, 2 * d1))
for i in range(16):
    print(f"   layer {i}: k {rel(d2[0, i], d1[0, i]):.2e} v {rel(d2[1, i], d1[1, i]):.2e}   lin k {rel(d3[0, i], 2 * d1[0, i]):.2e}")
