"""ms per captured train step (64x64 latents, SD-2.1 widths) of the tree in the CURRENT DIRECTORY, plus per-family
kernel time (CUPTI, PDL off).  Run from two checkouts on the same box to compare library versions like for like:
    python scripts/step_time.py ; (cd _base && python ../scripts/step_time.py)"""
import json
import os
import sys

sys.path.insert(0, os.getcwd())
import torch

from tests.unet_parity import make_inputs
from view_neti_b200 import ops
from view_neti_b200.sd21 import SD21, init_state_dict
from view_neti_b200.unet import UNet2DConditionModel

L = int(os.environ.get("LATENT", 64))
steps = int(os.environ.get("STEPS", 40))
model = UNet2DConditionModel(init_state_dict(SD21, 0), SD21, "cuda")
plan = model.engine.plan(1, L, L)
lat, t, tgt, ctx = make_inputs(SD21, 1, L, L, seed=1)
plan.latents.copy_(lat); plan.timesteps.copy_(t); plan.target.copy_(tgt)
for i in range(SD21.num_cross_layers):
    plan.ctx[0, i].copy_(ctx[f"CONTEXT_TENSOR_{i}"]); plan.ctx[1, i].copy_(ctx[f"CONTEXT_TENSOR_BYPASS_{i}"])
g = plan.capture("train")
res = []
for rep in range(3):
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    res.append(e0.elapsed_time(e1) / steps)
row = {"tree": os.getcwd(), "ms_per_step": [round(r, 4) for r in res], "launches": plan.launches["train"],
       "loss": float(plan.loss)}
if os.environ.get("CATS", "1") == "1":
    from bench import profile_categories
    ms, n, _ = profile_categories(plan)
    row["kernel_ms"] = {k: round(v, 3) for k, v in ms.items()}
print(json.dumps(row), flush=True)
