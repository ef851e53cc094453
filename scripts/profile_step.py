"""One warm-up + one eager train step (64x64 latents) for ncu launch lists / captures."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from tests.unet_parity import make_inputs
from view_neti_b200.sd21 import SD21, init_state_dict
from view_neti_b200.unet import UNet2DConditionModel

L = int(sys.argv[1]) if len(sys.argv) > 1 else 64
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
model = UNet2DConditionModel(init_state_dict(SD21, 0), SD21, "cuda")
plan = model.engine.plan(1, L, L)
lat, t, tgt, ctx = make_inputs(SD21, 1, L, L, seed=1)
plan.latents.copy_(lat); plan.timesteps.copy_(t); plan.target.copy_(tgt)
for i in range(16):
    plan.ctx[0, i].copy_(ctx[f"CONTEXT_TENSOR_{i}"]); plan.ctx[1, i].copy_(ctx[f"CONTEXT_TENSOR_BYPASS_{i}"])
for _ in range(steps):
    plan.train_step()
torch.cuda.synchronize()
print("loss", float(plan.loss))
