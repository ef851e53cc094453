"""Per-kernel GPU time of one eager train step (CUPTI via torch.profiler; no replay overhead)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile

from tests.unet_parity import make_inputs
from view_neti_b200.sd21 import SD21, init_state_dict
from view_neti_b200.unet import UNet2DConditionModel

L = int(sys.argv[1]) if len(sys.argv) > 1 else 64
nb = int(sys.argv[2]) if len(sys.argv) > 2 else 1
model = UNet2DConditionModel(init_state_dict(SD21, 0), SD21, "cuda")
plan = model.engine.plan(nb, L, L)
lat, t, tgt, ctx = make_inputs(SD21, nb, L, L, seed=1)
plan.latents.copy_(lat); plan.timesteps.copy_(t); plan.target.copy_(tgt)
for i in range(16):
    plan.ctx[0, i].copy_(ctx[f"CONTEXT_TENSOR_{i}"]); plan.ctx[1, i].copy_(ctx[f"CONTEXT_TENSOR_BYPASS_{i}"])
from view_neti_b200 import ops as _ops
_ops.set_pdl(False)
for _ in range(2):
    plan.train_step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    plan.train_step()
    torch.cuda.synchronize()
rows = []
for e in prof.key_averages():
    if e.device_type is not None and "cuda" in str(e.device_type).lower() or getattr(e, "self_device_time_total", 0) > 0:
        dt = getattr(e, "self_device_time_total", None) or getattr(e, "self_cuda_time_total", 0)
        if dt > 0:
            rows.append((dt, e.count, e.key))
rows.sort(reverse=True)
tot = sum(r[0] for r in rows)
print(f"total kernel time {tot:.0f} us")
for dt, n, k in rows[:40]:
    k = k.replace("void ", "").replace("(anonymous namespace)::", "")
    print(f"{dt:10.1f} us {n:5d} avg {dt / n:8.1f} {100 * dt / tot:5.1f}%  {k[:90]}")
