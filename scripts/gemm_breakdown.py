"""Per-shape time of every vn_gemm launch in one eager train step (CUPTI durations matched to call order)."""
import collections
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile

from tests.unet_parity import make_inputs
from view_neti_b200 import ops
from view_neti_b200.sd21 import SD21, init_state_dict
from view_neti_b200.unet import UNet2DConditionModel

L = int(sys.argv[1]) if len(sys.argv) > 1 else 64
model = UNet2DConditionModel(init_state_dict(SD21, 0), SD21, "cuda")
plan = model.engine.plan(1, L, L)
lat, t, tgt, ctx = make_inputs(SD21, 1, L, L, seed=1)
plan.latents.copy_(lat); plan.timesteps.copy_(t); plan.target.copy_(tgt)
from view_neti_b200 import ops as _ops
_ops.set_pdl(False)
for _ in range(2):
    plan.train_step()
torch.cuda.synchronize()
calls = []
og, oc = ops.gemm, ops.conv3x3


def g(A, B, D, **k):
    calls.append(("lin", A.numel() // A.shape[-1], B.shape[0], B.shape[1], 0))
    return og(A, B, D, **k)


def c(x, Wk, D, **k):
    calls.append(("conv", x.numel() // x.shape[-1], Wk.shape[0], Wk.shape[1], x.shape[1]))
    return oc(x, Wk, D, **k)


ops.gemm, ops.conv3x3 = g, c
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    plan.train_step()
    torch.cuda.synchronize()
ops.gemm, ops.conv3x3 = og, oc
evs = [e for e in prof.events() if "vn_gemm_kernel" in e.name]
evs.sort(key=lambda e: e.time_range.start)
assert len(evs) == len(calls), (len(evs), len(calls))
agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
for (kind, M, N, K, H), e in zip(calls, evs):
    dur = e.device_time if hasattr(e, "device_time") else e.cuda_time
    key = (kind, M, N, K)
    agg[key][0] += 1; agg[key][1] += dur; agg[key][2] += 2.0 * M * N * K
tot = sum(v[1] for v in agg.values())
print(f"gemm total {tot:.0f} us over {len(calls)} launches")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:60]:
    print(f"{v[1]:8.1f} us n={v[0]:3d} avg {v[1] / v[0]:7.1f} us {v[2] / v[1] / 1e6:7.1f} TFLOP/s  {k}")
by = collections.defaultdict(float)
for k, v in agg.items():
    by[(k[0], k[1])] += v[1]
print({f"{a}-M{m}": round(x) for (a, m), x in sorted(by.items(), key=lambda kv: -kv[1])})
