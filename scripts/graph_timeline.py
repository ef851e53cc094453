"""Timeline of ONE CUDA-graph replay of the train step (64x64 latents, PDL on, as benchmarked): CUPTI start / end of
every kernel node, written to gpurun_out/graph_timeline.json and summarised:

  * per kernel family: launches, sum of durations, sum of EXCLUSIVE time (wall time in which this kernel was the only
    one running on its stream position: end_i - max(start_i, end_{i-1}) along the main chain),
  * idle gaps on the device (no kernel running),
  * the 40 most expensive chain links.

    python scripts/graph_timeline.py [latent] [what=train|fwd]
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile

from tests.unet_parity import make_inputs
from view_neti_b200.sd21 import SD21, init_state_dict
from view_neti_b200.unet import UNet2DConditionModel

L = int(sys.argv[1]) if len(sys.argv) > 1 else 64
what = sys.argv[2] if len(sys.argv) > 2 else "train"
nb = int(sys.argv[3]) if len(sys.argv) > 3 else 1
model = UNet2DConditionModel(init_state_dict(SD21, 0), SD21, "cuda")
plan = model.engine.plan(nb, L, L)
lat, t, tgt, ctx = make_inputs(SD21, nb, L, L, seed=1)
plan.latents.copy_(lat); plan.timesteps.copy_(t); plan.target.copy_(tgt)
for i in range(16):
    plan.ctx[0, i].copy_(ctx[f"CONTEXT_TENSOR_{i}"]); plan.ctx[1, i].copy_(ctx[f"CONTEXT_TENSOR_BYPASS_{i}"])
g = plan.capture(what)
for _ in range(5):
    g.replay()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    g.replay()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    g.replay()
    torch.cuda.synchronize()
ev = []
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range is not None:
        ev.append((e.time_range.start, e.time_range.end, e.name, getattr(e, "stream", None)))
ev.sort()
t0 = ev[0][0]
rows = [{"name": n, "start_us": round(s - t0, 3), "dur_us": round(en - s, 3)} for s, en, n, _ in ev]
os.makedirs("gpurun_out", exist_ok=True)
with open(f"gpurun_out/graph_timeline_{what}_{L}.json", "w") as f:
    json.dump({"ms_per_replay": ms, "events": rows}, f)


def fam(n):
    for k, v in (("vn_gemm_kernel", "gemm"), ("attn_", "attn"), ("gn_", "groupnorm"), ("ln_kernel", "layernorm"),
                 ("geglu", "geglu")):
        if k in n:
            return v
    return "other"


span = ev[-1][1] - t0
busy_end, idle = t0, 0.0
excl, tot, cnt = {}, {}, {}
links = []
for s, en, n, _ in ev:
    f_ = fam(n)
    tot[f_] = tot.get(f_, 0.0) + (en - s)
    cnt[f_] = cnt.get(f_, 0) + 1
    if s > busy_end:
        idle += s - busy_end
    add = max(0.0, en - max(s, busy_end))
    excl[f_] = excl.get(f_, 0.0) + add
    links.append((add, en - s, n[:70], round(s - t0, 1)))
    busy_end = max(busy_end, en)
print(f"replay {ms:.3f} ms (events) ; profiled span {span / 1e3:.3f} ms ; {len(ev)} kernels ; idle {idle / 1e3:.3f} ms")
print(f"{'family':10s} {'n':>5s} {'sum dur ms':>11s} {'exclusive ms':>13s}")
for k in sorted(tot, key=lambda k: -excl[k]):
    print(f"{k:10s} {cnt[k]:5d} {tot[k] / 1e3:11.3f} {excl[k] / 1e3:13.3f}")
links.sort(reverse=True)
print("top chain links (exclusive us, duration us, kernel, start us):")
for a, d, n, s in links[:40]:
    print(f"  {a:7.1f} {d:7.1f}  {n}  @{s}")
# histogram of exclusive time per launch
import collections
h = collections.Counter()
for a, d, n, s in links:
    h[min(int(a), 30)] += 1
print("exclusive-us histogram:", sorted(h.items()))
