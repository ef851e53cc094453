"""Where the COMPLETE train step (Coach.train_step: conditioning path + UNet + mapper gradients + AdamW) spends its time:
CUPTI timeline of one step after warm-up - device busy / idle time, kernels grouped by family (ours vs torch glue), and the
host-side span of the step.  python scripts/full_step_profile.py"""
import collections
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile

from view_neti_b200.sd21 import SD21, init_state_dict
from view_neti_b200.training.coach import Coach
from view_neti_b200.training.synthetic import build_conditioning, synthetic_prompt
from view_neti_b200.unet import UNet2DConditionModel

dev = "cuda"
cond = build_conditioning(dev)
unet = UNet2DConditionModel(init_state_dict(SD21, 0), SD21, dev)
coach = Coach(cfg=None, unet=unet, conditioning=cond, optimizer=torch.optim.AdamW(cond.parameters(), lr=1e-3),
              generator=torch.Generator(device=dev).manual_seed(1))
batch = synthetic_prompt(1, dev)
latents = torch.randn(1, 4, 64, 64, device=dev)
for _ in range(6):
    coach.train_step(latents, batch)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    coach.train_step(latents, batch)
e1.record()
torch.cuda.synchronize()
print(f"full step {e0.elapsed_time(e1) / 20:.3f} ms")
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    coach.train_step(latents, batch)
    torch.cuda.synchronize()
ev = sorted([(e.time_range.start, e.time_range.end, e.name) for e in prof.events()
             if e.device_type == torch.autograd.DeviceType.CUDA], key=lambda t: t[0])
t0, busy, idle = ev[0][0], ev[0][0], 0.0
fam = collections.defaultdict(lambda: [0, 0.0, 0.0])
gaps = []
for s, en, n in ev:
    if s > busy:
        idle += s - busy
        gaps.append((s - busy, n[:60], round(s - t0, 1)))
    ours = any(k in n for k in ("vn_", "attn_", "gn_", "ln_kernel", "geglu", "gelu", "mapper", "adamw", "conv_", "im2col", "col2im",
                                "upsample", "gemv", "mse_", "cast_", "timestep", "softmax_rows", "nhwc", "seq_attn", "memset"))
    key = ("ours: " if ours else "torch: ") + n.replace("void ", "").replace("(anonymous namespace)::", "").split("(")[0].split("<")[0][:48]
    add = max(0.0, en - max(s, busy))
    fam[key][0] += 1; fam[key][1] += en - s; fam[key][2] += add
    busy = max(busy, en)
print(f"device span {(busy - t0) / 1e3:.3f} ms, idle {idle / 1e3:.3f} ms, kernels {len(ev)}")
for k, v in sorted(fam.items(), key=lambda kv: -kv[1][2])[:40]:
    print(f"  {k:60s} n={v[0]:4d} excl={v[2] / 1e3:7.3f} ms")
tt = sum(v[2] for k, v in fam.items() if k.startswith("torch"))
print(f"torch glue kernels: {sum(v[0] for k, v in fam.items() if k.startswith('torch'))} launches, {tt / 1e3:.3f} ms exclusive")
gaps.sort(reverse=True)
print("largest idle gaps (us, next kernel, at us):", gaps[:12])
