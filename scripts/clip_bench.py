"""Throughput of the batched conditioning-path encoder: the 16 per-UNet-layer CLIP passes of one train step stacked into
one [16*B, 77, 1024] call, 23 layers, forward + dgrad backward (CUDA-graph replay, CUDA events)."""
import json
import os
import sys

sys.path.insert(0, os.getcwd())
import torch

from view_neti_b200.models.clip_encoder import SD21_TEXT, ClipEncoderEngine, init_state_dict

cfg = SD21_TEXT
B = int(os.environ.get("B", 1))
steps = int(os.environ.get("STEPS", 20))
eng = ClipEncoderEngine(init_state_dict(cfg, 0), cfg)
plan = eng.plan(16 * B, 77)
plan.x_in.normal_()
plan.dy_in.normal_()
plan.run_forward(True)
plan.run_backward()
torch.cuda.synchronize()
res = {}
for what, fn in (("fwd", lambda: plan.run_forward()), ("bwd", lambda: plan.run_backward()),
                 ("fwd+bwd", lambda: (plan.run_forward(), plan.run_backward()))):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    res[what] = e0.elapsed_time(e1) / steps
rows = 16 * B * 77
C, I, L = cfg.hidden_size, cfg.intermediate_size, cfg.num_hidden_layers
gflop_fwd = L * (2.0 * rows * (4 * C * C + 2 * C * I) + 4.0 * 16 * B * cfg.num_attention_heads * 77 * 77 * 64 / 2) / 1e9
print(json.dumps({"nseq": 16 * B, "layers": L, "ms": {k: round(v, 3) for k, v in res.items()},
                  "gflop_fwd": round(gflop_fwd, 1), "tflops_fwd": round(gflop_fwd / res["fwd"], 1),
                  "tflops_fwd_bwd": round(2 * gflop_fwd / res["fwd+bwd"], 1)}))

if os.environ.get("PROFILE", "0") == "1":
    from torch.profiler import ProfilerActivity, profile
    from view_neti_b200 import ops
    ops.set_pdl(False)
    plan.forward(); plan.backward()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        plan.forward()
        plan.backward()
        torch.cuda.synchronize()
    ops.set_pdl(True)
    rows_ = sorted(prof.key_averages(), key=lambda e: -(getattr(e, "self_device_time_total", 0) or 0))
    for e in rows_[:12]:
        print(f"{e.self_device_time_total / 1e3:9.3f} ms  n={e.count:4d}  avg {e.self_device_time_total / max(1, e.count):8.1f} us  {e.key[:90]}")
