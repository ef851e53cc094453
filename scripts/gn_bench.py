"""Micro-benchmark of the one-launch GroupNorm (forward, backward) against the two-kernel form on the SD-2.1 shapes."""
import os
import sys

sys.path.insert(0, os.getcwd())
import torch

from view_neti_b200 import ops

BF = torch.bfloat16
dev = "cuda"
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 20
shapes = [(1, 4096, 320), (1, 4096, 640), (1, 4096, 960), (1, 1024, 640), (1, 1024, 1280), (1, 1024, 1920),
          (1, 256, 1280), (1, 256, 2560), (1, 64, 1280), (1, 64, 2560)]
only = os.environ.get("ONLY_IDX")
for idx, (nb, hw, C) in enumerate(shapes):
    if only is not None and idx != int(only):
        continue
    x = torch.randn(nb, hw, C, device=dev).to(BF)
    dy = torch.randn(nb, hw, C, device=dev).to(BF)
    y = torch.empty_like(x)
    gamma, beta = torch.ones(C, device=dev), torch.zeros(C, device=dev)
    stats = torch.zeros(nb, 32, 2, dtype=torch.float64, device=dev)
    red = torch.zeros(nb, 32, 2, dtype=torch.float64, device=dev)
    part = torch.empty(2, ops.groupnorm_partial_floats(nb), device=dev)
    row = [f"groupnorm nb{nb} hw{hw} C{C}:"]
    for fused in (1, 0):
        ops.set_groupnorm_fused(bool(fused))
        for name in ("fwd", "bwd"):
            ts = []
            for _ in range(iters + 3):
                ops.memset(part, 0xFF); stats.zero_() if name == "fwd" else red.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                if name == "fwd":
                    ops.groupnorm_fwd(x, gamma, beta, 1e-5, True, y, nb, hw, 32, stats, part[0])
                else:
                    ops.groupnorm_bwd_fused(x, dy, stats, red, part[1], gamma, beta, 1e-5, True, y, nb, hw, 32)
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1) * 1e3)
            ts = sorted(ts[3:])
            row.append(f"{'fused' if fused else '2-kernel'} {name} {ts[len(ts) // 2]:6.1f} us")
    ops.set_groupnorm_fused(True)
    print("  ".join(row), flush=True)
