"""CTA-pair (cta_group::2) vs single-CTA tiles of vn_gemm on the SD-2.1 layer shapes: median CUDA-event time per
launch, L2 flushed between launches (weights cold, as inside a step)."""
import os
import sys

sys.path.insert(0, os.getcwd())
import torch

from view_neti_b200 import ops

BF = torch.bfloat16
dev = "cuda"
ws = ops.Workspace(8192, 10240, dev)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
shapes = [("lin", 4096, 320, 320), ("lin", 4096, 960, 320), ("lin", 4096, 2560, 320), ("lin", 4096, 320, 1280),
          ("lin", 1024, 640, 640), ("lin", 1024, 5120, 640), ("lin", 1024, 640, 2560), ("lin", 1024, 1920, 640),
          ("lin", 256, 10240, 1280), ("lin", 256, 1280, 5120), ("lin", 256, 3840, 1280),
          ("conv", (1, 64, 64), 320, 320), ("conv", (1, 64, 64), 320, 960), ("conv", (1, 64, 64), 320, 640),
          ("conv", (1, 64, 64), 640, 640), ("conv", (1, 32, 32), 640, 640), ("conv", (1, 32, 32), 640, 1920),
          ("conv", (1, 32, 32), 1280, 1280), ("conv", (1, 16, 16), 1280, 1280), ("conv", (1, 16, 16), 1280, 2560)]


def med(f, iters=15):
    for _ in range(3):
        f()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


for kind, m, N, K in shapes:
    if kind == "lin":
        A = torch.randn(m, K, device=dev).to(BF); B = torch.randn(N, K, device=dev).to(BF)
        D = torch.empty(m, N, dtype=BF, device=dev); bias = torch.randn(N, device=dev)
        R = torch.randn(m, N, device=dev).to(BF)
        run = lambda **kw: ops.gemm(A, B, D, bias=bias, R=R, ws=ws, **kw)
        fl = 2.0 * m * N * K
        lab = f"lin  M{m} N{N} K{K}"
    else:
        nb, H, W = m
        x = torch.randn(nb, H, W, K, device=dev).to(BF); B = torch.randn(N, 9 * K, device=dev).to(BF)
        D = torch.empty(nb, H, W, N, dtype=BF, device=dev); bias = torch.randn(N, device=dev)
        R = torch.randn(nb, H, W, N, device=dev).to(BF)
        run = lambda **kw: ops.conv3x3(x, B, D, bias=bias, R=R, ws=ws, **kw)
        fl = 2.0 * nb * H * W * N * 9 * K
        lab = f"conv {H}x{W} C{K} N{N}"
    row = [f"{lab:26s}"]
    t_auto = med(lambda: run())
    row.append(f"auto {t_auto:6.1f}us {fl / t_auto / 1e6:6.0f}TF")
    for bn in (128, 256):
        t1 = med(lambda: run(force_bn=bn, force_split=1 + 512))
        t2 = med(lambda: run(force_bn=bn, force_split=1 + 256))
        row.append(f"| bn{bn}: 1cta {t1:6.1f}  pair {t2:6.1f}")
    print("  ".join(row), flush=True)
