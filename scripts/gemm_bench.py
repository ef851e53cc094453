"""Micro-benchmark of vn_gemm on the SD-2.1 layer shapes (CUDA events, L2 flushed between launches is NOT done:
weights of one layer fit in L2, as they do not inside a full step - treat numbers as upper bounds)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from view_neti_b200 import ops

BF = torch.bfloat16
dev = "cuda"
ws = ops.Workspace(8192, 10240, dev)
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 20
only = sys.argv[2] if len(sys.argv) > 2 else ""

# (kind, M or (nb,H,W), N, K or C)
shapes = [
    ("lin", 4096, 320, 320), ("lin", 4096, 960, 320), ("lin", 4096, 2560, 320), ("lin", 4096, 320, 1280),
    ("lin", 1024, 640, 640), ("lin", 1024, 5120, 640), ("lin", 1024, 640, 2560),
    ("lin", 256, 1280, 1280), ("lin", 256, 10240, 1280), ("lin", 256, 1280, 5120),
    ("lin", 77, 320, 1024), ("lin", 77, 1280, 1024), ("lin", 64, 1280, 1280),
    ("conv", (1, 64, 64), 320, 320), ("conv", (1, 64, 64), 320, 960), ("conv", (1, 64, 64), 320, 640),
    ("conv", (1, 32, 32), 640, 640), ("conv", (1, 32, 32), 640, 1920), ("conv", (1, 32, 32), 640, 320),
    ("conv", (1, 16, 16), 1280, 1280), ("conv", (1, 16, 16), 1280, 2560), ("conv", (1, 8, 8), 1280, 1280),
    ("conv", (1, 8, 8), 1280, 2560), ("conv", (1, 64, 64), 640, 640),
]
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
tot_t, tot_f = 0.0, 0.0
only_idx = os.environ.get("ONLY_IDX")
for idx, (kind, m, N, K) in enumerate(shapes):
    if only and only != kind:
        continue
    if only_idx is not None and idx != int(only_idx):
        continue
    if kind == "lin":
        A = torch.randn(m, K, device=dev).to(BF)
        B = torch.randn(N, K, device=dev).to(BF)
        D = torch.empty(m, N, dtype=BF, device=dev)
        R = torch.randn(m, N, device=dev).to(BF)
        bias = torch.randn(N, device=dev)
        f = lambda: ops.gemm(A, B, D, bias=bias, R=R, ws=ws, force_bn=int(os.environ.get('FORCE_BN', 0)), force_split=int(os.environ.get('FORCE_SPLIT', 0)))
        fl = 2.0 * m * N * K
        lab = f"lin  M{m} N{N} K{K}"
    else:
        nb, H, W = m
        x = torch.randn(nb, H, W, K, device=dev).to(BF)
        B = torch.randn(N, 9 * K, device=dev).to(BF)
        D = torch.empty(nb, H, W, N, dtype=BF, device=dev)
        R = torch.randn(nb, H, W, N, device=dev).to(BF)
        bias = torch.randn(N, device=dev)
        f = lambda: ops.conv3x3(x, B, D, bias=bias, R=R, ws=ws, force_bn=int(os.environ.get('FORCE_BN', 0)), force_split=int(os.environ.get('FORCE_SPLIT', 0)))
        fl = 2.0 * nb * H * W * N * 9 * K
        lab = f"conv {H}x{W} C{K} N{N}"
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    t = ts[len(ts) // 2]
    tot_t += t; tot_f += fl
    print(f"{lab:28s} {t:8.1f} us  {fl / t / 1e6:8.1f} TFLOP/s", flush=True)
print(f"sum {tot_t:.1f} us, {tot_f / tot_t / 1e6:.1f} TFLOP/s aggregate")
