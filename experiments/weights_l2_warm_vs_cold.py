"""Do the deep-level weight-streaming GEMMs run faster when their weights are already in L2?  (decides whether an L2
prefetch stream that runs one layer ahead of the chain is worth building)

Each shape is timed inside a CUDA graph of 8 launches (PDL on, as in the step):
  warm : the same weight buffer every launch (29-59 MB: stays in the 126 MB L2)
  cold : 8 different weight buffers, > 230 MB in rotation (every launch streams from HBM)
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from view_neti_b200 import ops

dev = "cuda"
BF = torch.bfloat16
ws = ops.Workspace(4096, 10240, dev)


def bench(kind, nb, hw, C, N, reps=8):
    H = W = hw
    if kind == "conv":
        x = torch.randn(nb, H, W, C, device=dev, dtype=BF)
        K = 9 * C
    else:
        x = torch.randn(nb * H * W, C, device=dev, dtype=BF)
        K = C
    wts = [torch.randn(N, K, device=dev, dtype=BF) * 0.02 for _ in range(reps)]
    D = torch.empty((nb, H, W, N) if kind == "conv" else (nb * H * W, N), device=dev, dtype=BF)
    out = {}
    for mode in ("warm", "cold"):
        def run():
            for i in range(reps):
                w = wts[0] if mode == "warm" else wts[i]
                if kind == "conv":
                    ops.conv3x3(x, w, D, ws=ws)
                else:
                    ops.gemm(x, w, D, ws=ws)
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            run(); torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=s):
                run()
        for _ in range(3):
            g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            g.replay()
        e1.record(); torch.cuda.synchronize()
        out[mode] = e0.elapsed_time(e1) / 20 / reps * 1e3
    mb = N * K * 2 / 1e6
    print(f"{kind:5s} M={nb*H*W:5d} N={N:5d} K={K:6d}  weights {mb:5.1f} MB   warm {out['warm']:6.2f} us   cold {out['cold']:6.2f} us"
          f"   cold HBM rate {mb / out['cold'] / 1e3:5.2f} TB/s", flush=True)


for a in [("conv", 1, 8, 1280, 1280), ("conv", 1, 16, 1280, 1280), ("conv", 1, 8, 2560, 1280), ("conv", 1, 16, 2560, 1280),
          ("lin", 1, 16, 1280, 1280), ("lin", 1, 16, 1280, 10240), ("lin", 1, 16, 5120, 1280), ("lin", 1, 8, 1280, 1280),
          ("conv", 1, 32, 640, 640), ("lin", 1, 32, 640, 640)]:
    bench(*a)
