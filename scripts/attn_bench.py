"""Micro-benchmark of the attention kernels (fwd, bwd) on the SD-2.1 shapes."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from view_neti_b200 import ops

BF = torch.bfloat16
dev = "cuda"
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 10
which = sys.argv[2] if len(sys.argv) > 2 else "fwd,bwd"
shapes = [(1, 5, 4096, 4096), (1, 10, 1024, 1024), (1, 20, 256, 256), (1, 20, 64, 64), (1, 5, 4096, 77), (1, 10, 1024, 77),
          (2, 5, 4096, 4096)]
only = os.environ.get("ONLY_IDX")
acc = torch.zeros(2 * 2 * 80 * 1280, dtype=torch.float64, device=dev)
ws = torch.empty(64 << 20, dtype=torch.uint8, device=dev)      # scratch for the split leftover items (persistent schedule)
for idx, (nb, h, nq, nk) in enumerate(shapes):
    if only is not None and idx != int(only):
        continue
    C = h * 64
    q = torch.randn(nb, nq, C, device=dev).to(BF)
    k = torch.randn(nb, nk, C, device=dev).to(BF)
    v = torch.randn(nb, nk, C, device=dev).to(BF)
    o = torch.empty_like(q)
    d_o = torch.randn_like(q)
    lse = torch.empty(nb, h, nq, device=dev)
    delta = torch.empty(nb, h, nq, device=dev)
    dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
    fl = 4.0 * nb * h * nq * nk * 64
    for name, f, mult in (("fwd", lambda: ops.attention_fwd(q, k, v, o, lse, h, ws=ws), 1.0),
                          ("bwd", lambda: ops.attention_bwd(q, k, v, o, lse, d_o, delta, dq, dk, dv, h,
                                                            dkv_acc=acc if nk < 256 else None, ws=ws), 2.5)):
        if name not in which:
            continue
        for _ in range(3):
            f()
        torch.cuda.synchronize()
        ts = []
        for _ in range(iters):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); f(); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        ts.sort()
        t = ts[len(ts) // 2]
        print(f"attn {name} nb{nb} h{h} nq{nq} nk{nk}: {t:8.1f} us  {fl * mult / t / 1e6:8.1f} TFLOP/s", flush=True)
