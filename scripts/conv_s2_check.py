"""Stride-2 3x3 convolutions straight from the NHWC tensor map (TMA element strides, vn_gemm modes 2 / 3) against the
explicit im2col + GEMM they replace: same k order (tap*C + c), so the results should agree bit for bit; then the time of
both forms at the UNet's and the VAE encoder's downsample shapes (L2 flushed between launches)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from view_neti_b200 import ops

BF = torch.bfloat16
dev = "cuda"
ws = ops.Workspace(8192, 8192, dev)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
g = torch.Generator(device=dev).manual_seed(0)


def timed(f, n=10):
    ts = []
    for _ in range(n):
        flush.zero_()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return sorted(ts)[len(ts) // 2]


ok = True
for (nb, H, W, C, pad) in [(1, 64, 64, 320, 1), (1, 32, 32, 640, 1), (1, 16, 16, 1280, 1), (2, 48, 64, 320, 1), (1, 9, 13, 64, 1),
                           (1, 512, 512, 128, 0), (1, 256, 256, 256, 0), (1, 128, 128, 512, 0), (2, 64, 96, 128, 0), (1, 6, 10, 64, 0)]:
    x = torch.randn(nb, H, W, C, device=dev, generator=g).to(BF)
    wk = (torch.randn(C, 9 * C, device=dev, generator=g) / (9 * C) ** 0.5).to(BF)
    bias = torch.randn(C, device=dev, generator=g)
    Ho, Wo = ((H - 1) // 2 + 1, (W - 1) // 2 + 1) if pad else ((H - 2) // 2 + 1, (W - 2) // 2 + 1)
    col = torch.empty(nb * Ho * Wo, 9 * C, dtype=BF, device=dev)
    d_ref = torch.empty(nb, Ho, Wo, C, dtype=BF, device=dev)
    d_new = torch.full((nb, Ho, Wo, C), 7.0, dtype=BF, device=dev)

    def old():
        (ops.im2col_s2 if pad else ops.im2col_s2_pad0)(x, col)
        ops.gemm(col, wk, d_ref.view(-1, C), bias=bias, ws=ws)

    def new():
        ops.conv3x3(x, wk, d_new, bias=bias, ws=ws, stride=2, pad=pad)

    old(); new()
    torch.cuda.synchronize()
    same = torch.equal(d_ref, d_new)
    err = float((d_ref.float() - d_new.float()).abs().max())
    ok &= err < 2e-2
    print(f"nb{nb} {H}x{W} C{C} pad{pad}: bit-equal {same}, max |diff| {err:.3g}, im2col+gemm {timed(old):7.1f} us, tma-strided {timed(new):7.1f} us")
print("OK" if ok else "MISMATCH")
sys.exit(0 if ok else 1)
