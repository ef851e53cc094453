"""In-kernel timelines (clock64 stamps written by vn_gemm / fused GroupNorm CTAs when vn_set_debug_buffer is armed):
where inside a launch the microseconds go.  L2 is flushed before every launch, so weights come from HBM as they do
inside a train step.  Prints, per shape, the median / max over CTAs of each phase's time since CTA entry, in ns."""
import os
import sys

# the stamps are compiled out of the product library: build the instrumented one first
#   VN_LIB_SUFFIX=_tl VN_CFLAGS=-DVN_TIMELINE python -m view_neti_b200.build
os.environ.setdefault("VN_LIB_SUFFIX", "_tl")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from view_neti_b200 import _abi, ops

BF = torch.bfloat16
dev = "cuda"
ws = ops.Workspace(8192, 10240, dev)
dbg = torch.zeros(16 * 2048, dtype=torch.int64, device=dev)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
lib = _abi.load()

GEMM_SLOTS = {2: "setup done", 3: "pdl_wait passed", 4: "first A load issued", 5: "tile-0 loads issued",
              6: "first stage landed", 7: "tile-0 MMAs issued", 8: "tile-0 accumulator ready", 12: "cluster sync 1",
              13: "cluster sync 2", 14: "tile-0 staged in smem", 15: "tile-0 store issued", 9: "epilogue done",
              10: "CTA exit"}
GN_SLOTS = {2: "pdl_wait passed", 3: "phase-1 loads+sums", 4: "CTA reduce, partials stored",
            6: "all partials arrived", 8: "totals + group constants", 7: "phase-2 stores issued"}


def report(label, slots, event_ms):
    torch.cuda.synchronize()
    d = dbg.view(-1, 16).cpu()
    used = d[:, 1] != 0
    d = d[used]
    n = d.shape[0]
    line = f"{label}: {n} CTAs, launch {event_ms * 1e3:.1f} us"
    if slots is GEMM_SLOTS:
        span = (d[:, 11].max() - d[:, 0].min()).item()
        cyc = (d[:, 10] - d[:, 1]).float()
        ns = (d[:, 11] - d[:, 0]).float().clamp(min=1)
        ghz = float((cyc / ns).median())
        line += f", first entry -> last exit {span / 1e3:.1f} us, entry skew {(d[:, 0].max() - d[:, 0].min()).item() / 1e3:.1f} us, {ghz:.2f} GHz"
    else:
        ghz = 1.9
    print(line)
    for s, name in slots.items():
        v = d[:, s]
        ok = v != 0
        if ok.sum() == 0:
            continue
        dt = (v[ok] - d[ok, 1]).float() / ghz
        print(f"    {name:28s} median {dt.median().item():8.0f} ns   max {dt.max().item():8.0f} ns")


def timed(f):
    flush.zero_()
    dbg.zero_()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); f(); e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1)


lib.vn_set_debug_buffer(dbg.data_ptr())
ops.set_pdl(False)
shapes = [("lin", 4096, 320, 320), ("lin", 1024, 640, 640), ("lin", 256, 1280, 1280), ("lin", 77, 320, 1024),
          ("lin", 64, 1280, 1280), ("lin", 4096, 2560, 320), ("lin", 256, 10240, 1280),
          ("conv", (1, 8, 8), 1280, 1280), ("conv", (1, 16, 16), 1280, 1280), ("conv", (1, 32, 32), 640, 640),
          ("conv", (1, 64, 64), 320, 320), ("conv", (1, 64, 64), 320, 960)]
for kind, m, N, K in shapes:
    if kind == "lin":
        A = torch.randn(m, K, device=dev).to(BF); B = torch.randn(N, K, device=dev).to(BF)
        D = torch.empty(m, N, dtype=BF, device=dev); bias = torch.randn(N, device=dev)
        f = lambda: ops.gemm(A, B, D, bias=bias, ws=ws)
        lab = f"lin M{m} N{N} K{K}"
    else:
        nb, H, W = m
        x = torch.randn(nb, H, W, K, device=dev).to(BF); B = torch.randn(N, 9 * K, device=dev).to(BF)
        D = torch.empty(nb, H, W, N, dtype=BF, device=dev); bias = torch.randn(N, device=dev)
        f = lambda: ops.conv3x3(x, B, D, bias=bias, ws=ws)
        lab = f"conv {H}x{W} C{K} N{N}"
    f(); f()
    ms = timed(f)
    report(lab, GEMM_SLOTS, ms)

for (nb, hw, C) in [(1, 4096, 320), (1, 4096, 960), (1, 1024, 640), (1, 256, 1280), (1, 64, 1280)]:
    x = torch.randn(nb, hw, C, device=dev).to(BF)
    dy = torch.randn(nb, hw, C, device=dev).to(BF)
    y = torch.empty_like(x)
    gamma, beta = torch.ones(C, device=dev), torch.zeros(C, device=dev)
    stats = torch.zeros(nb, 32, 2, dtype=torch.float64, device=dev)
    red = torch.zeros(nb, 32, 2, dtype=torch.float64, device=dev)
    bar = torch.empty(2, ops.groupnorm_partial_floats(nb), device=dev)

    def fwd():
        ops.groupnorm_fwd(x, gamma, beta, 1e-5, True, y, nb, hw, 32, stats, bar[0])

    def bwd():
        ops.groupnorm_bwd_fused(x, dy, stats, red, bar[1], gamma, beta, 1e-5, True, y, nb, hw, 32)

    for name, fn in (("fwd", fwd), ("bwd", bwd)):
        stats.zero_() if name == "fwd" else red.zero_()
        ops.memset(bar, 0xFF)
        ms = timed(fn)
        report(f"groupnorm {name} nb{nb} hw{hw} C{C} (cold)", GN_SLOTS, ms)
        # L2-warm (as inside a step: the producer just wrote x)
        stats.zero_() if name == "fwd" else red.zero_()
        ops.memset(bar, 0xFF); dbg.zero_()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        report(f"groupnorm {name} nb{nb} hw{hw} C{C} (warm)", GN_SLOTS, e0.elapsed_time(e1))
lib.vn_set_debug_buffer(None)
