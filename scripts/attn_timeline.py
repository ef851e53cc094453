"""In-kernel timeline of the attention forward (64x64 self-attention, CTA 0, iterations 8..15): clock64 stamps written by the
instrumented build.  Build it first:   VN_LIB_SUFFIX=_tl VN_CFLAGS=-DVN_TIMELINE python -m view_neti_b200.build
Prints, per key tile j, the cycle at which each role passed its events, relative to "S(8) issued"."""
import os
import sys

os.environ.setdefault("VN_LIB_SUFFIX", "_tl")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from view_neti_b200 import _abi, ops

dev, BF = "cuda", torch.bfloat16
lib = _abi.load()
dbg = torch.zeros(16 * 2048, dtype=torch.int64, device=dev)
nb, h, nq, nk = 1, 5, 4096, 4096
C = h * 64
q, k, v = (torch.randn(nb, n, C, device=dev).to(BF) for n in (nq, nk, nk))
o = torch.empty_like(q)
lse = torch.empty(nb, h, nq, device=dev)
ws = torch.empty(64 << 20, dtype=torch.uint8, device=dev)
for _ in range(3):
    ops.attention_fwd(q, k, v, o, lse, h, ws=ws)
torch.cuda.synchronize()
lib.vn_set_debug_buffer(dbg.data_ptr())
ops.attention_fwd(q, k, v, o, lse, h, ws=ws)
torch.cuda.synchronize()
lib.vn_set_debug_buffer(0)
d = dbg[:8 * 16].view(8, 16).cpu()
t0 = int(d[0, 0])
names = ["S(j) issued", "PV(j) issued", "s_full(j) seen", "arrive p_full(j)", "after acc PV(j-1)", "kv stage free", "TMA(j) issued",
         "kv_full(j) seen"]
print("cycles relative to S(8) issued; j = key tile")
print("j   " + "  ".join(f"{n:>18s}" for n in names))
for j in range(8):
    print(f"{j + 8:<3d} " + "  ".join(f"{(int(d[j, e]) - t0) if int(d[j, e]) else 0:18d}" for e in range(8)))

# ---- backward, dK/dV body of CTA 0 ----
d_o = torch.randn_like(q)
delta = torch.empty(nb, h, nq, device=dev)
dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
for _ in range(2):
    ops.attention_bwd(q, k, v, o, lse, d_o, delta, dq, dk, dv, h, ws=ws)
torch.cuda.synchronize()
dbg.zero_()
lib.vn_set_debug_buffer(dbg.data_ptr())
ops.attention_bwd(q, k, v, o, lse, d_o, delta, dq, dk, dv, h, ws=ws)
torch.cuda.synchronize()
lib.vn_set_debug_buffer(0)
d = dbg[1024:1024 + 8 * 16].view(8, 16).cpu()
t0 = int(d[0, 0])
names = ["SdP(i,h0) issued", "SdP(i,h1) issued", "dVdK(i,h0) issued", "dVdK(i,h1) issued", "s_full h0 seen", "s_full h1 seen",
         "arrive p_full h0", "arrive p_full h1"]
print("backward dK/dV body: cycles relative to SdP(8, half 0) issued; i = query tile")
print("i   " + "  ".join(f"{n:>17s}" for n in names))
for i in range(8):
    print(f"{i + 8:<3d} " + "  ".join(f"{(int(d[i, e]) - t0) if int(d[i, e]) else 0:17d}" for e in range(8)))
