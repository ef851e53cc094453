"""Times every distinct GEMM / conv shape of one train step under each (BN, split) configuration (10 back-to-back
launches inside a CUDA graph) and writes view_neti_b200/gemm_tuning.json, which ops.gemm / ops.conv3x3 consult."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from tests.unet_parity import make_inputs
from view_neti_b200 import ops
from view_neti_b200._abi import VNError
from view_neti_b200.sd21 import SD21, init_state_dict
from view_neti_b200.unet import UNet2DConditionModel

BF = torch.bfloat16
cases = [(1, 64, 64), (2, 64, 64), (1, 32, 32), (2, 96, 72)] if len(sys.argv) < 2 else [tuple(map(int, a.split("x"))) for a in sys.argv[1:]]
model = UNet2DConditionModel(init_state_dict(SD21, 0), SD21, "cuda")
seen = {}
orig_gemm, orig_conv = ops.gemm, ops.conv3x3


def rec_gemm(A, B, D, **kw):
    M = A.numel() // A.shape[-1]
    key = ops.tuning_key(0, M, B.shape[0], A.shape[-1], 0, 0)
    seen.setdefault(key, ("gemm", M, B.shape[0], A.shape[-1], None, kw.get("bias") is not None, kw.get("R") is not None,
                          D.dtype == torch.float32))
    return orig_gemm(A, B, D, **kw)


def rec_conv(x, Wk, D, **kw):
    nb, H, W, C = x.shape
    if kw.get("stride", 1) != 1:                    # strided tensor-map convs (3 per forward) keep the cost model's choice
        return orig_conv(x, Wk, D, **kw)
    key = ops.tuning_key(1, nb * H * W, Wk.shape[0], 9 * C, H, W)
    seen.setdefault(key, ("conv", nb * H * W, Wk.shape[0], C, (nb, H, W), kw.get("bias") is not None, kw.get("R") is not None, False))
    return orig_conv(x, Wk, D, **kw)


ops.TUNING.clear()
ops.gemm, ops.conv3x3 = rec_gemm, rec_conv
for nb, h, w in cases:
    plan = model.engine.plan(nb, h, w)
    lat, t, tgt, ctx = make_inputs(SD21, nb, h, w, seed=1)
    plan.latents.copy_(lat); plan.timesteps.copy_(t); plan.target.copy_(tgt)
    plan.train_step()
    torch.cuda.synchronize()
    del plan
    model.engine._plans.clear()
ops.gemm, ops.conv3x3 = orig_gemm, orig_conv
print(f"{len(seen)} distinct shapes", flush=True)

REP = 10
out = {}
tot_def = tot_best = 0.0
for key, (kind, M, N, K, geom, has_bias, has_r, f32) in seen.items():
    if kind == "gemm":
        A = torch.randn(M, K, device="cuda").to(BF)
        Bm = torch.randn(N, K, device="cuda").to(BF)
        D = torch.empty(M, N, dtype=torch.float32 if f32 else BF, device="cuda")
        R = torch.randn(M, N, device="cuda").to(BF) if has_r else None
        call = lambda bn, sp: orig_gemm(A, Bm, D, bias=bias, R=R, force_bn=bn, force_split=sp)
    else:
        nb, H, W = geom
        A = torch.randn(nb, H, W, K, device="cuda").to(BF)
        Bm = torch.randn(N, 9 * K, device="cuda").to(BF)
        D = torch.empty(nb, H, W, N, dtype=BF, device="cuda")
        R = torch.randn(nb, H, W, N, device="cuda").to(BF) if has_r else None
        call = lambda bn, sp: orig_conv(A, Bm, D, bias=bias, R=R, force_bn=bn, force_split=sp)
    bias = torch.randn(N, device="cuda") if has_bias else None
    res = {}
    # force_split code: low 4 bits split-K cluster size, +256 = CTA pairs (cta_group::2) on, +512 = pairs off
    cands = [(b, s + 512) for b in (64, 128, 192, 256) for s in (1, 2, 4, 8)] + [(160, 2 + 512), (160, 4 + 512)]
    if ((M + 127) // 128) % 2 == 0:              # CTA pairs need an even number of m-tiles (vn_gemm ignores the request otherwise)
        cands += [(b, 1 + 256) for b in (128, 192, 256)]
    for bn, sp in [(0, 0)] + cands:
        if os.environ.get("VN_TUNE_VERBOSE"):
            print(f"  {key} bn{bn} sp{sp}", flush=True)
        try:
            call(bn, sp)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                for _ in range(REP):
                    call(bn, sp)
            ts = []
            for _ in range(4):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); g.replay(); e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1) * 1e3 / REP)
            res[(bn, sp)] = min(ts)
        except (VNError, RuntimeError) as e:
            if "CUDA" in str(e) and "vn_gemm" not in str(e):
                raise
    d = res.pop((0, 0))
    (bn, sp), best = min(res.items(), key=lambda kv: kv[1])
    tot_def += d; tot_best += min(best, d)
    if best < 0.97 * d:
        out[key] = [bn, sp, round(best, 2), round(d, 2)]
    print(f"{key:34s} default {d:7.2f} us  best {best:7.2f} us  bn{bn} s{sp & 15}{' pair' if sp & 256 else ''}", flush=True)
print(f"sum over distinct shapes: default {tot_def:.1f} us, tuned {tot_best:.1f} us")
path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "gemm_tuning.json")
os.makedirs(os.path.dirname(path), exist_ok=True)
with open(path, "w") as f:
    json.dump(out, f, indent=0, sort_keys=True)
with open(path.replace(".json", "_seen.json"), "w") as f:      # every key timed in this run (winner or not): lets a merge drop stale entries
    json.dump(sorted(seen), f)
print("wrote", path, len(out), "entries")
