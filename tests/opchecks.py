"""Per-op parity checks of the CUDA library against plain PyTorch fp32 references on the same (bf16-rounded)
inputs.  Used by tests/test_ops_gpu.py (asserting) and scripts/gpu_diag.py (printing every error figure).
Each check returns a list of (label, rel_l2_error, tolerance)."""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

from view_neti_b200 import ops

DEV = "cuda"
BF = torch.bfloat16
# the fp32 references must be real fp32 (cuDNN convolutions default to TF32)
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def rel(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.detach().float(), b.detach().float()
    if not torch.isfinite(a).all():
        return float("inf")
    return float((a - b).norm() / (b.norm() + 1e-12))


def rnd(*shape, scale=1.0, seed=None):
    g = torch.Generator(device="cpu")
    g.manual_seed(seed if seed is not None else (hash(shape) & 0xFFFF))
    return (torch.randn(*shape, generator=g) * scale).to(DEV)


_ws = None


def ws():
    global _ws
    if _ws is None:
        _ws = ops.Workspace(8192, 10240, DEV, dkv_elems=2 * 2 * 80 * 1280)
    return _ws


# ------------------------------------------------------------------------------------------------
def check_gemm(M, N, K, bias=False, rowbias=False, resid=False, out_fp32=False, force_bn=0, force_split=0,
               strided=False):
    A = rnd(M, K, seed=1).to(BF)
    B = (rnd(N, K, seed=2) / math.sqrt(K)).to(BF)
    if strided:                                     # A and D as channel-slices of wider buffers
        Abuf = torch.zeros(M, K + 64, dtype=BF, device=DEV)
        Abuf[:, 64:] = A
        A = Abuf[:, 64:]
        Dbuf = torch.zeros(M, N + 32, dtype=torch.float32 if out_fp32 else BF, device=DEV)
        D = Dbuf[:, 32:]
    else:
        D = torch.empty(M, N, dtype=torch.float32 if out_fp32 else BF, device=DEV)
    bv = rnd(N, seed=3) if bias else None
    rpb = 64 if M % 64 == 0 else M
    rb = rnd(M // rpb, N, seed=4) if rowbias else None
    R = rnd(M, N, seed=5).to(BF) if resid else None
    ops.gemm(A, B, D, bias=bv, rowbias=rb, rows_per_batch=rpb if rowbias else 0, R=R, ws=ws(),
             force_bn=force_bn, force_split=force_split)
    ref = A.float() @ B.float().t()
    if bias:
        ref = ref + bv
    if rowbias:
        ref = ref + rb.repeat_interleave(rpb, 0)
    if resid:
        ref = ref + R.float()
    torch.cuda.synchronize()
    clean = float(ws().buf.float().abs().max())
    lab = f"gemm M{M} N{N} K{K} b{int(bias)} rb{int(rowbias)} r{int(resid)} f32{int(out_fp32)} bn{force_bn} sp{force_split} st{int(strided)}"
    return [(lab, rel(D, ref), 6e-3), (lab + " ws-clean", clean, 0.0)]


def conv_weight_to_k(w):          # [N, C, 3, 3] -> [N, 9*C] with k = tap*C + c
    return w.permute(0, 2, 3, 1).reshape(w.shape[0], -1).contiguous()


def conv_weight_to_k_dgrad(w):    # dgrad weights: [C, 9*N] with taps flipped
    return w.flip(2, 3).permute(1, 2, 3, 0).reshape(w.shape[1], -1).contiguous()


def check_conv(nb, H, W, Cc, N, bias=True, rowbias=False, resid=False, force_bn=0, force_split=0, dgrad=False):
    x = rnd(nb, H, W, Cc, seed=6).to(BF)
    w = (rnd(N, Cc, 3, 3, seed=7) / math.sqrt(9 * Cc)).to(BF)
    bv = rnd(N, seed=8) if bias else None
    rb = rnd(nb, N, seed=9) if rowbias else None
    R = rnd(nb, H, W, N, seed=10).to(BF) if resid else None
    D = torch.empty(nb, H, W, N, dtype=BF, device=DEV)
    ops.conv3x3(x, conv_weight_to_k(w), D, bias=bv, rowbias=rb, R=R, ws=ws(), force_bn=force_bn,
                force_split=force_split)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.float(), bv, padding=1).permute(0, 2, 3, 1)
    if rowbias:
        ref = ref + rb[:, None, None, :]
    if resid:
        ref = ref + R.float()
    out = [(f"conv nb{nb} {H}x{W} C{Cc} N{N} rb{int(rowbias)} r{int(resid)} bn{force_bn} sp{force_split}", rel(D, ref), 6e-3)]
    if dgrad:
        dy = rnd(nb, H, W, N, seed=11).to(BF)
        dx = torch.empty(nb, H, W, Cc, dtype=BF, device=DEV)
        ops.conv3x3(dy, conv_weight_to_k_dgrad(w), dx, ws=ws())
        xr = x.float().permute(0, 3, 1, 2).requires_grad_(True)
        F.conv2d(xr, w.float(), None, padding=1).backward(dy.float().permute(0, 3, 1, 2))
        out.append((f"conv-dgrad nb{nb} {H}x{W} C{Cc} N{N}", rel(dx, xr.grad.permute(0, 2, 3, 1)), 6e-3))
    return out


def check_groupnorm(nb, hw, Cc, silu, eps=1e-5, groups=32):
    x = (rnd(nb, hw, Cc, seed=12) * 1.5 + 0.3).to(BF)
    gamma, beta = 1 + 0.1 * rnd(Cc, seed=13), 0.1 * rnd(Cc, seed=14)
    stats = torch.zeros(nb, groups, 2, device=DEV, dtype=torch.float64)
    y = torch.empty_like(x)
    ops.groupnorm_stats(x, nb, hw, groups, stats)
    ops.groupnorm_apply(x, stats, gamma, beta, eps, silu, y, nb, hw, groups)
    xr = x.float().permute(0, 2, 1).requires_grad_(True)     # [nb, C, hw]
    yr = F.group_norm(xr, groups, gamma, beta, eps)
    if silu:
        yr = F.silu(yr)
    dy = rnd(nb, hw, Cc, seed=15).to(BF)
    yr.backward(dy.float().permute(0, 2, 1))
    red = torch.zeros(nb, groups, 2, device=DEV, dtype=torch.float64)
    add1 = rnd(nb, hw, Cc, seed=16).to(BF)
    dx = torch.empty_like(x)
    ops.groupnorm_bwd(x, dy, stats, red, gamma, beta, eps, silu, dx, nb, hw, groups, add1=add1)
    lab = f"groupnorm nb{nb} hw{hw} C{Cc} silu{int(silu)}"
    # one-launch forms (what the engine calls): statistics + grid barrier + apply; must reproduce the two-kernel results
    stats2 = torch.zeros_like(stats)
    red2 = torch.zeros_like(red)
    part = torch.empty(2, ops.groupnorm_partial_floats(nb), device=DEV)
    ops.memset(part, 0xFF)
    y2 = torch.empty_like(x)
    dx2 = torch.empty_like(x)
    ops.groupnorm_fwd(x, gamma, beta, eps, silu, y2, nb, hw, groups, stats2, part[0])
    ops.groupnorm_bwd_fused(x, dy, stats2, red2, part[1], gamma, beta, eps, silu, dx2, nb, hw, groups, add1=add1)
    return [(lab + " fwd", rel(y, yr.permute(0, 2, 1)), 6e-3),
            (lab + " bwd", rel(dx, xr.grad.permute(0, 2, 1) + add1.float()), 8e-3),
            (lab + " fused fwd", rel(y2, yr.permute(0, 2, 1)), 6e-3),
            (lab + " fused bwd", rel(dx2, xr.grad.permute(0, 2, 1) + add1.float()), 8e-3),
            # (per-thread fp32 partial sums cover different pixel sets in the two geometries: equal up to fp32 rounding)
            (lab + " fused ~ two-kernel (y)", rel(y2, y), 2e-4),
            (lab + " fused ~ two-kernel (dx)", rel(dx2, dx), 2e-4),
            (lab + " fused stats", float((stats2 - stats).norm() / stats.norm()), 1e-5),
            (lab + " fused red", float((red2 - red).abs().max() / (red.abs().max() + 1e-30)), 1e-3)]


def check_layernorm(rows, Cc):
    x = (rnd(rows, Cc, seed=17) * 2 + 0.5).to(BF)
    gamma, beta = 1 + 0.1 * rnd(Cc, seed=18), 0.1 * rnd(Cc, seed=19)
    y = torch.empty_like(x)
    stats = torch.empty(rows, 2, device=DEV)
    ops.layernorm_fwd(x, gamma, beta, 1e-5, y, stats, rows)
    xr = x.float().requires_grad_(True)
    yr = F.layer_norm(xr, (Cc,), gamma, beta, 1e-5)
    dy = rnd(rows, Cc, seed=20).to(BF)
    yr.backward(dy.float())
    add = rnd(rows, Cc, seed=21).to(BF)
    dx = torch.empty_like(x)
    ops.layernorm_bwd(x, dy, gamma, stats, dx, rows, add=add)
    return [(f"layernorm {rows}x{Cc} fwd", rel(y, yr), 6e-3),
            (f"layernorm {rows}x{Cc} bwd", rel(dx, xr.grad + add.float()), 8e-3)]


def check_layernorm_f32(rows, Cc):
    """fp32-stream forms (CLIP text encoder residual stream): fp32 x / add / dx, bf16 y and bf16 copy of dx."""
    x = rnd(rows, Cc, seed=27) * 2 + 0.5
    gamma, beta = 1 + 0.1 * rnd(Cc, seed=28), 0.1 * rnd(Cc, seed=29)
    y = torch.empty(rows, Cc, dtype=BF, device=DEV)
    stats = torch.empty(rows, 2, device=DEV)
    ops.layernorm_fwd_f32(x, gamma, beta, 1e-5, y, stats, rows)
    xr = x.clone().requires_grad_(True)
    yr = F.layer_norm(xr, (Cc,), gamma, beta, 1e-5)
    dy = rnd(rows, Cc, seed=30).to(BF)
    yr.backward(dy.float())
    add = rnd(rows, Cc, seed=31)
    dx, dxb = torch.empty_like(x), torch.empty(rows, Cc, dtype=BF, device=DEV)
    ops.layernorm_bwd_f32(x, dy, gamma, stats, dx, rows, add=add, dx_bf16=dxb)
    dx2 = torch.empty_like(x)
    ops.layernorm_bwd_f32(x, dy, gamma, stats, dx2, rows, add=None, dx_bf16=None)
    return [(f"layernorm f32 {rows}x{Cc} fwd", rel(y, yr), 4e-3),
            (f"layernorm f32 {rows}x{Cc} bwd", rel(dx, xr.grad + add), 2e-5),
            (f"layernorm f32 {rows}x{Cc} bwd bf16 copy", rel(dxb, xr.grad + add), 4e-3),
            (f"layernorm f32 {rows}x{Cc} bwd no add", rel(dx2, xr.grad), 2e-5)]


def check_gemm_f32_residual(M, N, K, force_bn=0, force_split=0):
    """fp32 output with an fp32 residual (direct epilogue; split-K finish)."""
    A = rnd(M, K, seed=1).to(BF)
    B = (rnd(N, K, seed=2) / math.sqrt(K)).to(BF)
    bv, R = rnd(N, seed=3), rnd(M, N, seed=5) * 3
    D = torch.empty(M, N, device=DEV)
    ops.gemm(A, B, D, bias=bv, R=R, ws=ws(), force_bn=force_bn, force_split=force_split)
    ref = A.float() @ B.float().t() + bv + R
    return [(f"gemm f32 residual M{M} N{N} K{K} bn{force_bn} sp{force_split}", rel(D, ref), 2e-5)]


def check_geglu(rows, Fd):
    h = rnd(rows, 2 * Fd, seed=22).to(BF)
    y = torch.empty(rows, Fd, dtype=BF, device=DEV)
    ops.geglu_fwd(h, y, rows)
    hr = h.float().requires_grad_(True)
    a, g = hr.chunk(2, -1)
    yr = a * F.gelu(g)
    dy = rnd(rows, Fd, seed=23).to(BF)
    yr.backward(dy.float())
    dh = torch.empty_like(h)
    ops.geglu_bwd(h, dy, dh, rows)
    return [(f"geglu {rows}x{Fd} fwd", rel(y, yr), 6e-3), (f"geglu {rows}x{Fd} bwd", rel(dh, hr.grad), 8e-3)]


def attn_ref(q, k, v, heads, scale, causal=False):
    nb, nq, Cc = q.shape
    nk = k.shape[1]
    qh = q.float().reshape(nb, nq, heads, 64).transpose(1, 2)
    kh = k.float().reshape(nb, nk, heads, 64).transpose(1, 2)
    vh = v.float().reshape(nb, nk, heads, 64).transpose(1, 2)
    s = (qh @ kh.transpose(-1, -2)) * scale
    if causal:
        s = s + torch.full((nq, nk), float("-inf"), device=s.device).triu(1)
    p = s.softmax(-1)
    o = (p @ vh).transpose(1, 2).reshape(nb, nq, Cc)
    return o, torch.logsumexp(s, -1)


def check_attention(nb, heads, nq, nk, bwd=True, use_acc=False, strided=False, causal=False):
    Cc = heads * 64
    if strided:   # q/k/v as slices of one fused [nb, n, 3C] buffer (self-attention layout)
        assert nq == nk
        qkv = (rnd(nb, nq, 3 * Cc, seed=24) * 1.2).to(BF)
        q, k, v = qkv[..., :Cc], qkv[..., Cc:2 * Cc], qkv[..., 2 * Cc:]
    else:
        q = (rnd(nb, nq, Cc, seed=25) * 1.2).to(BF)
        k = (rnd(nb, nk, Cc, seed=26) * 1.2).to(BF)
        v = rnd(nb, nk, Cc, seed=27).to(BF)
    o = torch.empty(nb, nq, Cc, dtype=BF, device=DEV)
    lse = torch.empty(nb, heads, nq, device=DEV)
    ops.attention_fwd(q, k, v, o, lse, heads, causal=causal)
    qr, kr, vr = (t.float().detach().clone().requires_grad_(True) for t in (q, k, v))
    oref, lseref = attn_ref(qr, kr, vr, heads, 0.125, causal=causal)
    lab = f"attn nb{nb} h{heads} nq{nq} nk{nk} acc{int(use_acc)} st{int(strided)} causal{int(causal)}"
    out = [(lab + " o", rel(o, oref), 8e-3), (lab + " lse", rel(lse, lseref), 1e-4)]
    # the balanced schedule (leftover items split along the keys, merged by the fix-up launch) against the plain one
    need = ops.attention_fwd_workspace_bytes(nb, heads, nq, nk)
    if need > 0:
        wsb = torch.full((need,), 0xFF, dtype=torch.uint8, device=DEV)          # NaN bit patterns: every slot must be written
        o2, lse2 = torch.full_like(o, float("nan")), torch.full_like(lse, float("nan"))
        ops.attention_fwd(q, k, v, o2, lse2, heads, causal=causal, ws=wsb)
        out += [(lab + " o (balanced vs reference)", rel(o2, oref), 8e-3), (lab + " lse (balanced)", rel(lse2, lseref), 1e-4),
                (lab + " o (balanced vs plain)", rel(o2, o), 4e-3)]
    if bwd:
        d_o = rnd(nb, nq, Cc, seed=28).to(BF)
        oref.backward(d_o.float())
        delta = torch.empty(nb, heads, nq, device=DEV)
        dq = torch.empty(nb, nq, Cc, dtype=BF, device=DEV)
        dk = torch.empty(nb, nk, Cc, dtype=BF, device=DEV)
        dv = torch.empty(nb, nk, Cc, dtype=BF, device=DEV)
        acc = ws().dkv if use_acc else None
        ops.attention_bwd(q, k, v, o, lse, d_o, delta, dq, dk, dv, heads, dkv_acc=acc, causal=causal)
        out += [(lab + " dq", rel(dq, qr.grad), 1.2e-2), (lab + " dk", rel(dk, kr.grad), 1.2e-2),
                (lab + " dv", rel(dv, vr.grad), 1.2e-2)]
        # the persistent schedule (whole items back to back per CTA, leftover items split, parts added by the fix-up launch)
        needb = ops.attention_bwd_workspace_bytes(nb, heads, nq, nk, True)
        if needb > 0 and not use_acc:
            wsb = torch.full((needb,), 0xFF, dtype=torch.uint8, device=DEV)
            dq2, dk2, dv2 = (torch.full_like(t, float("nan")) for t in (dq, dk, dv))
            ops.attention_bwd(q, k, v, o, lse, d_o, delta, dq2, dk2, dv2, heads, causal=causal, ws=wsb)
            out += [(lab + " dq (persistent)", rel(dq2, qr.grad), 1.2e-2), (lab + " dk (persistent)", rel(dk2, kr.grad), 1.2e-2),
                    (lab + " dv (persistent)", rel(dv2, vr.grad), 1.2e-2), (lab + " dq persistent vs plain", rel(dq2, dq), 4e-3)]
            needp = ops.attention_bwd_workspace_bytes(nb, heads, nq, nk, False)          # dq pruned: only dK/dV items
            if needp > 0:
                wsp = torch.full((needp,), 0xFF, dtype=torch.uint8, device=DEV)
                dk3, dv3 = torch.full_like(dk, float("nan")), torch.full_like(dv, float("nan"))
                ops.attention_bwd(q, k, v, o, lse, d_o, delta, None, dk3, dv3, heads, causal=causal, ws=wsp)
                out += [(lab + " dk (persistent, dq pruned)", rel(dk3, kr.grad), 1.2e-2),
                        (lab + " dv (persistent, dq pruned)", rel(dv3, vr.grad), 1.2e-2)]
        if use_acc:
            torch.cuda.synchronize()
            out.append((lab + " acc-clean", float(acc.abs().max()), 0.0))
    return out


def check_resample(nb, H, W, Cc):
    x = rnd(nb, H, W, Cc, seed=29).to(BF)
    y = torch.empty(nb, 2 * H, 2 * W, Cc, dtype=BF, device=DEV)
    ops.upsample2x_fwd(x, y)
    yr = F.interpolate(x.float().permute(0, 3, 1, 2), scale_factor=2.0, mode="nearest").permute(0, 2, 3, 1)
    dy = rnd(nb, 2 * H, 2 * W, Cc, seed=30).to(BF)
    dx = torch.empty_like(x)
    ops.upsample2x_bwd(dy, dx)
    dxr = F.avg_pool2d(dy.float().permute(0, 3, 1, 2), 2).permute(0, 2, 3, 1) * 4
    out = [(f"upsample {H}x{W}x{Cc} fwd", rel(y, yr), 1e-6), (f"upsample {H}x{W}x{Cc} bwd", rel(dx, dxr), 6e-3)]
    # stride-2 conv = im2col + gemm ; dgrad = gemm + col2im
    N = Cc
    w = (rnd(N, Cc, 3, 3, seed=31) / math.sqrt(9 * Cc)).to(BF)
    Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    col = torch.empty(nb * Ho * Wo, 9 * Cc, dtype=BF, device=DEV)
    ops.im2col_s2(x, col)
    D = torch.empty(nb * Ho * Wo, N, dtype=BF, device=DEV)
    ops.gemm(col, conv_weight_to_k(w), D, ws=ws())
    xr = x.float().permute(0, 3, 1, 2).requires_grad_(True)
    ref = F.conv2d(xr, w.float(), None, stride=2, padding=1)
    out.append((f"downsample {H}x{W}x{Cc} fwd", rel(D.reshape(nb, Ho, Wo, N), ref.permute(0, 2, 3, 1)), 6e-3))
    # the same convolution with its taps read straight from the NHWC tensor map (TMA element strides 2, vn_gemm mode 2)
    bv = rnd(N, seed=34)
    D2 = torch.full((nb, Ho, Wo, N), 7.0, dtype=BF, device=DEV)
    ops.conv3x3(x, conv_weight_to_k(w), D2, bias=bv, ws=ws(), stride=2, pad=1)
    out.append((f"downsample {H}x{W}x{Cc} fwd (strided tensor map)",
                rel(D2, ref.detach().permute(0, 2, 3, 1) + bv), 6e-3))
    # and the VAE encoder's form: F.pad(x, (0, 1, 0, 1)) then stride 2 without padding (vn_gemm mode 3)
    if H >= 2 and W >= 2:
        Hp, Wp = (H - 2) // 2 + 1, (W - 2) // 2 + 1
        D3 = torch.full((nb, Hp, Wp, N), 7.0, dtype=BF, device=DEV)
        ops.conv3x3(x, conv_weight_to_k(w), D3, bias=bv, ws=ws(), stride=2, pad=0)
        ref3 = F.conv2d(F.pad(x.float().permute(0, 3, 1, 2), (0, 1, 0, 1)), w.float(), bv, stride=2)
        out.append((f"downsample {H}x{W}x{Cc} fwd (strided tensor map, pad (0,1,0,1))", rel(D3, ref3.permute(0, 2, 3, 1)), 6e-3))
    dyo = rnd(nb, Ho, Wo, N, seed=32).to(BF)
    ref.backward(dyo.float().permute(0, 3, 1, 2))
    Wt = w.permute(2, 3, 1, 0).reshape(9 * Cc, N).contiguous()        # [9C, N]: dcol = dy @ Wk  (B = Wk^T)
    dcol = torch.empty(nb * Ho * Wo, 9 * Cc, dtype=BF, device=DEV)
    ops.gemm(dyo.reshape(-1, N), Wt, dcol, ws=ws())
    add = rnd(nb, H, W, Cc, seed=33).to(BF)
    dxx = torch.empty_like(x)
    ops.col2im_s2(dcol, dxx, add=add)
    out.append((f"downsample {H}x{W}x{Cc} dgrad", rel(dxx, xr.grad.permute(0, 2, 3, 1) + add.float()), 8e-3))
    return out


def check_edge_convs(nb, H, W, Cw):
    x = rnd(nb, 4, H, W, seed=34)
    w = rnd(Cw, 4, 3, 3, seed=35) / 6.0
    b = rnd(Cw, seed=36)
    y = torch.empty(nb, H, W, Cw, dtype=BF, device=DEV)
    ops.conv_in_fwd(x, w, b, y)
    out = [(f"conv_in {H}x{W}x{Cw}", rel(y, F.conv2d(x, w, b, padding=1).permute(0, 2, 3, 1)), 4e-3)]
    xo = rnd(nb, H, W, Cw, seed=37).to(BF)
    wo = rnd(4, Cw, 3, 3, seed=38) / math.sqrt(9 * Cw)
    bo = rnd(4, seed=39)
    yo = torch.empty(nb, 4, H, W, device=DEV)
    ops.conv_out_fwd(xo, wo, bo, yo)
    xr = xo.float().permute(0, 3, 1, 2).requires_grad_(True)
    ref = F.conv2d(xr, wo, bo, padding=1)
    out.append((f"conv_out {H}x{W}x{Cw}", rel(yo, ref), 1e-4))
    dy = rnd(nb, 4, H, W, seed=40)
    ref.backward(dy)
    dx = torch.empty(nb, H, W, Cw, dtype=BF, device=DEV)
    ops.conv_out_bwd(dy, wo, dx)
    out.append((f"conv_out_bwd {H}x{W}x{Cw}", rel(dx, xr.grad.permute(0, 2, 3, 1)), 4e-3))
    return out


def check_misc():
    out = []
    t = torch.tensor([0, 1, 500, 999], dtype=torch.int64, device=DEV)
    emb = torch.empty(4, 320, device=DEV)
    ops.timestep_sinusoid(t, emb)
    half = 160
    ex = -math.log(10000.0) * torch.arange(half, dtype=torch.float32, device=DEV) / half
    e = t[:, None].float() * torch.exp(ex)[None]
    out.append(("timestep_sinusoid", rel(emb, torch.cat([e.cos(), e.sin()], -1)), 2e-5))
    x = rnd(2, 1280, seed=41)
    Wt = (rnd(2000, 1280, seed=42) / 36.0).to(BF)
    b = rnd(2000, seed=43)
    y = torch.empty(2, 2000, device=DEV)
    ops.gemv(x, Wt, b, y, silu_in=True)
    out.append(("gemv silu", rel(y, F.silu(x) @ Wt.float().t() + b), 1e-5))
    pred, tgt = rnd(2, 4, 64, 64, seed=44), rnd(2, 4, 64, 64, seed=45)
    loss = torch.zeros(1, device=DEV)
    dp = torch.empty_like(pred)
    ops.mse_loss(pred, tgt, loss, dp)
    pr = pred.clone().requires_grad_(True)
    lr = F.mse_loss(pr, tgt)
    lr.backward()
    out += [("mse loss", rel(loss, lr.detach().reshape(1)), 1e-5), ("mse dpred", rel(dp, pr.grad), 1e-5)]
    for vpred in (0, 1):
        lat, eu, ec = rnd(1, 4, 32, 32, seed=46), rnd(1, 4, 32, 32, seed=47), rnd(1, 4, 32, 32, seed=48)
        at, ap = 0.4, 0.6
        m = eu + 7.5 * (ec - eu)
        if vpred:
            x0 = at ** 0.5 * lat - (1 - at) ** 0.5 * m
            eps = at ** 0.5 * m + (1 - at) ** 0.5 * lat
        else:
            x0 = (lat - (1 - at) ** 0.5 * m) / at ** 0.5
            eps = m
        ref = ap ** 0.5 * x0 + (1 - ap) ** 0.5 * eps
        l2 = lat.clone()
        ops.cfg_ddim_step(l2, eu, ec, 7.5, at, ap, vpred)
        out.append((f"cfg_ddim vpred{vpred}", rel(l2, ref), 1e-5))
    src = rnd(100, 64, seed=49).to(BF)
    add = rnd(100, 64, seed=50).to(BF)
    dstbuf = torch.zeros(100, 192, dtype=BF, device=DEV)
    ops.copy2d(src, dstbuf[:, 64:128], add=add)
    out.append(("copy2d+add", rel(dstbuf[:, 64:128], src.float() + add.float()), 4e-3))
    out.append(("copy2d untouched", float(dstbuf[:, :64].abs().max() + dstbuf[:, 128:].abs().max()), 0.0))
    xf = rnd(1000, seed=51)
    xb = torch.empty(1000, dtype=BF, device=DEV)
    ops.cast_f32_bf16(xf, xb)
    xf2 = torch.empty(1000, device=DEV)
    ops.cast_bf16_f32(xb, xf2)
    out.append(("cast roundtrip", rel(xf2, xf.to(BF).float()), 0.0))
    return out


def all_checks():
    """(callable, kwargs) list; sizes follow the SD-2.1 layer shapes."""
    L = []
    for (M, N, K) in [(4096, 320, 320), (1024, 640, 640), (256, 1280, 1280), (64, 1280, 1280), (77, 320, 1024),
                      (4096, 2560, 320), (4096, 320, 1280), (1000, 328, 192)]:
        L.append((check_gemm, dict(M=M, N=N, K=K, bias=True, resid=True)))
    for bn in (64, 128, 256):
        for sp in (1, 2, 4, 8):
            L.append((check_gemm, dict(M=512, N=640, K=1280, bias=True, rowbias=True, resid=True, force_bn=bn, force_split=sp)))
    # multicast clusters along M (force_split bits 4..7 = cluster size)
    for bn in (64, 128, 256):
        for mc in (1, 2, 4):
            L.append((check_gemm, dict(M=1536, N=640, K=640, bias=True, resid=True, force_bn=bn, force_split=1 + 16 * mc)))
    L.append((check_gemm, dict(M=300, N=328, K=192, bias=True, resid=True, force_bn=128, force_split=1 + 16 * 2)))
    L.append((check_gemm, dict(M=700, N=328, K=192, bias=True, rowbias=True, resid=True, force_bn=64, force_split=1 + 16 * 4)))
    L.append((check_gemm, dict(M=20000, N=320, K=320, bias=True, resid=True, force_bn=128, force_split=1 + 16 * 4)))
    L.append((check_gemm, dict(M=200, N=328, K=1024, bias=True, resid=True, force_bn=128, force_split=4)))
    L.append((check_gemm, dict(M=77, N=1024, K=1280, out_fp32=True, force_bn=64, force_split=8)))
    L.append((check_gemm, dict(M=6000, N=2560, K=320, bias=True, resid=True, force_bn=256, force_split=1)))
    L.append((check_gemm, dict(M=20000, N=320, K=320, bias=True, resid=True, force_bn=128, force_split=1)))
    # CTA pairs (cta_group::2): force_split = 1 (no split-K) + 256 (pairs on)
    L.append((check_gemm, dict(M=4096, N=320, K=320, bias=True, resid=True, force_bn=128, force_split=257)))
    L.append((check_gemm, dict(M=20480, N=320, K=320, bias=True, resid=True, force_bn=128, force_split=257)))
    L.append((check_gemm, dict(M=1024, N=2560, K=640, bias=True, force_bn=256, force_split=257)))
    L.append((check_gemm, dict(M=512, N=328, K=1024, bias=True, resid=True, force_bn=256, force_split=257)))
    # BN 160 (split-K only) and BN 192
    L.append((check_gemm, dict(M=1024, N=640, K=2560, bias=True, resid=True, force_bn=160, force_split=4)))
    L.append((check_gemm, dict(M=300, N=328, K=1024, bias=True, resid=True, force_bn=160, force_split=2)))
    L.append((check_gemm, dict(M=1024, N=1920, K=640, bias=True, resid=True, force_bn=192, force_split=1)))
    L.append((check_gemm, dict(M=1024, N=1920, K=640, bias=True, resid=True, force_bn=192, force_split=257)))
    L.append((check_gemm, dict(M=256, N=3840, K=1280, bias=True, force_bn=192, force_split=2)))
    L.append((check_gemm, dict(M=300, N=320, K=640, out_fp32=True, strided=True, bias=True)))
    L.append((check_gemm, dict(M=300, N=320, K=640, strided=True, resid=True, force_split=2)))
    for (nb, H, W, Cc, N) in [(1, 64, 64, 320, 320), (1, 32, 32, 640, 640), (1, 16, 16, 1280, 1280),
                              (1, 8, 8, 1280, 1280), (2, 16, 16, 128, 64), (1, 48, 64, 64, 64), (2, 8, 8, 2560, 1280),
                              (1, 64, 64, 960, 320)]:
        L.append((check_conv, dict(nb=nb, H=H, W=W, Cc=Cc, N=N, rowbias=True, resid=True, dgrad=(Cc <= 1280))))
    L.append((check_conv, dict(nb=1, H=64, W=64, Cc=320, N=320, force_bn=256, force_split=257, rowbias=True, resid=True)))
    L.append((check_conv, dict(nb=1, H=32, W=32, Cc=640, N=640, force_bn=128, force_split=257, rowbias=True, resid=True)))
    L.append((check_conv, dict(nb=4, H=64, W=64, Cc=64, N=192, force_bn=128, force_split=257, rowbias=True, resid=True)))
    L.append((check_conv, dict(nb=1, H=64, W=64, Cc=320, N=320, force_bn=160, force_split=2, rowbias=True, resid=True)))
    L.append((check_conv, dict(nb=1, H=16, W=16, Cc=1280, N=1280, force_bn=160, force_split=4, rowbias=True, resid=True)))
    L.append((check_conv, dict(nb=1, H=16, W=16, Cc=640, N=320, force_bn=64, force_split=4)))
    L.append((check_conv, dict(nb=1, H=32, W=32, Cc=320, N=320, force_bn=256, force_split=1, rowbias=True, resid=True)))
    L.append((check_conv, dict(nb=2, H=8, W=8, Cc=1280, N=1280, force_bn=128, force_split=8, rowbias=True, resid=True)))
    L.append((check_conv, dict(nb=3, H=64, W=64, Cc=64, N=192, force_bn=64, force_split=1, rowbias=True, resid=True)))
    L.append((check_conv, dict(nb=1, H=12, W=16, Cc=128, N=128, force_bn=128, force_split=2, rowbias=True, resid=True)))
    L.append((check_conv, dict(nb=3, H=32, W=32, Cc=64, N=192, force_bn=64, force_split=1 + 16 * 4, rowbias=True, resid=True)))
    L.append((check_conv, dict(nb=1, H=24, W=16, Cc=128, N=320, force_bn=128, force_split=1 + 16 * 2, rowbias=True, resid=True)))
    L.append((check_conv, dict(nb=1, H=64, W=64, Cc=320, N=320, force_bn=256, force_split=1 + 16 * 4, rowbias=True, resid=True)))
    for (nb, hw, Cc, silu) in [(1, 4096, 320, True), (2, 1024, 640, False), (1, 64, 2560, True), (1, 256, 1920, True),
                               (1, 4096, 960, True), (1, 4096, 640, False), (1, 256, 2560, True), (2, 6912, 960, True),
                               (1, 3072, 320, True), (1, 16, 128, True), (3, 4, 512, False), (1, 1024, 1280, True)]:
        L.append((check_groupnorm, dict(nb=nb, hw=hw, Cc=Cc, silu=silu, eps=1e-5 if silu else 1e-6)))
    for (rows, Cc) in [(4096, 320), (1024, 640), (77, 1280)]:
        L.append((check_layernorm, dict(rows=rows, Cc=Cc)))
    for (rows, Cc) in [(1232, 1024), (77, 128), (300, 2048)]:
        L.append((check_layernorm_f32, dict(rows=rows, Cc=Cc)))
    L.append((check_gemm_f32_residual, dict(M=1232, N=1024, K=1024)))
    L.append((check_gemm_f32_residual, dict(M=1232, N=1024, K=4096)))
    L.append((check_gemm_f32_residual, dict(M=154, N=128, K=256, force_bn=64, force_split=2)))
    L.append((check_gemm_f32_residual, dict(M=300, N=1024, K=1024, force_bn=128, force_split=1)))
    L.append((check_geglu, dict(rows=1024, Fd=2560)))
    for (nb, heads, n) in [(1, 5, 4096), (1, 10, 1024), (2, 20, 256), (1, 20, 64), (1, 4, 3072)]:
        L.append((check_attention, dict(nb=nb, heads=heads, nq=n, nk=n, strided=(n == 1024))))
    # shapes whose work items leave a small last wave on 148 SMs: balanced forward schedule (key-split leftover items)
    L.append((check_attention, dict(nb=3, heads=5, nq=4096, nk=4096, bwd=False)))           # 480 items: 3 rounds + 36 left
    L.append((check_attention, dict(nb=1, heads=5, nq=4000, nk=4000)))                      # ragged last tiles, 160 items
    L.append((check_attention, dict(nb=2, heads=10, nq=1024, nk=1024, strided=True)))       # 160 items, 8 key blocks
    L.append((check_attention, dict(nb=1, heads=10, nq=1024, nk=1024)))                     # bwd: 80 + 80 items = 148 + 12
    L.append((check_attention, dict(nb=1, heads=3, nq=3200, nk=3200, causal=True)))         # causal across key / query parts
    L.append((check_attention, dict(nb=10, heads=16, nq=128, nk=1024, bwd=False, causal=True)))  # causal + key parts with no visible key
    for (nb, heads, n) in [(1, 5, 4096), (2, 10, 1024), (1, 20, 256), (1, 20, 64)]:
        L.append((check_attention, dict(nb=nb, heads=heads, nq=n, nk=77, use_acc=True)))
    L.append((check_attention, dict(nb=1, heads=2, nq=200, nk=77, use_acc=False)))
    # causal mask (CLIP text encoder: 77 tokens; and multi-tile shapes crossing the diagonal)
    L.append((check_attention, dict(nb=16, heads=16, nq=77, nk=77, strided=True, causal=True)))
    L.append((check_attention, dict(nb=2, heads=2, nq=300, nk=300, causal=True)))
    L.append((check_attention, dict(nb=1, heads=1, nq=128, nk=128, causal=True)))
    L.append((check_resample, dict(nb=1, H=16, W=16, Cc=128)))
    L.append((check_resample, dict(nb=2, H=12, W=16, Cc=64)))
    L.append((check_edge_convs, dict(nb=2, H=32, W=32, Cw=320)))
    L.append((check_misc, dict()))
    return L
