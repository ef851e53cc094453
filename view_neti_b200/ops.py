"""Tensor-level wrappers over the C-ABI (include/viewneti.h): torch tensors in, raw pointers across.

Activations are bf16, channel-last, possibly channel-slices of a wider buffer: every wrapper takes
"rows x C" views whose last dim is contiguous and passes the row stride as `ld`.  Nothing here
computes: each function is exactly one library call (there is no CPU / eager fallback).
"""
from __future__ import annotations

import contextlib

import ctypes as C
from typing import Optional

import torch

from . import _abi
from ._abi import AttnDesc, GemmDesc, check, ptr, stream

BF16 = torch.bfloat16


def _ld(t: torch.Tensor) -> int:
    """Row stride (elements) of a [..., rows, C] view with contiguous last dim and collapsible leading dims."""
    assert t.stride(-1) == 1, "last dim must be contiguous"
    ld = t.stride(-2)
    exp = ld * t.shape[-2]
    for i in range(t.dim() - 3, -1, -1):          # leading dims must collapse onto the row dim
        if t.shape[i] != 1:
            assert t.stride(i) == exp, f"view does not collapse to rows: shape {tuple(t.shape)} stride {t.stride()}"
            exp *= t.shape[i]
    return ld


def _pix_ld(x: torch.Tensor) -> int:
    """Pixel stride (elements) of an NHWC view [nb,H,W,C] whose pixels are evenly strided; size-1 dims carry no
    stride information in torch, so they are skipped."""
    nb, H, W, Cc = x.shape
    assert x.stride(3) == 1
    ld = x.stride(2) if W > 1 else (x.stride(1) if H > 1 else (x.stride(0) if nb > 1 else Cc))
    if W > 1 and H > 1:
        assert x.stride(1) == W * ld, f"rows of the NHWC view are not evenly strided: {tuple(x.shape)} {x.stride()}"
    if nb > 1 and H * W > 1:
        assert x.stride(0) == H * W * ld, f"images of the NHWC view are not evenly strided: {tuple(x.shape)} {x.stride()}"
    return ld


def tuning_key(mode: int, M: int, N: int, K: int, H: int, W: int) -> str:
    return f"{mode}:{M}:{N}:{K}:{H}:{W}"


def _load_tuning():
    """(BN, split) choices measured on B200 by scripts/gemm_autotune.py for the SD-2.1 layer shapes; shapes not in the
    table use the library's cost model."""
    import json
    import os
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "gemm_tuning.json")
    if not os.path.exists(path):
        return {}
    with open(path) as f:
        return {k: (int(v[0]), int(v[1])) for k, v in json.load(f).items()}


TUNING = _load_tuning()


class Workspace:
    """Split-K scratch for vn_gemm (zero on entry, left zero by every call) + fp32 dK/dV accumulator."""

    def __init__(self, max_m: int, max_n: int, device, dkv_elems: int = 0):
        lib = _abi.load()
        self.bytes = int(lib.vn_gemm_workspace_bytes(max_m, max_n))
        self.buf = torch.zeros(self.bytes, dtype=torch.uint8, device=device)
        self.dkv = torch.zeros(max(dkv_elems, 1), dtype=torch.float64, device=device)


def gemm(A: torch.Tensor, B: torch.Tensor, D: torch.Tensor, *, bias=None, rowbias=None, rows_per_batch: int = 0,
         R=None, ws: Optional[Workspace] = None, force_bn: int = 0, force_split: int = 0,
         b_dynamic: bool = False) -> torch.Tensor:
    """D[M,N] = A[M,K] @ B[N,K]^T (+bias[N]) (+rowbias[row // rows_per_batch, N]) (+R[M,N]).

    b_dynamic: B was written by an earlier launch on this stream (not a frozen weight).  The kernel requests its first
    B tiles before it waits for its predecessor (include/viewneti.h, vn_gemm), so such a product is launched without
    programmatic dependent launch."""
    if b_dynamic and _PDL:
        with pdl_off():
            return gemm(A, B, D, bias=bias, rowbias=rowbias, rows_per_batch=rows_per_batch, R=R, ws=ws,
                        force_bn=force_bn, force_split=force_split)
    lib = _abi.load()
    M, K = A.shape[-2] * (A.numel() // (A.shape[-1] * A.shape[-2])), A.shape[-1]
    N = B.shape[0]
    assert B.shape[1] == K and A.dtype == BF16 and B.dtype == BF16
    d = GemmDesc()
    d.mode = 0
    d.M, d.N, d.K = M, N, K
    d.A, d.lda = ptr(A), _ld(A)
    d.B, d.ldb = ptr(B), B.stride(0)
    d.D, d.ldd = ptr(D), _ld(D)
    d.out_fp32 = 1 if D.dtype == torch.float32 else 0
    d.bias = ptr(bias)
    if rowbias is not None:
        d.rowbias, d.ld_rowbias, d.rows_per_batch = ptr(rowbias), rowbias.stride(0), rows_per_batch
    if R is not None:
        d.R, d.ldr = ptr(R), _ld(R)
        if R.dtype == torch.float32:
            assert D.dtype == torch.float32, "an fp32 residual needs an fp32 output"
            d.r_fp32 = 1
    if ws is not None:
        d.workspace, d.workspace_bytes = ptr(ws.buf), ws.bytes
    if not force_bn and not force_split:
        force_bn, force_split = TUNING.get(tuning_key(0, M, N, K, 0, 0), (0, 0))
    d.force_bn, d.force_split = force_bn, force_split
    check(lib.vn_gemm(C.byref(d), stream()), "vn_gemm")
    return D


def conv3x3(x: torch.Tensor, Wk: torch.Tensor, D: torch.Tensor, *, bias=None, rowbias=None, R=None,
            ws: Optional[Workspace] = None, force_bn: int = 0, force_split: int = 0, stride: int = 1,
            pad: int = 1) -> torch.Tensor:
    """Implicit-GEMM 3x3 conv.  x [nb,H,W,C] NHWC view, Wk [N, 9*C] (k = tap*C + c), D [nb,Ho,Wo,N].
    stride 1 / pad 1; stride 2 / pad 1 (Downsample2D); stride 2 / pad 0 = zero beyond the far edge only (VAE encoder)."""
    lib = _abi.load()
    nb, H, W, Cc = x.shape
    N = Wk.shape[0]
    assert Wk.shape[1] == 9 * Cc and (stride, pad) in ((1, 1), (2, 1), (2, 0))
    Ho, Wo = (H, W) if stride == 1 else (((H - 1) // 2 + 1, (W - 1) // 2 + 1) if pad else ((H - 2) // 2 + 1, (W - 2) // 2 + 1))
    assert tuple(D.shape) == (nb, Ho, Wo, N), (tuple(D.shape), (nb, Ho, Wo, N))
    d = GemmDesc()
    d.mode = 1 if stride == 1 else (2 if pad else 3)
    d.M, d.N, d.K = nb * Ho * Wo, N, 9 * Cc
    d.nb, d.H, d.W, d.C = nb, H, W, Cc
    d.A, d.lda = ptr(x), _pix_ld(x)
    d.B, d.ldb = ptr(Wk), Wk.stride(0)
    d.D, d.ldd = ptr(D), _pix_ld(D)
    d.out_fp32 = 1 if D.dtype == torch.float32 else 0
    d.bias = ptr(bias)
    if rowbias is not None:
        d.rowbias, d.ld_rowbias, d.rows_per_batch = ptr(rowbias), rowbias.stride(0), Ho * Wo
    if R is not None:
        d.R, d.ldr = ptr(R), _pix_ld(R)
    if ws is not None:
        d.workspace, d.workspace_bytes = ptr(ws.buf), ws.bytes
    if not force_bn and not force_split:
        force_bn, force_split = TUNING.get(tuning_key(d.mode, nb * Ho * Wo, N, 9 * Cc, H, W), (0, 0))
    d.force_bn, d.force_split = force_bn, force_split
    check(lib.vn_gemm(C.byref(d), stream()), "vn_gemm(conv)")
    return D


# ---- GroupNorm / LayerNorm / GEGLU -------------------------------------------------------------
def groupnorm_stats(x, nb, hw, groups, stats):
    check(_abi.load().vn_groupnorm_stats(ptr(x), _ld(x), nb, hw, x.shape[-1], groups, ptr(stats), stream()), "gn_stats")


def groupnorm_apply(x, stats, gamma, beta, eps, silu, y, nb, hw, groups):
    check(_abi.load().vn_groupnorm_apply(ptr(x), _ld(x), ptr(stats), ptr(gamma), ptr(beta), eps, int(silu), ptr(y),
                                         _ld(y), nb, hw, x.shape[-1], groups, stream()), "gn_apply")


def groupnorm_bwd(x, dy, stats, red, gamma, beta, eps, silu, dx, nb, hw, groups, add1=None, add2=None):
    lib = _abi.load()
    Cc = x.shape[-1]
    check(lib.vn_groupnorm_bwd_stats(ptr(x), _ld(x), ptr(dy), _ld(dy), ptr(stats), ptr(gamma), ptr(beta), eps, int(silu),
                                     ptr(red), nb, hw, Cc, groups, stream()), "gn_bwd_stats")
    check(lib.vn_groupnorm_bwd_apply(ptr(x), _ld(x), ptr(dy), _ld(dy), ptr(stats), ptr(red), ptr(gamma), ptr(beta), eps,
                                     int(silu), ptr(add1), _ld(add1) if add1 is not None else 0, ptr(add2),
                                     _ld(add2) if add2 is not None else 0, ptr(dx), _ld(dx), nb, hw, Cc, groups,
                                     stream()), "gn_bwd_apply")


def groupnorm_partial_floats(nb: int) -> int:
    """Floats of `partials` one fused GroupNorm call needs (64 per CTA, at most one CTA per SM)."""
    return int(_abi.load().vn_groupnorm_partial_floats(nb))


def groupnorm_fwd(x, gamma, beta, eps, silu, y, nb, hw, groups, stats, partials):
    """GroupNorm(+SiLU) in one launch.  `partials`: float32 [groupnorm_partial_floats(nb)], every byte 0xff on entry
    (None -> two-kernel form); `stats` fp64 [nb, groups, 2] is written."""
    check(_abi.load().vn_groupnorm_fwd(ptr(x), _ld(x), ptr(gamma), ptr(beta), eps, int(silu), ptr(y), _ld(y), nb, hw,
                                       x.shape[-1], groups, ptr(stats), ptr(partials), stream()), "gn_fwd")


def groupnorm_bwd_fused(x, dy, stats, red, partials, gamma, beta, eps, silu, dx, nb, hw, groups, add1=None, add2=None):
    """dx = GN^T(dy) (+add1) (+add2) in one launch; `partials` as in groupnorm_fwd, `red` is written."""
    check(_abi.load().vn_groupnorm_bwd(ptr(x), _ld(x), ptr(dy), _ld(dy), ptr(stats), ptr(gamma), ptr(beta), eps,
                                       int(silu), ptr(add1), _ld(add1) if add1 is not None else 0, ptr(add2),
                                       _ld(add2) if add2 is not None else 0, ptr(dx), _ld(dx), nb, hw, x.shape[-1],
                                       groups, ptr(red), ptr(partials), stream()), "gn_bwd")


def memset(t: torch.Tensor, byte_value: int) -> None:
    """Fill a contiguous tensor with one byte value (a memset node when captured in a CUDA graph)."""
    assert t.is_contiguous()
    check(_abi.load().vn_memset(ptr(t), int(byte_value), t.numel() * t.element_size(), stream()), "memset")


def set_groupnorm_fused(enabled: bool) -> None:
    _abi.load().vn_set_groupnorm_fused(1 if enabled else 0)


def layernorm_fwd(x, gamma, beta, eps, y, stats, rows):
    check(_abi.load().vn_layernorm_fwd(ptr(x), _ld(x), ptr(gamma), ptr(beta), eps, ptr(y), _ld(y), ptr(stats), rows,
                                       x.shape[-1], stream()), "ln_fwd")


def layernorm_bwd(x, dy, gamma, stats, dx, rows, add=None):
    check(_abi.load().vn_layernorm_bwd(ptr(x), _ld(x), ptr(dy), _ld(dy), ptr(gamma), ptr(stats), ptr(add),
                                       _ld(add) if add is not None else 0, ptr(dx), _ld(dx), rows, x.shape[-1],
                                       stream()), "ln_bwd")


def layernorm_fwd_f32(x, gamma, beta, eps, y, stats, rows):
    """fp32 rows in, bf16 normalised rows out (the text encoder's fp32 residual stream)."""
    assert x.dtype == torch.float32 and y.dtype == BF16
    check(_abi.load().vn_layernorm_fwd_f32(ptr(x), _ld(x), ptr(gamma), ptr(beta), eps, ptr(y), _ld(y), ptr(stats), rows,
                                           x.shape[-1], stream()), "ln_fwd_f32")


def layernorm_bwd_f32(x, dy, gamma, stats, dx, rows, add=None, dx_bf16=None):
    """dx (fp32) = LN^T(dy) (+ add, fp32); optionally also a bf16 copy of dx (A operand of the next dgrad GEMM)."""
    assert x.dtype == torch.float32 and dx.dtype == torch.float32 and (add is None or add.dtype == torch.float32)
    check(_abi.load().vn_layernorm_bwd_f32(ptr(x), _ld(x), ptr(dy), _ld(dy), ptr(gamma), ptr(stats), ptr(add),
                                           _ld(add) if add is not None else 0, ptr(dx), _ld(dx), ptr(dx_bf16),
                                           _ld(dx_bf16) if dx_bf16 is not None else 0, rows, x.shape[-1], stream()),
          "ln_bwd_f32")


def geglu_fwd(h, y, rows):
    check(_abi.load().vn_geglu_fwd(ptr(h), _ld(h), ptr(y), _ld(y), rows, y.shape[-1], stream()), "geglu_fwd")


def geglu_bwd(h, dy, dh, rows):
    check(_abi.load().vn_geglu_bwd(ptr(h), _ld(h), ptr(dy), _ld(dy), ptr(dh), _ld(dh), rows, dy.shape[-1], stream()),
          "geglu_bwd")


# ---- attention ----------------------------------------------------------------------------------
def _attn_desc(q, k, v, o, lse, heads, scale) -> AttnDesc:
    """q,o: [nb, nq, heads*64] views; k,v: [nb, nk, heads*64] views (last dim contiguous)."""
    d = AttnDesc()
    d.nb, d.nq, d.nk, d.heads, d.scale = q.shape[0], q.shape[1], k.shape[1], heads, scale
    d.q, d.ldq, d.bsq = ptr(q), q.stride(1), q.stride(0)
    d.k, d.ldk, d.bsk = ptr(k), k.stride(1), k.stride(0)
    d.v, d.ldv, d.bsv = ptr(v), v.stride(1), v.stride(0)
    d.o, d.ldo, d.bso = ptr(o), o.stride(1), o.stride(0)
    d.lse = ptr(lse)
    return d


def attention_fwd_workspace_bytes(nb: int, heads: int, nq: int, nk: int) -> int:
    """Scratch the forward wants for its leftover-item key split (0: the shape runs one CTA per item)."""
    return int(_abi.load().vn_attention_fwd_workspace_bytes(nb, heads, nq, nk))


def attention_fwd(q, k, v, o, lse, heads, scale=0.125, causal=False, ws: Optional[torch.Tensor] = None):
    """ws: optional scratch (uint8 / any dtype, >= attention_fwd_workspace_bytes) that lets the kernel balance the work
    items over the SMs; without it every item runs whole on its own CTA."""
    d = _attn_desc(q, k, v, o, lse, heads, scale)
    d.causal = int(causal)
    if ws is not None:
        d.ws, d.ws_bytes = ptr(ws), ws.numel() * ws.element_size()
    check(_abi.load().vn_attention_fwd(C.byref(d), stream()), "attention_fwd")


def attention_bwd_workspace_bytes(nb: int, heads: int, nq: int, nk: int, has_dq: bool = True) -> int:
    return int(_abi.load().vn_attention_bwd_workspace_bytes(nb, heads, nq, nk, int(has_dq)))


def attention_bwd(q, k, v, o, lse, d_o, delta, dq, dk, dv, heads, scale=0.125, dkv_acc=None, causal=False,
                  ws: Optional[torch.Tensor] = None, defer_finish: bool = False):
    """defer_finish (with dkv_acc): dK / dV stay in the fp64 accumulator; pass the returned descriptor to
    attention_dkv_finish - on any stream ordered after this call - to convert them and zero the accumulator again."""
    d = _attn_desc(q, k, v, o, lse, heads, scale)
    d.causal = int(causal)
    if ws is not None:
        d.ws, d.ws_bytes = ptr(ws), ws.numel() * ws.element_size()
    d.d_o, d.lddo, d.bsdo = ptr(d_o), d_o.stride(1), d_o.stride(0)
    d.delta = ptr(delta)
    if dq is not None:
        d.dq, d.lddq, d.bsdq = ptr(dq), dq.stride(1), dq.stride(0)
    d.dk, d.lddk, d.bsdk = ptr(dk), dk.stride(1), dk.stride(0)
    d.dv, d.lddv, d.bsdv = ptr(dv), dv.stride(1), dv.stride(0)
    d.dkv_acc = ptr(dkv_acc)
    d.defer_dkv_finish = 1 if (defer_finish and dkv_acc is not None) else 0
    check(_abi.load().vn_attention_bwd(C.byref(d), stream()), "attention_bwd")
    return d


def attention_dkv_finish(d) -> None:
    """Second half of attention_bwd(..., defer_finish=True): fp64 accumulator -> bf16 dK / dV on the CURRENT stream."""
    check(_abi.load().vn_attention_dkv_finish(C.byref(d), stream()), "attention_dkv_finish")


def seq_attention_fwd(q, k, v, o, lse, heads, scale=0.125, causal=True):
    """Short-sequence (<= 128 tokens) attention, optional causal mask: CLIPAttention core of the text transformer."""
    d = _attn_desc(q, k, v, o, lse, heads, scale)
    check(_abi.load().vn_seq_attention_fwd(C.byref(d), int(causal), stream()), "seq_attention_fwd")


def seq_attention_bwd(q, k, v, o, lse, d_o, dq, dk, dv, heads, scale=0.125, causal=True):
    d = _attn_desc(q, k, v, o, lse, heads, scale)
    d.d_o, d.lddo, d.bsdo = ptr(d_o), d_o.stride(1), d_o.stride(0)
    d.dq, d.lddq, d.bsdq = ptr(dq), dq.stride(1), dq.stride(0)
    d.dk, d.lddk, d.bsdk = ptr(dk), dk.stride(1), dk.stride(0)
    d.dv, d.lddv, d.bsdv = ptr(dv), dv.stride(1), dv.stride(0)
    check(_abi.load().vn_seq_attention_bwd(C.byref(d), int(causal), stream()), "seq_attention_bwd")


def gelu_fwd(h, y, rows):
    check(_abi.load().vn_gelu_fwd(ptr(h), _ld(h), ptr(y), _ld(y), rows, h.shape[-1], stream()), "gelu_fwd")


def gelu_bwd(h, dy, dh, rows):
    check(_abi.load().vn_gelu_bwd(ptr(h), _ld(h), ptr(dy), _ld(dy), ptr(dh), _ld(dh), rows, h.shape[-1], stream()),
          "gelu_bwd")


# ---- resampling / edge convs / glue -------------------------------------------------------------
def upsample2x_fwd(x, y):
    nb, H, W, Cc = x.shape
    check(_abi.load().vn_upsample2x_fwd(ptr(x), _pix_ld(x), ptr(y), _pix_ld(y), nb, H, W, Cc, stream()), "upsample")


def upsample2x_bwd(dy, dx):
    nb, H, W, Cc = dx.shape
    check(_abi.load().vn_upsample2x_bwd(ptr(dy), _pix_ld(dy), ptr(dx), _pix_ld(dx), nb, H, W, Cc, stream()),
          "upsample_bwd")


def im2col_s2(x, col):
    nb, H, W, Cc = x.shape
    check(_abi.load().vn_im2col_s2(ptr(x), _pix_ld(x), ptr(col), nb, H, W, Cc, stream()), "im2col_s2")


def im2col_s2_pad0(x, col):
    """VAE-encoder downsample: zero pad (0,1,0,1), 3x3 taps at stride 2.  col [nb*(H//2)*(W//2), 9*C]."""
    nb, H, W, Cc = x.shape
    check(_abi.load().vn_im2col_s2_pad0(ptr(x), _pix_ld(x), ptr(col), nb, H, W, Cc, stream()), "im2col_s2_pad0")


def im2col_thin(x, col):
    """x NCHW fp32 [nb,Ct,H,W] -> col bf16 [nb*H*W, Kpad] (k = tap*Ct + ct, zero-padded to Kpad % 64 == 0)."""
    nb, Ct, H, W = x.shape
    assert x.dtype == torch.float32 and x.is_contiguous() and col.dtype == BF16 and col.shape[0] == nb * H * W
    check(_abi.load().vn_im2col_thin(ptr(x), ptr(col), col.stride(0), nb, Ct, H, W, stream()), "im2col_thin")


def softmax_rows(S, P, scale):
    """P = softmax(scale * S) over the last axis; S fp32 [rows, cols], P bf16 [rows, cols]."""
    rows, cols = S.shape
    assert S.dtype == torch.float32 and P.dtype == torch.bfloat16 and P.shape == S.shape
    check(_abi.load().vn_softmax_rows(ptr(S), S.stride(0), ptr(P), P.stride(0), rows, cols, float(scale), stream()),
          "softmax_rows")


def col2im_s2(dcol, dx, add=None):
    nb, H, W, Cc = dx.shape
    check(_abi.load().vn_col2im_s2(ptr(dcol), ptr(add), _pix_ld(add) if add is not None else 0, ptr(dx), _pix_ld(dx),
                                   nb, H, W, Cc, stream()), "col2im_s2")


def conv_in_fwd(x, w, bias, y):
    nb, Cin, H, W = x.shape
    check(_abi.load().vn_conv_in_fwd(ptr(x), ptr(w), ptr(bias), ptr(y), _pix_ld(y), nb, Cin, H, W, y.shape[-1],
                                     stream()), "conv_in")


def conv_out_fwd(x, w, bias, y):
    nb, H, W, Cin = x.shape
    check(_abi.load().vn_conv_out_fwd(ptr(x), _pix_ld(x), ptr(w), ptr(bias), ptr(y), nb, Cin, H, W, y.shape[1],
                                      stream()), "conv_out")


def conv_out_bwd(dy, w, dx):
    nb, Cout, H, W = dy.shape
    check(_abi.load().vn_conv_out_bwd(ptr(dy), ptr(w), ptr(dx), _pix_ld(dx), nb, dx.shape[-1], H, W, Cout, stream()),
          "conv_out_bwd")


def nhwc_to_nchw_thin(src, out):
    """src fp32 [nb, H, W, Cpad] (first Ct channels used) -> out fp32 [nb, Ct, H, W]."""
    nb, Ct, H, W = out.shape
    assert src.dtype == torch.float32 and out.dtype == torch.float32 and out.is_contiguous() and src.stride(-1) == 1
    check(_abi.load().vn_nhwc_to_nchw_thin(ptr(src), _pix_ld(src), ptr(out), nb, Ct, H * W, stream()), "nhwc_to_nchw_thin")


def timestep_sinusoid(t, out):
    check(_abi.load().vn_timestep_sinusoid(ptr(t), ptr(out), out.shape[0], out.shape[1], stream()), "timestep")


GEMV_MAX_BATCH = 8          # kGemvMaxB in csrc/vn_small.cu: rows of x one launch keeps in registers


def gemv(x, W, bias, y, silu_in=False):
    """y[b] = W @ (silu?)(x[b]) + bias for a handful of rows (time embedding).  Batches beyond the kernel's 8 rows per
    launch (training batches > 8, CFG batches of sd_pipeline_call with num_images_per_prompt > 4) go in chunks."""
    lib = _abi.load()
    for b0 in range(0, x.shape[0], GEMV_MAX_BATCH):
        xb, yb = x[b0:b0 + GEMV_MAX_BATCH], y[b0:b0 + GEMV_MAX_BATCH]
        check(lib.vn_gemv(ptr(xb), xb.stride(0), ptr(W), ptr(bias), ptr(yb), yb.stride(0), xb.shape[0], W.shape[0],
                          W.shape[1], int(silu_in), stream()), "gemv")


def cast_f32_bf16(x, y):
    check(_abi.load().vn_cast_f32_bf16(ptr(x), ptr(y), x.numel(), stream()), "cast")


def cast_bf16_f32(x, y):
    check(_abi.load().vn_cast_bf16_f32(ptr(x), ptr(y), x.numel(), stream()), "cast")


def copy2d(src, dst, add=None):
    rows = src.numel() // src.shape[-1]
    check(_abi.load().vn_copy2d(ptr(src), _ld(src), ptr(add), _ld(add) if add is not None else 0, ptr(dst), _ld(dst),
                                rows, src.shape[-1], stream()), "copy2d")


def mse_loss(pred, target, loss, dpred=None, loss_scale=1.0):
    check(_abi.load().vn_mse_loss(ptr(pred), ptr(target), pred.numel(), loss_scale, ptr(loss), ptr(dpred), stream()),
          "mse")


def cfg_dpmpp_step(latents, eps_u, eps_c, x0_prev, guidance, p, q, A, B0, B1):
    """CFG combine + DPM-Solver++(2M) update in place; x0_prev carries the data prediction to the next step."""
    assert latents.dtype == torch.float32 and x0_prev.dtype == torch.float32 and x0_prev.numel() == latents.numel()
    check(_abi.load().vn_cfg_dpmpp_step(ptr(latents), ptr(eps_u), ptr(eps_c), ptr(x0_prev), latents.numel(), guidance,
                                        p, q, A, B0, B1, stream()), "cfg_dpmpp_step")


def cfg_ddim_step(latents, eps_u, eps_c, guidance, acp_t, acp_prev, vpred):
    check(_abi.load().vn_cfg_ddim_step(ptr(latents), ptr(eps_u), ptr(eps_c), latents.numel(), guidance, acp_t, acp_prev,
                                       int(vpred), stream()), "cfg_ddim")


def mapper_param_count(dim: int) -> int:
    return int(_abi.load().vn_mapper_param_count(dim))


def mapper_fwd(x, Wf, params, norm_scale, word, bypass, saved):
    B, nfeat = x.shape
    check(_abi.load().vn_mapper_fwd(ptr(x), ptr(Wf), ptr(params), float(norm_scale), ptr(word), ptr(bypass), ptr(saved),
                                    B, nfeat, word.shape[1], stream()), "mapper_fwd")


def mapper_bwd(d_word, d_bypass, params, saved, norm_scale, d_params, scratch):
    B, dim = d_word.shape
    check(_abi.load().vn_mapper_bwd(ptr(d_word), ptr(d_bypass), ptr(params), ptr(saved), float(norm_scale), ptr(d_params),
                                    ptr(scratch), B, dim, stream()), "mapper_bwd")


def adamw_step(params, grads, exp_avg, exp_avg_sq, lr, beta1, beta2, eps, weight_decay, step, grad_scale=1.0):
    check(_abi.load().vn_adamw_step(ptr(params), ptr(grads), ptr(exp_avg), ptr(exp_avg_sq), params.numel(), lr, beta1,
                                    beta2, eps, weight_decay, step, grad_scale, stream()), "adamw_step")


def launch_count() -> int:
    return int(_abi.load().vn_launch_count())


def launch_count_reset() -> None:
    _abi.load().vn_launch_count_reset()


_PDL = True


def set_pdl(enabled: bool) -> None:
    """Programmatic dependent launch on/off (off to time kernels in isolation, or around a launch whose early reads
    depend on its predecessor: gemm(b_dynamic=True))."""
    global _PDL
    _PDL = bool(enabled)
    _abi.load().vn_set_pdl(1 if enabled else 0)


@contextlib.contextmanager
def pdl_off():
    prev = _PDL
    set_pdl(False)
    try:
        yield
    finally:
        set_pdl(prev)
