"""TEST INFRASTRUCTURE — torch (CPU) restatement of the few `view_neti_b200.ops` entry points the VAE launch sequence
uses, so the HOST logic of models/vae.py (weight layouts, the quant_conv fold, buffer roles and their reuse, call order,
the contracts the kernels put on their callers) is checked against the oracle where there is no GPU.  Never on a product
path: tests swap it in for `ops` by monkeypatching.  Same storage contract as the kernels: NHWC bf16 activations, fp32
math inside an op.  The contracts the real kernels rely on are asserted here:
  * GroupNorm statistics slots are zero on entry, `partials` is all-0xFF (or None);
  * a GEMM B operand written by an earlier op of the sequence needs b_dynamic=True (include/viewneti.h, vn_gemm);
  * outputs never alias inputs."""
from __future__ import annotations

import torch
import torch.nn.functional as F

from view_neti_b200 import _abi  # noqa: F401  (VNError for the code under test)

BF = torch.bfloat16
_written = set()


def reset():
    _written.clear()


def _mark(t: torch.Tensor):
    _written.add(t.untyped_storage().data_ptr())


def _no_alias(out: torch.Tensor, *ins):
    for t in ins:
        if t is not None:
            assert t.untyped_storage().data_ptr() != out.untyped_storage().data_ptr(), "output aliases an input"


class Workspace:
    def __init__(self, max_m, max_n, device, dkv_elems: int = 0):
        self.bytes = 256


def groupnorm_partial_floats(nb: int) -> int:
    return 64


def memset(t: torch.Tensor, byte_value: int) -> None:
    t.view(torch.uint8).fill_(byte_value)


def groupnorm_fwd(x, gamma, beta, eps, silu, y, nb, hw, groups, stats, partials):
    assert x.shape == (nb, hw, x.shape[-1]) and x.dtype == BF and y.dtype == BF
    assert stats.shape == (nb, groups, 2) and stats.dtype == torch.float64 and float(stats.abs().sum()) == 0.0
    if partials is not None:
        assert bool((partials.view(torch.uint8) == 0xFF).all()), "partials must be preset to 0xff"
        partials.zero_()                                         # the kernel consumes its slot
    _no_alias(y, x)
    C = x.shape[-1]
    xf = x.float().view(nb, hw, groups, C // groups)
    stats[..., 0] = xf.double().sum(dim=(1, 3))
    stats[..., 1] = (xf.double() ** 2).sum(dim=(1, 3))
    out = F.group_norm(x.float().transpose(1, 2), groups, gamma, beta, eps).transpose(1, 2)
    y.copy_(F.silu(out) if silu else out)
    _mark(y)


def conv3x3(x, Wk, D, *, bias=None, rowbias=None, R=None, ws=None, force_bn=0, force_split=0, stride=1, pad=1):
    nb, H, W, C = x.shape
    N = Wk.shape[0]
    assert Wk.shape == (N, 9 * C) and rowbias is None and C % 64 == 0 and N % 8 == 0
    _no_alias(D, x, R)
    w = Wk.float().view(N, 3, 3, C).permute(0, 3, 1, 2)
    xi = x.float().permute(0, 3, 1, 2)
    if stride == 2 and pad == 0:
        xi, pad = F.pad(xi, (0, 1, 0, 1)), 0
    o = F.conv2d(xi, w, bias, stride=stride, padding=pad).permute(0, 2, 3, 1)
    assert D.shape == o.shape
    D.copy_(o + R.float() if R is not None else o)
    _mark(D)


def gemm(A, B, D, *, bias=None, rowbias=None, rows_per_batch=0, R=None, ws=None, force_bn=0, force_split=0,
         b_dynamic=False):
    assert A.dim() == 2 and B.dim() == 2 and A.shape[1] == B.shape[1] and D.shape == (A.shape[0], B.shape[0])
    assert A.dtype == BF and B.dtype == BF and rowbias is None
    assert A.shape[1] % 64 == 0 and B.shape[0] % 8 == 0
    for t in (A, B, D):
        assert t.stride(-1) == 1 and t.stride(0) % 8 == 0 and t.data_ptr() % 16 == 0
    assert b_dynamic or B.untyped_storage().data_ptr() not in _written, "computed B operand needs b_dynamic=True"
    _no_alias(D, A, B, R)
    o = A.float() @ B.float().t()
    if bias is not None:
        o = o + bias
    if R is not None:
        o = o + R.float()
    D.copy_(o)
    _mark(D)


def im2col_s2_pad0(x, col):
    nb, H, W, C = x.shape
    Ho, Wo = H // 2, W // 2
    assert col.shape == (nb * Ho * Wo, 9 * C)
    xp = F.pad(x.float().permute(0, 3, 1, 2), (0, 1, 0, 1))
    u = F.unfold(xp, 3, stride=2)                                # [nb, C*9, Ho*Wo], channel-major (c*9 + tap)
    u = u.view(nb, C, 9, Ho * Wo).permute(0, 3, 2, 1).reshape(nb * Ho * Wo, 9 * C)
    col.copy_(u)
    _mark(col)


def im2col_thin(x, col):
    nb, Ct, H, W = x.shape
    assert x.dtype == torch.float32 and x.is_contiguous() and col.dtype == BF
    assert col.shape[0] == nb * H * W and col.shape[1] % 64 == 0 and col.shape[1] >= 9 * Ct
    u = F.unfold(x, 3, padding=1).view(nb, Ct, 9, H * W).permute(0, 3, 2, 1).reshape(nb * H * W, 9 * Ct)
    col.zero_()
    col[:, : 9 * Ct] = u
    _mark(col)


def softmax_rows(S, P, scale):
    assert S.dtype == torch.float32 and P.dtype == BF and S.shape == P.shape and S.shape[1] % 4 == 0
    P.copy_(torch.softmax(S * scale, dim=-1))
    _mark(P)


def conv_in_fwd(x, w, bias, y):
    assert x.dtype == torch.float32 and x.is_contiguous() and y.dtype == BF and y.shape[-1] % 8 == 0
    y.copy_(F.conv2d(x, w, bias, padding=1).permute(0, 2, 3, 1))
    _mark(y)


def conv_out_fwd(x, w, bias, y):
    assert x.dtype == BF and y.dtype == torch.float32 and y.is_contiguous() and y.shape[1] <= 8
    y.copy_(F.conv2d(x.float().permute(0, 3, 1, 2), w, bias, padding=1))


def upsample2x_fwd(x, y):
    nb, H, W, C = x.shape
    assert y.shape == (nb, 2 * H, 2 * W, C)
    _no_alias(y, x)
    y.copy_(x.repeat_interleave(2, dim=1).repeat_interleave(2, dim=2))
    _mark(y)
