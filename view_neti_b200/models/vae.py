"""SD-2.1 VAE (`AutoencoderKL`) on the library's kernels: the frozen image <-> latent maps around the hot path.

The reference calls it twice:
  * reference training/coach.py:165-169, every train step, no grad:
        latents = vae.encode(pixel_values).latent_dist.sample().detach() * vae.config.scaling_factor
  * reference sd_pipeline_call.py:115, once per generated image: `pipeline.decode_latents(latents)`.
`AutoencoderKL` below keeps those call shapes (`.encode(x).latent_dist.sample()`, `.decode(z).sample`,
`.config.scaling_factor`); `decode_latents` is the pipeline method the reference reaches through diffusers.

No checkpoint exists on this machine, so - as for the UNet (sd21.py) - weights are seeded random tensors of the exact
architecture under diffusers' state_dict keys (83 653 863 parameters); a real `vae/diffusion_pytorch_model` state
dict loads through the same constructor.  Topology restated from the public SD-2.1 `vae/config.json`; see oracle/vae.py.

Everything runs NHWC bf16 with fp32 accumulation: 3x3 convolutions and 1x1 / Linear layers on vn_gemm (tcgen05),
GroupNorm(+SiLU) on the GroupNorm kernels, the stride-2 encoder convolutions as im2col + GEMM, the two few-channel edge
convolutions on the thin-conv kernels.  The single-head 512-wide attention of the mid blocks does not fit the
head_dim-64 attention kernels; it is three GEMMs around a row softmax (vn_softmax_rows), 0.4 % of the encoder's FLOPs.
"""
from __future__ import annotations

import math
import os
from collections import OrderedDict
from dataclasses import dataclass
from types import SimpleNamespace
from typing import Dict, List, Optional, Tuple

import torch

from .. import ops

BF = torch.bfloat16
F32 = torch.float32


@dataclass(frozen=True)
class VAEConfig:
    in_channels: int = 3
    out_channels: int = 3
    latent_channels: int = 4
    block_out_channels: Tuple[int, ...] = (128, 256, 512, 512)
    layers_per_block: int = 2
    norm_num_groups: int = 32
    norm_eps: float = 1e-6
    scaling_factor: float = 0.18215
    sample_size: int = 768

    @property
    def downscale(self) -> int:
        return 2 ** (len(self.block_out_channels) - 1)


SD21_VAE = VAEConfig()
# same topology at a quarter of the width, for parity tests the CPU oracle finishes in seconds
TINY_VAE = VAEConfig(block_out_channels=(64, 128, 128, 128), sample_size=64)


# ---- parameter table (diffusers 0.14 AutoencoderKL state_dict keys) ----------------------------------
def _resnet(p: str, cin: int, cout: int) -> List[Tuple[str, Tuple[int, ...], str]]:
    t = [(f"{p}.norm1.weight", (cin,), "gamma"), (f"{p}.norm1.bias", (cin,), "beta"),
         (f"{p}.conv1.weight", (cout, cin, 3, 3), "w"), (f"{p}.conv1.bias", (cout,), "b"),
         (f"{p}.norm2.weight", (cout,), "gamma"), (f"{p}.norm2.bias", (cout,), "beta"),
         (f"{p}.conv2.weight", (cout, cout, 3, 3), "w_res"), (f"{p}.conv2.bias", (cout,), "b")]
    if cin != cout:
        t += [(f"{p}.conv_shortcut.weight", (cout, cin, 1, 1), "w"), (f"{p}.conv_shortcut.bias", (cout,), "b")]
    return t


def _mid(p: str, c: int) -> List[Tuple[str, Tuple[int, ...], str]]:
    a = f"{p}.attentions.0"
    t = [(f"{a}.group_norm.weight", (c,), "gamma"), (f"{a}.group_norm.bias", (c,), "beta")]
    for n in ("query", "key", "value"):
        t += [(f"{a}.{n}.weight", (c, c), "w"), (f"{a}.{n}.bias", (c,), "b")]
    t += [(f"{a}.proj_attn.weight", (c, c), "w_res"), (f"{a}.proj_attn.bias", (c,), "b")]
    return t + _resnet(f"{p}.resnets.0", c, c) + _resnet(f"{p}.resnets.1", c, c)


def param_table(cfg: VAEConfig = SD21_VAE) -> List[Tuple[str, Tuple[int, ...], str]]:
    ch, L = cfg.block_out_channels, cfg.latent_channels
    p: List[Tuple[str, Tuple[int, ...], str]] = [("encoder.conv_in.weight", (ch[0], cfg.in_channels, 3, 3), "w"),
                                                 ("encoder.conv_in.bias", (ch[0],), "b")]
    cin = ch[0]
    for i, cout in enumerate(ch):
        for j in range(cfg.layers_per_block):
            p += _resnet(f"encoder.down_blocks.{i}.resnets.{j}", cin, cout)
            cin = cout
        if i < len(ch) - 1:
            p += [(f"encoder.down_blocks.{i}.downsamplers.0.conv.weight", (cout, cout, 3, 3), "w"),
                  (f"encoder.down_blocks.{i}.downsamplers.0.conv.bias", (cout,), "b")]
    p += _mid("encoder.mid_block", ch[-1])
    p += [("encoder.conv_norm_out.weight", (ch[-1],), "gamma"), ("encoder.conv_norm_out.bias", (ch[-1],), "beta"),
          ("encoder.conv_out.weight", (2 * L, ch[-1], 3, 3), "w"), ("encoder.conv_out.bias", (2 * L,), "b"),
          ("quant_conv.weight", (2 * L, 2 * L, 1, 1), "w"), ("quant_conv.bias", (2 * L,), "b"),
          ("post_quant_conv.weight", (L, L, 1, 1), "w"), ("post_quant_conv.bias", (L,), "b"),
          ("decoder.conv_in.weight", (ch[-1], L, 3, 3), "w"), ("decoder.conv_in.bias", (ch[-1],), "b")]
    p += _mid("decoder.mid_block", ch[-1])
    rev = tuple(reversed(ch))
    cin = rev[0]
    for i, cout in enumerate(rev):
        for j in range(cfg.layers_per_block + 1):
            p += _resnet(f"decoder.up_blocks.{i}.resnets.{j}", cin, cout)
            cin = cout
        if i < len(ch) - 1:
            p += [(f"decoder.up_blocks.{i}.upsamplers.0.conv.weight", (cout, cout, 3, 3), "w"),
                  (f"decoder.up_blocks.{i}.upsamplers.0.conv.bias", (cout,), "b")]
    p += [("decoder.conv_norm_out.weight", (ch[0],), "gamma"), ("decoder.conv_norm_out.bias", (ch[0],), "beta"),
          ("decoder.conv_out.weight", (cfg.out_channels, ch[0], 3, 3), "w"),
          ("decoder.conv_out.bias", (cfg.out_channels,), "b")]
    return p


def num_params(cfg: VAEConfig = SD21_VAE) -> int:
    return sum(math.prod(s) for _, s, _ in param_table(cfg))


def init_state_dict(cfg: VAEConfig = SD21_VAE, seed: int = 0) -> Dict[str, torch.Tensor]:
    """Seeded variance-preserving init, same recipe as sd21.init_state_dict (CPU generator: identical everywhere)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}
    for name, shape, kind in param_table(cfg):
        if kind in ("w", "w_res"):
            std = (1.0 if kind == "w" else 0.5) / math.sqrt(math.prod(shape[1:]))
            sd[name] = torch.randn(shape, generator=g) * std
        elif kind == "gamma":
            sd[name] = 1.0 + 0.1 * torch.randn(shape, generator=g)
        else:
            sd[name] = 0.05 * torch.randn(shape, generator=g)
    return sd


_ATTN_RENAMES = ((".to_q.", ".query."), (".to_k.", ".key."), (".to_v.", ".value."), (".to_out.0.", ".proj_attn."))


def normalise_keys(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """Checkpoints saved by diffusers >= 0.18 name the mid-block attention `to_q / to_k / to_v / to_out.0`; the
    version range the reference runs on and the SD-2.1 hub files use `query / key / value / proj_attn`.
    Both load; missing or unexpected keys are an error."""
    out = {}
    for k, v in sd.items():
        if ".attentions." in k:
            for new, old in _ATTN_RENAMES:
                k = k.replace(new, old)
        out[k] = v
    return out


def check_state_dict(sd: Dict[str, torch.Tensor], cfg: VAEConfig) -> None:
    want = {n: s for n, s, _ in param_table(cfg)}
    missing = sorted(set(want) - set(sd))
    extra = sorted(set(sd) - set(want))
    bad = [k for k in want if k in sd and tuple(sd[k].shape) != want[k]
           and not (len(want[k]) == 4 and want[k][2:] == (1, 1) and tuple(sd[k].shape) == want[k][:2])]
    if missing or extra or bad:
        raise ops._abi.VNError(f"VAE state dict does not match the SD-2.1 layout: missing {missing[:3]} unexpected {extra[:3]} "
                               f"shape mismatch {bad[:3]}")


# ---- engine --------------------------------------------------------------------------------------
def _require_cuda(dev: torch.device) -> None:
    if dev.type != "cuda":
        raise ops._abi.VNError("VAEEngine needs a CUDA device: this path has no CPU fallback")


class _Res:
    pass


class VAEEngine:
    """Weights in kernel layout + launch sequences.  Activations are [nb, H*W, C] bf16; scratch buffers are keyed by
    role and shape and reused along the chain (forward only, one stream, stream order keeps reuse safe)."""

    # GroupNorm inputs up to this many elements take the one-launch form; above it the two-kernel form.  Measured on
    # B200 at 512 x 512: the one-launch form wins at every size of this network, even where a CTA's slab no longer
    # fits in registers and is re-read from L2 (encode 2.01 -> 1.91 ms, decode 3.52 -> 3.32 ms), so the default is "always"
    FUSED_GN_ELEMS = int(os.environ.get("VN_VAE_GN_FUSED_ELEMS", 1 << 40))
    # the four few-channel edge convolutions (3 -> C, C -> 8, 4 -> C, C -> 3) on vn_gemm: conv_in as im2col_thin + GEMM
    # with K padded to 64, conv_out as the implicit 3x3 GEMM with N padded to 8.  False: the CUDA-core thin-conv
    # kernels of the UNet (fp32 weights; 683 us for the decoder's 128 -> 3 at 512 x 512 against ~50 us on tensor cores)
    THIN_ON_GEMM = os.environ.get("VN_VAE_THIN_GEMM", "1") != "0"
    # the three stride-2 encoder convolutions without im2col (TMA element strides; 0: im2col + GEMM cross-check)
    S2_TMA = os.environ.get("VN_CONV_S2_TMA", "1") != "0"

    def __init__(self, state_dict: Dict[str, torch.Tensor], cfg: VAEConfig = SD21_VAE, device="cuda"):
        self.cfg = cfg
        self.dev = torch.device(device)
        _require_cuda(self.dev)
        self.ws = ops.Workspace(8192, 8192, self.dev)
        self._pools: "OrderedDict[tuple, Dict[tuple, torch.Tensor]]" = OrderedDict()   # call signature -> scratch
        self._bufs: Dict[tuple, torch.Tensor] = {}
        state_dict = normalise_keys(state_dict)
        check_state_dict(state_dict, cfg)
        self._prep(state_dict)
        n_gn = 2 * len(self.res) + 4
        self._stats = torch.zeros(n_gn, 64, cfg.norm_num_groups, 2, dtype=torch.float64, device=self.dev)
        self._parts = torch.empty(n_gn, ops.groupnorm_partial_floats(1), dtype=F32, device=self.dev)
        self._gn_i = 0

    # -- weights --
    def _prep(self, sd: Dict[str, torch.Tensor]) -> None:
        dev = self.dev
        f32 = lambda k: sd[k].detach().to(device=dev, dtype=F32).contiguous()          # noqa: E731
        conv = lambda k: sd[k].detach().to(dev, BF).permute(0, 2, 3, 1).reshape(sd[k].shape[0], -1).contiguous()  # noqa: E731
        lin = lambda w: w.detach().to(dev, BF).contiguous()                             # noqa: E731
        self.res: Dict[str, _Res] = {}
        self.samplers: Dict[str, Tuple[torch.Tensor, torch.Tensor]] = {}
        self.attn: Dict[str, SimpleNamespace] = {}
        for key in sd:
            if key.endswith(".conv1.weight"):
                p = key[: -len(".conv1.weight")]
                r = _Res()
                r.cin, r.cout = sd[key].shape[1], sd[key].shape[0]
                r.n1 = (f32(p + ".norm1.weight"), f32(p + ".norm1.bias"))
                r.n2 = (f32(p + ".norm2.weight"), f32(p + ".norm2.bias"))
                r.c1, r.c1b = conv(key), f32(p + ".conv1.bias")
                r.c2, r.c2b = conv(p + ".conv2.weight"), f32(p + ".conv2.bias")
                r.sc = None
                if p + ".conv_shortcut.weight" in sd:
                    w = sd[p + ".conv_shortcut.weight"]
                    r.sc, r.scb = lin(w.reshape(w.shape[0], w.shape[1])), f32(p + ".conv_shortcut.bias")
                self.res[p] = r
            elif key.endswith("samplers.0.conv.weight"):
                p = key[: -len(".conv.weight")]
                self.samplers[p] = (conv(key), f32(p + ".conv.bias"))
            elif key.endswith(".group_norm.weight"):
                p = key[: -len(".group_norm.weight")]
                a = SimpleNamespace()
                a.c = sd[key].shape[0]
                a.gn = (f32(p + ".group_norm.weight"), f32(p + ".group_norm.bias"))
                a.qk = lin(torch.cat([sd[p + ".query.weight"], sd[p + ".key.weight"]], 0))
                a.qkb = torch.cat([f32(p + ".query.bias"), f32(p + ".key.bias")]).contiguous()
                a.v, a.vb = lin(sd[p + ".value.weight"]), f32(p + ".value.bias")
                a.o, a.ob = lin(sd[p + ".proj_attn.weight"]), f32(p + ".proj_attn.bias")
                self.attn[p] = a
        self.enc_in = (f32("encoder.conv_in.weight"), f32("encoder.conv_in.bias"))
        self.enc_norm = (f32("encoder.conv_norm_out.weight"), f32("encoder.conv_norm_out.bias"))
        # quant_conv (1x1, 8 -> 8) directly follows encoder.conv_out (3x3, C -> 8) with nothing in between: one conv
        # with W' = Wq . Wout, b' = Wq . bout + bq  (composed in fp64, stored fp32)
        wq = sd["quant_conv.weight"].detach().double().flatten(1)
        wo, bo = sd["encoder.conv_out.weight"].detach().double(), sd["encoder.conv_out.bias"].detach().double()
        self.enc_out = (torch.einsum("ab,bcij->acij", wq, wo).to(dev, F32).contiguous(),
                        (wq @ bo + sd["quant_conv.bias"].detach().double()).to(dev, F32).contiguous())
        # post_quant_conv precedes a zero-PADDED conv, so its bias cannot be folded; it stays a 4 x 4 channel mix
        self.post_quant = (f32("post_quant_conv.weight").flatten(1), f32("post_quant_conv.bias"))
        self.dec_in = (f32("decoder.conv_in.weight"), f32("decoder.conv_in.bias"))
        self.dec_norm = (f32("decoder.conv_norm_out.weight"), f32("decoder.conv_norm_out.bias"))
        self.dec_out = (f32("decoder.conv_out.weight"), f32("decoder.conv_out.bias"))
        # GEMM forms of the edge convolutions
        def thin_in(w, b):          # [Cout, Ct, 3, 3] -> [Cout, Kpad], k = tap*Ct + ct (vn_im2col_thin)
            k = 9 * w.shape[1]
            wk = torch.zeros(w.shape[0], (k + 63) // 64 * 64, dtype=BF, device=dev)
            wk[:, :k] = w.permute(0, 2, 3, 1).reshape(w.shape[0], k).to(BF)
            return wk, b

        def thin_out(w, b):         # [Ct, C, 3, 3] -> [8, 9*C] (rows >= Ct zero), k = tap*C + c (implicit 3x3 GEMM)
            wk = torch.zeros(8, 9 * w.shape[1], dtype=BF, device=dev)
            wk[: w.shape[0]] = w.permute(0, 2, 3, 1).reshape(w.shape[0], -1).to(BF)
            bk = torch.zeros(8, dtype=F32, device=dev)
            bk[: w.shape[0]] = b
            return wk, bk

        self.enc_in_g, self.dec_in_g = thin_in(*self.enc_in), thin_in(*self.dec_in)
        self.enc_out_g, self.dec_out_g = thin_out(*self.enc_out), thin_out(*self.dec_out)

    def weight_bytes(self) -> int:
        seen, n = set(), 0
        for obj in list(self.res.values()) + list(self.attn.values()):
            for t in vars(obj).values():
                for u in (t if isinstance(t, tuple) else (t,)):
                    if torch.is_tensor(u) and u.data_ptr() not in seen:
                        seen.add(u.data_ptr())
                        n += u.numel() * u.element_size()
        for w, b in self.samplers.values():
            n += w.numel() * 2 + b.numel() * 4
        return n

    # -- buffers --
    def _buf(self, role: str, shape, dtype=BF) -> torch.Tensor:
        key = (role, tuple(shape), dtype)
        t = self._bufs.get(key)
        if t is None:
            t = self._bufs[key] = torch.empty(tuple(shape), dtype=dtype, device=self.dev)
        return t

    MAX_POOLS = 4          # scratch is kept for the last few (map, batch, height, width) signatures only

    def _begin(self, nb: int, sig: tuple = ()) -> None:
        if nb > self._stats.shape[1]:
            raise ops._abi.VNError(f"VAEEngine: batch {nb} > {self._stats.shape[1]}")
        pool = self._pools.pop(sig, None)
        self._pools[sig] = self._bufs = pool if pool is not None else {}
        while len(self._pools) > self.MAX_POOLS:       # stream-ordered allocator: freeing under in-flight kernels is safe
            self._pools.popitem(last=False)
        self._stats.zero_()
        ops.memset(self._parts, 0xFF)
        self._gn_i = 0

    # -- pieces --
    def _gn(self, x: torch.Tensor, gb, silu: bool, role: str = "gn") -> torch.Tensor:
        nb, hw, C = x.shape
        i = self._gn_i
        self._gn_i += 1
        y = self._buf(role, x.shape)
        fused = nb * hw * C <= self.FUSED_GN_ELEMS
        ops.groupnorm_fwd(x, gb[0], gb[1], self.cfg.norm_eps, silu, y, nb, hw, self.cfg.norm_num_groups,
                          self._stats[i, :nb], self._parts[i] if fused else None)
        return y

    def _conv_in(self, x_in: torch.Tensor, w32, wg, out: torch.Tensor, H: int, W: int) -> None:
        """x_in NCHW fp32 [nb,Ct,H,W] -> out [nb, H*W, C] bf16."""
        nb, C = out.shape[0], out.shape[-1]
        if self.THIN_ON_GEMM:
            col = self._buf("thin.col", (nb * H * W, wg[0].shape[1]))
            ops.im2col_thin(x_in, col)
            ops.gemm(col, wg[0], out.view(nb * H * W, C), bias=wg[1], ws=self.ws)
        else:
            ops.conv_in_fwd(x_in, w32[0], w32[1], out.view(nb, H, W, C))

    def _conv_out(self, a: torch.Tensor, w32, wg, H: int, W: int) -> torch.Tensor:
        """a [nb, H*W, C] bf16 -> fresh NCHW fp32 [nb, Ct, H, W]."""
        nb, C, Ct = a.shape[0], a.shape[-1], w32[0].shape[0]
        if self.THIN_ON_GEMM:
            d = self._buf("thin.out", (nb, H, W, 8), F32)
            ops.conv3x3(a.view(nb, H, W, C), wg[0], d, bias=wg[1], ws=self.ws, force_bn=64, force_split=1)
            return d[..., :Ct].permute(0, 3, 1, 2).contiguous()
        y = torch.empty(nb, Ct, H, W, dtype=F32, device=self.dev)
        ops.conv_out_fwd(a.view(nb, H, W, C), w32[0], w32[1], y)
        return y

    def _resnet(self, name: str, x: torch.Tensor, H: int, W: int, out_role: str) -> torch.Tensor:
        """ResnetBlock2D without time embedding (oracle/vae.py:resnet)."""
        r = self.res[name]
        nb, hw = x.shape[0], H * W
        a1 = self._gn(x, r.n1, True)
        h1 = self._buf("h1", (nb, hw, r.cout))
        ops.conv3x3(a1.view(nb, H, W, r.cin), r.c1, h1.view(nb, H, W, r.cout), bias=r.c1b, ws=self.ws)
        a2 = self._gn(h1, r.n2, True)
        if r.sc is not None:
            sc = self._buf("sc", (nb, hw, r.cout))
            ops.gemm(x.view(nb * hw, r.cin), r.sc, sc.view(nb * hw, r.cout), bias=r.scb, ws=self.ws)
        else:
            sc = x
        out = self._buf(out_role, (nb, hw, r.cout))
        ops.conv3x3(a2.view(nb, H, W, r.cout), r.c2, out.view(nb, H, W, r.cout), bias=r.c2b,
                    R=sc.view(nb, H, W, r.cout), ws=self.ws)
        return out

    def _attention(self, name: str, x: torch.Tensor, out_role: str) -> torch.Tensor:
        """AttentionBlock, one head over all C channels (oracle/vae.py:attention)."""
        a = self.attn[name]
        nb, hw, C = x.shape
        if hw % 64 != 0:
            raise ops._abi.VNError(f"VAE attention: {hw} latent pixels, need a multiple of 64")
        t = self._gn(x, a.gn, False)
        qk = self._buf("attn.qk", (nb, hw, 2 * C))
        ops.gemm(t.view(nb * hw, C), a.qk, qk.view(nb * hw, 2 * C), bias=a.qkb, ws=self.ws)
        vt = self._buf("attn.vt", (C, hw))
        S = self._buf("attn.s", (hw, hw), F32)
        P = self._buf("attn.p", (hw, hw))
        o = self._buf("attn.o", (nb, hw, C))
        for b in range(nb):
            # V^T = Wv X^T straight out of the GEMM (the P V product wants V as its [N, K] operand); the value bias is
            # added after P V instead - softmax rows sum to one.  These three products read a B operand that an
            # earlier launch wrote, which vn_gemm only allows without programmatic dependent launch (viewneti.h).
            ops.gemm(a.v, t[b], vt, ws=self.ws, b_dynamic=True)
            ops.gemm(qk[b, :, :C], qk[b, :, C:], S, ws=self.ws, b_dynamic=True)
            ops.softmax_rows(S, P, 1.0 / math.sqrt(C))
            ops.gemm(P, vt, o[b], bias=a.vb, ws=self.ws, b_dynamic=True)
        out = self._buf(out_role, (nb, hw, C))
        ops.gemm(o.view(nb * hw, C), a.o, out.view(nb * hw, C), bias=a.ob, R=x.view(nb * hw, C), ws=self.ws)
        return out

    def _mid(self, p: str, x: torch.Tensor, H: int, W: int) -> torch.Tensor:
        x = self._resnet(p + ".resnets.0", x, H, W, "m1")          # roles of its own: never aliases the caller's x0 / x1
        x = self._attention(p + ".attentions.0", x, "m0")
        return self._resnet(p + ".resnets.1", x, H, W, "m1")

    # -- the two maps --
    @torch.no_grad()
    def encode_moments(self, pixel_values: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """[nb,3,H,W] (any float dtype, values in [-1,1]) -> (mean, logvar) fp32 [nb,4,H/8,W/8], logvar clamped to
        [-30, 20] (DiagonalGaussianDistribution).  The returned tensors are fresh; scratch is reused by the next call."""
        cfg = self.cfg
        nb, cin, H, W = pixel_values.shape
        if cin != cfg.in_channels or H % cfg.downscale or W % cfg.downscale:
            raise ops._abi.VNError(f"VAE encode: bad input shape {tuple(pixel_values.shape)}")
        x_in = pixel_values.detach().to(device=self.dev, dtype=F32).contiguous()
        self._begin(nb, ("encode", nb, H, W))
        ch = cfg.block_out_channels
        x = self._buf("x0", (nb, H * W, ch[0]))
        self._conv_in(x_in, self.enc_in, self.enc_in_g, x, H, W)
        flip = 1
        for i in range(len(ch)):
            for j in range(cfg.layers_per_block):
                x = self._resnet(f"encoder.down_blocks.{i}.resnets.{j}", x, H, W, f"x{flip}")
                flip ^= 1
            if i < len(ch) - 1:
                wf, bias = self.samplers[f"encoder.down_blocks.{i}.downsamplers.0"]
                Ho, Wo = H // 2, W // 2
                y = self._buf(f"x{flip}", (nb, Ho * Wo, ch[i]))
                flip ^= 1
                if self.S2_TMA:
                    ops.conv3x3(x.view(nb, H, W, ch[i]), wf, y.view(nb, Ho, Wo, ch[i]), bias=bias, ws=self.ws, stride=2, pad=0)
                else:
                    col = self._buf("col", (nb * Ho * Wo, 9 * ch[i]))
                    ops.im2col_s2_pad0(x.view(nb, H, W, ch[i]), col)
                    ops.gemm(col, wf, y.view(nb * Ho * Wo, ch[i]), bias=bias, ws=self.ws)
                x = y
                H, W = Ho, Wo
        x = self._mid("encoder.mid_block", x, H, W)
        a = self._gn(x, self.enc_norm, True)
        m = self._conv_out(a, self.enc_out, self.enc_out_g, H, W)
        mean, logvar = m.chunk(2, dim=1)
        return mean, logvar.clamp(-30.0, 20.0)

    @torch.no_grad()
    def decode(self, z: torch.Tensor) -> torch.Tensor:
        """latents already divided by the scaling factor [nb,4,h,w] -> image fp32 [nb,3,8h,8w]."""
        cfg = self.cfg
        nb, L, H, W = z.shape
        if L != cfg.latent_channels:
            raise ops._abi.VNError(f"VAE decode: bad latent shape {tuple(z.shape)}")
        z = z.detach().to(device=self.dev, dtype=F32)
        # post_quant_conv: a 4 x 4 channel mix of 4 h w numbers (see _prep for why it is not folded into conv_in)
        z = (torch.einsum("ab,nbhw->nahw", self.post_quant[0], z) + self.post_quant[1].view(1, -1, 1, 1)).contiguous()
        self._begin(nb, ("decode", nb, H, W))
        ch = cfg.block_out_channels
        rev = tuple(reversed(ch))
        x = self._buf("x0", (nb, H * W, rev[0]))
        self._conv_in(z, self.dec_in, self.dec_in_g, x, H, W)
        x = self._mid("decoder.mid_block", x, H, W)
        flip = 0
        for i, cout in enumerate(rev):
            for j in range(cfg.layers_per_block + 1):
                x = self._resnet(f"decoder.up_blocks.{i}.resnets.{j}", x, H, W, f"x{flip}")
                flip ^= 1
            if i < len(ch) - 1:
                wf, bias = self.samplers[f"decoder.up_blocks.{i}.upsamplers.0"]
                up = self._buf("up", (nb, 4 * H * W, cout))
                ops.upsample2x_fwd(x.view(nb, H, W, cout), up.view(nb, 2 * H, 2 * W, cout))
                H, W = 2 * H, 2 * W
                x = self._buf(f"x{flip}", (nb, H * W, cout))
                flip ^= 1
                ops.conv3x3(up.view(nb, H, W, cout), wf, x.view(nb, H, W, cout), bias=bias, ws=self.ws)
        a = self._gn(x, self.dec_norm, True)
        return self._conv_out(a, self.dec_out, self.dec_out_g, H, W)

    def scratch_bytes(self) -> int:
        return sum(t.numel() * t.element_size() for pool in self._pools.values() for t in pool.values())


# ---- drop-in surface ---------------------------------------------------------------------------------
class DiagonalGaussianDistribution:
    """What `vae.encode(x).latent_dist` is in diffusers: `.sample(generator=None)`, `.mode()`, `.mean`, `.logvar`, `.std`."""

    def __init__(self, mean: torch.Tensor, logvar: torch.Tensor):
        self.mean, self.logvar = mean, logvar
        self.std = torch.exp(0.5 * logvar)
        self.var = torch.exp(logvar)

    def sample(self, generator: Optional[torch.Generator] = None) -> torch.Tensor:
        # diffusers' randn_tensor: draw on the generator's device (a CPU generator is legal), then move
        gdev = generator.device if generator is not None else self.mean.device
        noise = torch.randn(self.mean.shape, generator=generator, device=gdev, dtype=self.mean.dtype).to(self.mean.device)
        return self.mean + self.std * noise

    def mode(self) -> torch.Tensor:
        return self.mean


class AutoencoderKL(torch.nn.Module):
    """The object reference training/coach.py:628-633 gets from `AutoencoderKL.from_pretrained(..., subfolder="vae")`,
    as far as the reference touches it: `.encode(x).latent_dist.sample()`, `.decode(z).sample`, `.config.scaling_factor`,
    `.requires_grad_(False)`, `.to(device, dtype=...)`, `.eval()`.  Frozen: neither map records autograd history
    (the reference detaches the encoder output and decodes under no_grad)."""

    def __init__(self, state_dict: Optional[Dict[str, torch.Tensor]] = None, cfg: VAEConfig = SD21_VAE, device="cuda"):
        super().__init__()
        self.config = SimpleNamespace(**{f: getattr(cfg, f) for f in cfg.__dataclass_fields__})
        self.cfg = cfg
        sd = state_dict if state_dict is not None else init_state_dict(cfg, 0)
        self.engine = VAEEngine(sd, cfg, device)
        self.dtype = F32

    @classmethod
    def from_pretrained(cls, path: str, subfolder: Optional[str] = None, revision: Optional[str] = None, device="cuda",
                        cfg: VAEConfig = SD21_VAE, **_) -> "AutoencoderKL":
        """`AutoencoderKL.from_pretrained(pretrained_model_name_or_path, subfolder="vae", revision=...)` (reference
        training/coach.py:628-633) for a LOCAL diffusers directory (no network here)."""
        root = os.path.join(path, subfolder) if subfolder else path
        for fn in ("diffusion_pytorch_model.safetensors", "diffusion_pytorch_model.bin", "diffusion_pytorch_model.pt"):
            p = os.path.join(root, fn)
            if not os.path.exists(p):
                continue
            if fn.endswith(".safetensors"):
                from safetensors.torch import load_file
                return cls(load_file(p), cfg, device)
            return cls(torch.load(p, map_location="cpu"), cfg, device)
        raise FileNotFoundError(f"no VAE checkpoint under {root} (AutoencoderKL(None) builds seeded synthetic weights)")

    def to(self, *args, **kwargs):           # weights live in kernel layout on the engine's device; dtype is advisory
        for a in list(args) + list(kwargs.values()):
            if isinstance(a, torch.dtype):
                self.dtype = a
        return self

    def encode(self, x: torch.Tensor, return_dict: bool = True):
        mean, logvar = self.engine.encode_moments(x)
        dist = DiagonalGaussianDistribution(mean.to(self.dtype), logvar.to(self.dtype))
        return SimpleNamespace(latent_dist=dist) if return_dict else (dist,)

    def decode(self, z: torch.Tensor, return_dict: bool = True):
        img = self.engine.decode(z).to(self.dtype)
        return SimpleNamespace(sample=img) if return_dict else (img,)


def decode_latents(vae: AutoencoderKL, latents: torch.Tensor):
    """StableDiffusionPipeline.decode_latents as reached from reference sd_pipeline_call.py:115: numpy NHWC float32
    image in [0, 1]."""
    img = vae.decode(latents / vae.config.scaling_factor).sample
    return (img / 2 + 0.5).clamp(0, 1).cpu().permute(0, 2, 3, 1).float().numpy()
