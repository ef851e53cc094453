"""Static execution plan of the frozen SD-2.1 UNet on the sm_100a library: forward and the dgrad-only backward
that stops at the 32 XTI context tensors.

Replaces, for the path reference training/coach.py:197-214 / sd_pipeline_call.py:78-94,
  * diffusers `UNet2DConditionModel.forward`  (ResnetBlock2D, Transformer2DModel, Down/Upsample2D)
  * reference models/xti_attention_processor.py:9-57 for all 32 attention modules (K from CONTEXT_TENSOR_i, V from
    CONTEXT_TENSOR_BYPASS_i, layer index bound statically instead of the mutable `this_idx` counter)
  * torch autograd's generic backward: weights are frozen (coach.py:647-648), so only data gradients are computed,
    no weight gradients, and nothing upstream of the first cross-attention's K/V is visited.

Host side is Python; every arithmetic op is one C-ABI call (view_neti_b200.ops).  All buffers are allocated
once per (batch, h, w) plan, so a whole step is CUDA-graph capturable (`capture_*`).  Layout: activations
bf16 channel-last [nb, h*w, C]; skip tensors are written straight into their slice of the up-block concat
buffer (torch.cat never runs); norm / bias / time-embedding parameters fp32.
"""
from __future__ import annotations

import os
from typing import Dict, List, Tuple

import torch

from . import ops
from .sd21 import SD21, UNetConfig, up_block_resnet_channels

BF = torch.bfloat16
F32 = torch.float32
_S2_TMA = os.environ.get("VN_CONV_S2_TMA", "1") != "0"      # forward stride-2 convs without im2col (0: im2col + GEMM cross-check)


def _lin(w: torch.Tensor, dev) -> Tuple[torch.Tensor, torch.Tensor]:
    """Linear weight [out, in] -> (forward B operand [out, in], dgrad B operand [in, out]) in bf16."""
    wb = w.to(device=dev, dtype=BF)
    return wb.contiguous(), wb.t().contiguous()


def _conv(w: torch.Tensor, dev) -> Tuple[torch.Tensor, torch.Tensor]:
    """Conv weight [out, in, 3, 3] -> implicit-GEMM operands: fwd [out, 9*in] (k = tap*in + c) and
    dgrad [in, 9*out] with the taps flipped."""
    wb = w.to(device=dev, dtype=BF)
    f = wb.permute(0, 2, 3, 1).reshape(wb.shape[0], -1).contiguous()
    b = wb.flip(2, 3).permute(1, 2, 3, 0).reshape(wb.shape[1], -1).contiguous()
    return f, b


class _Res:
    pass


class _Xf:
    pass


class UNetEngine:
    """Weights + one static plan per input shape.  `forward` / `backward` are plain launch sequences."""

    def __init__(self, state_dict: Dict[str, torch.Tensor], cfg: UNetConfig = SD21, device="cuda"):
        self.cfg = cfg
        self.dev = torch.device(device)
        if self.dev.type != "cuda":
            raise ops._abi.VNError("UNetEngine needs a CUDA device: the hot path has no CPU fallback")
        self._plans: Dict[Tuple[int, int, int], "_Plan"] = {}
        self.use_graphs = True      # drop-in API: replay CUDA graphs from the second step on
        self._prep_weights(state_dict)

    # ------------------------------------------------------------------------------------------------
    def _prep_weights(self, sd: Dict[str, torch.Tensor]) -> None:
        cfg, dev = self.cfg, self.dev
        f32 = lambda k: sd[k].to(device=dev, dtype=F32).contiguous()   # noqa: E731
        self.conv_in_w, self.conv_in_b = f32("conv_in.weight"), f32("conv_in.bias")
        self.conv_out_w, self.conv_out_b = f32("conv_out.weight"), f32("conv_out.bias")
        self.norm_out = (f32("conv_norm_out.weight"), f32("conv_norm_out.bias"))
        # The three few-channel edge convolutions on the tensor-core GEMM (as in models/vae.py): conv_in (4 -> C) and the
        # dgrad of conv_out (4 -> C) as vn_im2col_thin (K = 9*4 zero-padded to one 64-wide k-block) + GEMM; conv_out
        # (C -> 4) as the implicit 3x3 GEMM with N padded to 8 and an fp32 destination.  The CUDA-core thin-conv kernels
        # they replace cost 39 / 26 / 72 us of a 7.6 ms step at 64x64 (profiles/r2_graph_timeline_v0.txt); VN_UNET_THIN_GEMM=0
        # keeps them as a cross-check.
        import os
        self.thin_gemm = os.environ.get("VN_UNET_THIN_GEMM", "1") != "0"

        def thin_in(w):             # [Cout, Ct, 3, 3] -> [Cout, 64], k = tap*Ct + ct
            k = 9 * w.shape[1]
            wk = torch.zeros(w.shape[0], (k + 63) // 64 * 64, dtype=BF, device=dev)
            wk[:, :k] = w.permute(0, 2, 3, 1).reshape(w.shape[0], k).to(device=dev, dtype=BF)
            return wk

        w_in, w_out = sd["conv_in.weight"], sd["conv_out.weight"]
        self.conv_in_g = thin_in(w_in)
        self.conv_out_g = torch.zeros(8, 9 * w_out.shape[1], dtype=BF, device=dev)            # [8, 9*C], rows >= 4 zero
        self.conv_out_g[: w_out.shape[0]] = w_out.permute(0, 2, 3, 1).reshape(w_out.shape[0], -1).to(device=dev, dtype=BF)
        self.conv_out_gb = torch.zeros(8, dtype=F32, device=dev)
        self.conv_out_gb[: w_out.shape[0]] = sd["conv_out.bias"].to(dev, F32)
        # dgrad of conv_out = a 4 -> C convolution with the taps flipped and in / out channels swapped
        self.conv_out_bwd_g = thin_in(w_out.flip(2, 3).permute(1, 0, 2, 3))
        self.te1 = (sd["time_embedding.linear_1.weight"].to(dev, BF).contiguous(), f32("time_embedding.linear_1.bias"))
        self.te2 = (sd["time_embedding.linear_2.weight"].to(dev, BF).contiguous(), f32("time_embedding.linear_2.bias"))
        self.res: Dict[str, _Res] = {}
        self.xf: Dict[str, _Xf] = {}
        self.down: Dict[str, tuple] = {}
        self.up: Dict[str, tuple] = {}
        temb_w, temb_b, off = [], [], 0
        for key in sd:
            if key.endswith(".conv1.weight"):
                p = key[: -len(".conv1.weight")]
                r = _Res()
                r.cin, r.cout = sd[key].shape[1], sd[key].shape[0]
                r.n1 = (f32(p + ".norm1.weight"), f32(p + ".norm1.bias"))
                r.n2 = (f32(p + ".norm2.weight"), f32(p + ".norm2.bias"))
                r.c1f, r.c1b = _conv(sd[key], dev)
                r.c1bias = f32(p + ".conv1.bias")
                r.c2f, r.c2b = _conv(sd[p + ".conv2.weight"], dev)
                r.c2bias = f32(p + ".conv2.bias")
                r.temb_off = off
                off += r.cout
                temb_w.append(sd[p + ".time_emb_proj.weight"].to(dev, BF))
                temb_b.append(f32(p + ".time_emb_proj.bias"))
                if p + ".conv_shortcut.weight" in sd:
                    w = sd[p + ".conv_shortcut.weight"]
                    r.scf, r.scb = _lin(w.reshape(w.shape[0], w.shape[1]), dev)
                    r.scbias = f32(p + ".conv_shortcut.bias")
                else:
                    r.scf = None
                self.res[p] = r
            elif key.endswith(".proj_in.weight"):
                p = key[: -len(".proj_in.weight")]
                b = p + ".transformer_blocks.0"
                x = _Xf()
                x.c = sd[key].shape[0]
                x.gn = (f32(p + ".norm.weight"), f32(p + ".norm.bias"))
                x.pif, x.pib = _lin(sd[key], dev)
                x.pibias = f32(p + ".proj_in.bias")
                x.pof, x.pob = _lin(sd[p + ".proj_out.weight"], dev)
                x.pobias = f32(p + ".proj_out.bias")
                x.ln = [(f32(f"{b}.norm{i}.weight"), f32(f"{b}.norm{i}.bias")) for i in (1, 2, 3)]
                qkv = torch.cat([sd[f"{b}.attn1.to_q.weight"], sd[f"{b}.attn1.to_k.weight"], sd[f"{b}.attn1.to_v.weight"]], 0)
                x.qkvf, x.qkvb = _lin(qkv, dev)
                x.o1f, x.o1b = _lin(sd[f"{b}.attn1.to_out.0.weight"], dev)
                x.o1bias = f32(f"{b}.attn1.to_out.0.bias")
                x.q2f, x.q2b = _lin(sd[f"{b}.attn2.to_q.weight"], dev)
                x.k2f, x.k2b = _lin(sd[f"{b}.attn2.to_k.weight"], dev)
                x.v2f, x.v2b = _lin(sd[f"{b}.attn2.to_v.weight"], dev)
                # K and V projections of a layer as ONE product: A = [ctx_k ; ctx_v] (rows), B = [Wk ; Wv] -> D [2*rows, 2C] whose
                # diagonal blocks are K and V (the off-diagonal blocks are never read); same for the two context gradients.
                # Halves the 64 tiny (77-row) launches of a step for 2x of their negligible FLOPs.
                x.kv2f = torch.cat([x.k2f, x.v2f], 0).contiguous()          # [2C, 1024]
                x.kv2b = torch.cat([x.k2b, x.v2b], 0).contiguous()          # [2 * 1024, C]
                x.o2f, x.o2b = _lin(sd[f"{b}.attn2.to_out.0.weight"], dev)
                x.o2bias = f32(f"{b}.attn2.to_out.0.bias")
                x.ff1f, x.ff1b = _lin(sd[f"{b}.ff.net.0.proj.weight"], dev)
                x.ff1bias = f32(f"{b}.ff.net.0.proj.bias")
                x.ff2f, x.ff2b = _lin(sd[f"{b}.ff.net.2.weight"], dev)
                x.ff2bias = f32(f"{b}.ff.net.2.bias")
                self.xf[p] = x
            elif key.endswith(".downsamplers.0.conv.weight"):
                p = key[: -len(".conv.weight")]
                w = sd[key].to(dev, BF)
                wf = w.permute(0, 2, 3, 1).reshape(w.shape[0], -1).contiguous()          # [out, 9*in]
                self.down[p] = (wf, wf.t().contiguous(), f32(p + ".conv.bias"))          # dgrad: dcol = dy @ wf
            elif key.endswith(".upsamplers.0.conv.weight"):
                p = key[: -len(".conv.weight")]
                f, b = _conv(sd[key], dev)
                self.up[p] = (f, b, f32(p + ".conv.bias"))
        self.temb_w = torch.cat(temb_w, 0).contiguous()
        self.temb_b = torch.cat(temb_b, 0).contiguous()
        self.temb_total = off

    # ------------------------------------------------------------------------------------------------
    def plan(self, nb: int, h: int, w: int) -> "_Plan":
        key = (nb, h, w)
        if key not in self._plans:
            self._plans[key] = _Plan(self, nb, h, w)
        return self._plans[key]

    def weight_bytes(self) -> int:
        tot = 0
        seen = set()
        for obj in list(self.res.values()) + list(self.xf.values()):
            for v in vars(obj).values():
                for t in (v if isinstance(v, (tuple, list)) else (v,)):
                    for u in (t if isinstance(t, (tuple, list)) else (t,)):
                        if torch.is_tensor(u) and u.data_ptr() not in seen:
                            seen.add(u.data_ptr())
                            tot += u.numel() * u.element_size()
        return tot


class _Plan:
    """All buffers and the op order for one (nb, h, w)."""

    def __init__(self, eng: UNetEngine, nb: int, h: int, w: int):
        cfg = eng.cfg
        nlev = len(cfg.block_out_channels)
        assert h % (1 << (nlev - 1)) == 0 and w % (1 << (nlev - 1)) == 0, "latent size must be divisible by 8"
        self.eng, self.nb, self.h, self.w = eng, nb, h, w
        self.dev = eng.dev
        self.bufs: Dict[str, torch.Tensor] = {}
        self.L = cfg.context_len
        self.n_layers = cfg.num_cross_layers
        ch0 = cfg.block_out_channels[0]
        self.ws = ops.Workspace(nb * h * w, 8 * ch0, self.dev,
                                dkv_elems=2 * nb * self.L * max(cfg.block_out_channels))
        # static inputs / outputs
        self.latents = self.buf("in.latents", (nb, cfg.in_channels, h, w), F32)
        self.timesteps = torch.zeros(nb, dtype=torch.int64, device=self.dev)
        # contexts: stored [layer][k|v][nb][77][1024] so that a layer's K and V sources are adjacent rows of one GEMM operand;
        # `ctx` is the [k|v][layer] view every caller indexes (ctx[0, l] = CONTEXT_TENSOR_l, ctx[1, l] = ..._BYPASS_l)
        self.ctx_store = self.buf("in.ctx", (self.n_layers, 2, nb, self.L, cfg.cross_attention_dim), F32)
        self.ctx = self.ctx_store.permute(1, 0, 2, 3, 4)
        self.eps = self.buf("out.eps", (nb, cfg.out_channels, h, w), F32)
        self.d_eps = self.buf("in.d_eps", (nb, cfg.out_channels, h, w), F32)
        # context gradients: one [2*nb*77, 2*1024] product per layer, dK-context / dV-context are its diagonal blocks;
        # `d_ctx` is the [k|v][layer][nb][77][1024] strided view of those blocks
        D = cfg.cross_attention_dim
        R = nb * self.L
        self.d_ctx_store = self.buf("out.d_ctx", (self.n_layers, 2 * R, 2 * D), F32)
        self.d_ctx = self.d_ctx_store.as_strided((2, self.n_layers, nb, self.L, D),
                                                 (R * 2 * D + D, 2 * R * 2 * D, self.L * 2 * D, 2 * D, 1))
        self.target = self.buf("in.target", (nb, cfg.out_channels, h, w), F32)
        self.loss = self.buf("out.loss", (1,), F32)
        n_gn = 2 * (len(eng.res)) + len(eng.xf) + 1
        self.stat_f = self.buf("gn.stats", (n_gn, nb, cfg.norm_num_groups, 2), torch.float64)     # (sum x, sum x^2)
        self.stat_b = self.buf("gn.red", (n_gn, nb, cfg.norm_num_groups, 2), torch.float64)       # backward reductions
        # per-CTA partial sums through which the CTAs of a fused GroupNorm exchange statistics: one private slot per
        # GroupNorm, all bytes preset to 0xff by ONE memset per pass (a written word is its own arrival flag)
        npf = ops.groupnorm_partial_floats(nb)
        self.part_f = self.buf("gn.part_f", (n_gn, npf), torch.float32)
        self.part_b = self.buf("gn.part_b", (n_gn, npf), torch.float32)
        self._stat_slots: Dict[str, int] = {}
        # side stream: the 32 context projections (forward) and the 32 context-gradient GEMMs (backward) do not sit on
        # the UNet's dependency chain; they run concurrently with it and are joined by events (captured in the graph)
        self.side = torch.cuda.Stream(device=self.dev)
        # scratch of the self-attention forward's work balancing (largest request over the attention levels of this plan)
        need = max([max(ops.attention_fwd_workspace_bytes(nb, cfg.block_out_channels[i] // 64, (h >> i) * (w >> i), (h >> i) * (w >> i)),
                        ops.attention_bwd_workspace_bytes(nb, cfg.block_out_channels[i] // 64, (h >> i) * (w >> i), (h >> i) * (w >> i)))
                    for i in range(nlev)] + [0])
        self.attn_ws = torch.empty(need, dtype=torch.uint8, device=self.dev) if need > 0 else None
        self.ev_temb = torch.cuda.Event()
        self.ev_kv = [torch.cuda.Event() for _ in range(self.n_layers)]
        self.ev_dkv = [torch.cuda.Event() for _ in range(self.n_layers)]
        self.ev_fin = [torch.cuda.Event() for _ in range(self.n_layers)]   # dK / dV finished on the side stream (accumulator free)
        self._fin_ev = None
        self.graphs: Dict[str, torch.cuda.CUDAGraph] = {}
        self.launches: Dict[str, int] = {}
        self._saved = False
        self.trace_f: Dict[str, torch.Tensor] = {}     # block name -> output view   (parity debugging)
        self.trace_b: Dict[str, torch.Tensor] = {}     # block name -> d(output) view

    # ---- buffers ------------------------------------------------------------------------------------
    def buf(self, name: str, shape, dtype=BF) -> torch.Tensor:
        t = self.bufs.get(name)
        if t is None:
            t = torch.zeros(tuple(shape), dtype=dtype, device=self.dev)
            self.bufs[name] = t
        assert tuple(t.shape) == tuple(shape) and t.dtype == dtype, name
        return t

    def _stat(self, name: str, arena: torch.Tensor) -> torch.Tensor:
        """[nb, groups, 2] fp32 GroupNorm statistics slot; each arena is zeroed by ONE memset per pass."""
        i = self._stat_slots.get(name)
        if i is None:
            i = self._stat_slots[name] = len([k for k in self._stat_slots if k.endswith(name[-3:])])
            assert i < arena.shape[0], "statistics arena too small"
        return arena[i]

    def _fork(self) -> torch.cuda.Event:
        """Make the side stream wait for everything enqueued so far on the current stream; returns a fresh event
        the caller records on the side stream and later waits on (fork / join, capturable)."""
        e = torch.cuda.Event()
        e.record(torch.cuda.current_stream())
        self.side.wait_event(e)
        return torch.cuda.Event()

    def _xf_names(self) -> List[str]:
        from .sd21 import cross_attn_layer_names
        return cross_attn_layer_names(self.eng.cfg)

    def act_bytes(self) -> int:
        return sum(t.numel() * t.element_size() for t in self.bufs.values())

    # ---- forward pieces -----------------------------------------------------------------------------
    def _gn(self, name, x, gb, eps, silu, hw):
        """GroupNorm(+SiLU) of a [nb, hw, C] view; stats are kept for the backward."""
        G = self.eng.cfg.norm_num_groups
        st = self._stat(name + ".st", self.stat_f)
        y = self.buf(name + ".y", (self.nb, hw, x.shape[-1]))
        ops.groupnorm_fwd(x, gb[0], gb[1], eps, silu, y, self.nb, hw, G, st, self.part_f[self._stat_slots[name + ".st"]])
        return y

    def _gn_bwd(self, name, x, dy, gb, eps, silu, hw, dx, add1=None, add2=None):
        G = self.eng.cfg.norm_num_groups
        st = self._stat(name + ".st", self.stat_f)
        red = self._stat(name + ".red", self.stat_b)
        ops.groupnorm_bwd_fused(x, dy, st, red, self.part_b[self._stat_slots[name + ".red"]], gb[0], gb[1], eps, silu, dx,
                                self.nb, hw, G, add1=add1, add2=add2)

    def _res_fwd(self, name, x, H, W, out):
        """diffusers ResnetBlock2D.  x: [nb, hw, cin] view, out: [nb, hw, cout] view."""
        r = self.eng.res[name]
        nb, hw = self.nb, H * W
        eps = self.eng.cfg.norm_eps
        self.trace_f[name] = out
        a1 = self._gn(name + ".gn1", x, r.n1, eps, True, hw)
        h1 = self.buf(name + ".h1", (nb, hw, r.cout))
        tp = self.bufs["temb.proj"][:, r.temb_off:r.temb_off + r.cout]
        ops.conv3x3(a1.view(nb, H, W, r.cin), r.c1f, h1.view(nb, H, W, r.cout), bias=r.c1bias, rowbias=tp, ws=self.ws)
        if r.scf is not None:
            # the 1x1 shortcut only meets the main chain again at conv2's residual input: side stream
            sc = self.buf(name + ".sc", (nb, hw, r.cout))
            ev = self._fork()
            with torch.cuda.stream(self.side):
                ops.gemm(x, r.scf, sc, bias=r.scbias, ws=self.ws)
                ev.record(self.side)
        else:
            sc, ev = x, None
        a2 = self._gn(name + ".gn2", h1, r.n2, eps, True, hw)
        if ev is not None:
            torch.cuda.current_stream().wait_event(ev)
        ops.conv3x3(a2.view(nb, H, W, r.cout), r.c2f, out.unflatten(1, (H, W)), bias=r.c2bias,
                    R=sc.unflatten(1, (H, W)), ws=self.ws)

    def _res_bwd(self, name, x, H, W, dout, dx, extra=None):
        """dx[nb,hw,cin] = d(ResnetBlock2D)/dx . dout (+ extra)."""
        r = self.eng.res[name]
        nb, hw = self.nb, H * W
        eps = self.eng.cfg.norm_eps
        self.trace_b[name] = dout
        ev = None
        if r.scf is not None:                              # shortcut dgrad runs beside the conv2 -> GN2 -> conv1 chain
            dsc = self.buf(name + ".dsc", (nb, hw, r.cin))
            ev = self._fork()
            with torch.cuda.stream(self.side):
                ops.gemm(dout, r.scb, dsc, ws=self.ws)
                ev.record(self.side)
        else:
            dsc = dout
        da2 = self.buf(name + ".da2", (nb, hw, r.cout))
        ops.conv3x3(dout.unflatten(1, (H, W)), r.c2b, da2.view(nb, H, W, r.cout), ws=self.ws)
        dh1 = self.buf(name + ".dh1", (nb, hw, r.cout))
        self._gn_bwd(name + ".gn2", self.bufs[name + ".h1"], da2, r.n2, eps, True, hw, dh1)
        da1 = self.buf(name + ".da1", (nb, hw, r.cin))
        ops.conv3x3(dh1.view(nb, H, W, r.cout), r.c1b, da1.view(nb, H, W, r.cin), ws=self.ws)
        if ev is not None:
            torch.cuda.current_stream().wait_event(ev)
        self._gn_bwd(name + ".gn1", x, da1, r.n1, eps, True, hw, dx, add1=dsc, add2=extra)

    def _kv2(self, name, c):
        """[2*nb*77, 2C] result of a layer's fused K|V projection and its two diagonal blocks as [nb, 77, C] views."""
        nb, L = self.nb, self.L
        kv = self.buf(name + ".kv2", (2 * nb * L, 2 * c))
        return kv, kv[:nb * L, :c].unflatten(0, (nb, L)), kv[nb * L:, c:].unflatten(0, (nb, L))

    def _xf_fwd(self, name, layer, x, H, W, out):
        """diffusers Transformer2DModel(1 BasicTransformerBlock) with XTIAttenProc semantics."""
        t = self.eng.xf[name]
        cfg = self.eng.cfg
        nb, hw, c, L = self.nb, H * W, t.c, self.L
        heads = c // 64
        rows = nb * hw
        self.trace_f[name] = out
        g = self._gn(name + ".gn", x, t.gn, cfg.xf_norm_eps, False, hw)
        t0 = self.buf(name + ".t0", (nb, hw, c))
        ops.gemm(g, t.pif, t0, bias=t.pibias, ws=self.ws)
        # --- attn1: self-attention (xti_attention_processor.py:25-26,32-33) ---
        n1 = self.buf(name + ".n1", (nb, hw, c))
        ops.layernorm_fwd(t0, t.ln[0][0], t.ln[0][1], cfg.ln_eps, n1, self.buf(name + ".ln1", (rows, 2), F32), rows)
        qkv = self.buf(name + ".qkv", (nb, hw, 3 * c))
        ops.gemm(n1, t.qkvf, qkv, ws=self.ws)
        o1 = self.buf(name + ".o1", (nb, hw, c))
        ops.attention_fwd(qkv[..., :c], qkv[..., c:2 * c], qkv[..., 2 * c:], o1,
                          self.buf(name + ".lse1", (nb, heads, hw), F32), heads, ws=self.attn_ws)
        t1 = self.buf(name + ".t1", (nb, hw, c))
        ops.gemm(o1, t.o1f, t1, bias=t.o1bias, R=t0, ws=self.ws)
        # --- attn2: XTI cross-attention, K from CONTEXT_TENSOR_i, V from CONTEXT_TENSOR_BYPASS_i (:16-22,38-42) ---
        n2 = self.buf(name + ".n2", (nb, hw, c))
        ops.layernorm_fwd(t1, t.ln[1][0], t.ln[1][1], cfg.ln_eps, n2, self.buf(name + ".ln2", (rows, 2), F32), rows)
        q2 = self.buf(name + ".q2", (nb, hw, c))
        ops.gemm(n2, t.q2f, q2, ws=self.ws)
        _, k2, v2 = self._kv2(name, c)                    # projected on the side stream at the start of forward()
        torch.cuda.current_stream().wait_event(self.ev_kv[layer])
        o2 = self.buf(name + ".o2", (nb, hw, c))
        ops.attention_fwd(q2, k2, v2, o2, self.buf(name + ".lse2", (nb, heads, hw), F32), heads)
        t2 = self.buf(name + ".t2", (nb, hw, c))
        ops.gemm(o2, t.o2f, t2, bias=t.o2bias, R=t1, ws=self.ws)
        # --- GEGLU feed-forward ---
        n3 = self.buf(name + ".n3", (nb, hw, c))
        ops.layernorm_fwd(t2, t.ln[2][0], t.ln[2][1], cfg.ln_eps, n3, self.buf(name + ".ln3", (rows, 2), F32), rows)
        hff = self.buf(name + ".hff", (nb, hw, 8 * c))
        ops.gemm(n3, t.ff1f, hff, bias=t.ff1bias, ws=self.ws)
        gg = self.buf(name + ".gg", (nb, hw, 4 * c))
        ops.geglu_fwd(hff, gg, rows)
        t3 = self.buf(name + ".t3", (nb, hw, c))
        ops.gemm(gg, t.ff2f, t3, bias=t.ff2bias, R=t2, ws=self.ws)
        ops.gemm(t3, t.pof, out, bias=t.pobias, R=x, ws=self.ws)

    def _xf_bwd(self, name, layer, x, H, W, dout, dx, extra=None, first=False):
        """Backward of _xf_fwd: writes d_ctx[k|v][layer]; unless `first`, also dx (+ dout residual + extra)."""
        t = self.eng.xf[name]
        cfg = self.eng.cfg
        nb, hw, c, L = self.nb, H * W, t.c, self.L
        heads = c // 64
        rows = nb * hw
        B = self.bufs
        self.trace_b[name] = dout
        dt3 = self.buf(name + ".dt3", (nb, hw, c))
        ops.gemm(dout, t.pob, dt3, ws=self.ws)
        dgg = self.buf(name + ".dgg", (nb, hw, 4 * c))
        ops.gemm(dt3, t.ff2b, dgg, ws=self.ws)
        dhff = self.buf(name + ".dhff", (nb, hw, 8 * c))
        ops.geglu_bwd(B[name + ".hff"], dgg, dhff, rows)
        dn3 = self.buf(name + ".dn", (nb, hw, c))
        ops.gemm(dhff, t.ff1b, dn3, ws=self.ws)
        dt2 = self.buf(name + ".dt2", (nb, hw, c))
        ops.layernorm_bwd(B[name + ".t2"], dn3, t.ln[2][0], B[name + ".ln3"], dt2, rows, add=dt3)
        do2 = self.buf(name + ".do", (nb, hw, c))
        ops.gemm(dt2, t.o2b, do2, ws=self.ws)
        dq2 = None if first else self.buf(name + ".dq2", (nb, hw, c))
        dkv2 = self.buf(name + ".dkv2", (2 * nb * L, c))              # [dK ; dV] rows: one operand for both context gradients
        dk2, dv2 = dkv2[:nb * L].view(nb, L, c), dkv2[nb * L:].view(nb, L, c)
        _, k2, v2 = self._kv2(name, c)
        if self._fin_ev is not None:                                     # the shared fp64 accumulator is free (zero) again
            torch.cuda.current_stream().wait_event(self._fin_ev)
        desc = ops.attention_bwd(B[name + ".q2"], k2, v2, B[name + ".o2"], B[name + ".lse2"], do2,
                                 self.buf(name + ".delta", (nb, heads, hw), F32), dq2, dk2, dv2, heads, dkv_acc=self.ws.dkv,
                                 defer_finish=True)
        # dK / dV -> bf16 and d CONTEXT_TENSOR_l / d CONTEXT_TENSOR_BYPASS_l (fp32 out) leave the chain: side stream, joined in
        # backward(); only dQ continues on the main chain
        self.ev_dkv[layer].record(torch.cuda.current_stream())
        self.side.wait_event(self.ev_dkv[layer])
        with torch.cuda.stream(self.side):
            ops.attention_dkv_finish(desc)
            self.ev_fin[layer].record(self.side)
            ops.gemm(dkv2, t.kv2b, self.d_ctx_store[layer], ws=self.ws)
        self._fin_ev = self.ev_fin[layer]
        if first:
            return                                                       # nothing upstream depends on the contexts
        dn2 = dn3
        ops.gemm(dq2, t.q2b, dn2, ws=self.ws)
        dt1 = self.buf(name + ".dt1", (nb, hw, c))
        ops.layernorm_bwd(B[name + ".t1"], dn2, t.ln[1][0], B[name + ".ln2"], dt1, rows, add=dt2)
        do1 = do2
        ops.gemm(dt1, t.o1b, do1, ws=self.ws)
        qkv = B[name + ".qkv"]
        dqkv = self.buf(name + ".dqkv", (nb, hw, 3 * c))
        ops.attention_bwd(qkv[..., :c], qkv[..., c:2 * c], qkv[..., 2 * c:], B[name + ".o1"], B[name + ".lse1"], do1,
                          B[name + ".delta"], dqkv[..., :c], dqkv[..., c:2 * c], dqkv[..., 2 * c:], heads, ws=self.attn_ws)
        dn1 = dn3
        ops.gemm(dqkv, t.qkvb, dn1, ws=self.ws)
        dt0 = dt3
        ops.layernorm_bwd(B[name + ".t0"], dn1, t.ln[0][0], B[name + ".ln1"], dt0, rows, add=dt1)
        dg = dt2
        ops.gemm(dt0, t.pib, dg, ws=self.ws)
        self._gn_bwd(name + ".gn", x, dg, t.gn, cfg.xf_norm_eps, False, hw, dx, add1=dout, add2=extra)

    # ---- whole network ------------------------------------------------------------------------------
    def forward(self) -> torch.Tensor:
        """eps = UNet(latents, timesteps, ctx) on the static buffers (coach.py:197-198)."""
        eng, cfg = self.eng, self.eng.cfg
        nb, h, w = self.nb, self.h, self.w
        ch = cfg.block_out_channels
        nlev = len(ch)
        B = self.bufs
        self.stat_f.zero_()
        ops.memset(self.part_f, 0xFF)
        # Side stream, concurrent with conv_in / the first GroupNorm: the time embedding (sinusoid -> Linear -> SiLU ->
        # Linear, then all 22 ResBlock projections in one launch; first needed by the first ResBlock's conv1), the bf16 cast
        # of the contexts and K = to_k(CONTEXT_TENSOR_l), V = to_v(CONTEXT_TENSOR_BYPASS_l) for all 16 layers
        # (xti_attention_processor.py:38-42)
        main = torch.cuda.current_stream()
        self.side.wait_stream(main)
        with torch.cuda.stream(self.side):
            sin = self.buf("temb.sin", (nb, ch[0]), F32)
            ops.timestep_sinusoid(self.timesteps, sin)
            e1 = self.buf("temb.e1", (nb, cfg.time_embed_dim), F32)
            ops.gemv(sin, eng.te1[0], eng.te1[1], e1)
            temb = self.buf("temb.e2", (nb, cfg.time_embed_dim), F32)
            ops.gemv(e1, eng.te2[0], eng.te2[1], temb, silu_in=True)
            ops.gemv(temb, eng.temb_w, eng.temb_b, self.buf("temb.proj", (nb, eng.temb_total), F32), silu_in=True)
            self.ev_temb.record(self.side)
            ctxb = self.buf("ctx.bf16", tuple(self.ctx_store.shape))
            ops.cast_f32_bf16(self.ctx_store, ctxb)
            for l, name in enumerate(self._xf_names()):
                t = eng.xf[name]
                ops.gemm(ctxb[l].view(2 * nb * self.L, -1), t.kv2f, self._kv2(name, t.c)[0], ws=self.ws)
                self.ev_kv[l].record(self.side)

        # concat buffers of the up path; skip tensors are produced directly into their slices
        cats: Dict[Tuple[int, int], torch.Tensor] = {}
        for i in range(nlev):
            lev = nlev - 1 - i
            hw = (h >> lev) * (w >> lev)
            for j in range(cfg.layers_per_block + 1):
                hid, skip, _ = up_block_resnet_channels(cfg, i, j)
                cats[(i, j)] = self.buf(f"cat.{i}.{j}", (nb, hw, hid + skip))
        order = [(i, j) for i in range(nlev) for j in range(cfg.layers_per_block + 1)]   # pop order
        n_skips = len(order)

        def skip_home(k: int) -> torch.Tensor:      # k-th pushed skip is popped (n_skips-1-k)-th
            i, j = order[n_skips - 1 - k]
            hid, sk, _ = up_block_resnet_channels(cfg, i, j)
            return cats[(i, j)][..., hid:hid + sk]

        self._skip_home = skip_home
        self._cats = cats
        k = 0
        x = skip_home(k); k += 1
        if eng.thin_gemm:
            col = self.buf("in.col", (nb * h * w, eng.conv_in_g.shape[1]))
            ops.im2col_thin(self.latents, col)
            ops.gemm(col, eng.conv_in_g, x, bias=eng.conv_in_b, ws=self.ws)
        else:
            ops.conv_in_fwd(self.latents, eng.conv_in_w, eng.conv_in_b, x.unflatten(1, (h, w)))
        main.wait_event(self.ev_temb)          # the time-embedding projections are first read by the next block's conv1
        layer = 0
        H, W = h, w
        for i in range(nlev):
            has_attn = cfg.down_has_attn[i]
            for j in range(cfg.layers_per_block):
                name = f"down_blocks.{i}.resnets.{j}"
                if has_attn:
                    ro = self.buf(name + ".out", (nb, H * W, ch[i]))
                    self._res_fwd(name, x, H, W, ro)
                    out = skip_home(k); k += 1
                    self._xf_fwd(f"down_blocks.{i}.attentions.{j}", layer, ro, H, W, out)
                    layer += 1
                else:
                    out = skip_home(k); k += 1
                    self._res_fwd(name, x, H, W, out)
                x = out
            if i < nlev - 1:
                wf, _, bias = eng.down[f"down_blocks.{i}.downsamplers.0"]
                Ho, Wo = H // 2, W // 2
                out = skip_home(k); k += 1
                if _S2_TMA:      # taps straight from the NHWC tensor map (element strides 2): no im2col buffer
                    ops.conv3x3(x.unflatten(1, (H, W)), wf, out.unflatten(1, (Ho, Wo)), bias=bias, ws=self.ws, stride=2)
                else:
                    col = self.buf(f"down.{i}.col", (nb * Ho * Wo, 9 * ch[i]))
                    ops.im2col_s2(x.unflatten(1, (H, W)), col)
                    ops.gemm(col, wf, out, bias=bias, ws=self.ws)
                x = out
                H, W = Ho, Wo
        c = ch[-1]
        m0 = self.buf("mid.r0", (nb, H * W, c))
        self._res_fwd("mid_block.resnets.0", x, H, W, m0)
        m1 = self.buf("mid.xf", (nb, H * W, c))
        self._xf_fwd("mid_block.attentions.0", layer, m0, H, W, m1)
        layer += 1
        # the last mid resnet writes straight into the first concat buffer
        hid0 = up_block_resnet_channels(cfg, 0, 0)[0]
        x = cats[(0, 0)][..., :hid0]
        self._res_fwd("mid_block.resnets.1", m1, H, W, x)
        rev = list(reversed(ch))
        has_attn_up = list(reversed(cfg.down_has_attn))
        for i in range(nlev):
            nres = cfg.layers_per_block + 1
            for j in range(nres):
                name = f"up_blocks.{i}.resnets.{j}"
                last = j == nres - 1
                # where does this layer's output go?  next concat slice, the upsampler input, or the final buffer
                if not last:
                    nhid = up_block_resnet_channels(cfg, i, j + 1)[0]
                    dest = cats[(i, j + 1)][..., :nhid]
                else:
                    dest = self.buf(f"up.{i}.out", (nb, H * W, rev[i]))
                if has_attn_up[i]:
                    ro = self.buf(name + ".out", (nb, H * W, rev[i]))
                    self._res_fwd(name, cats[(i, j)], H, W, ro)
                    self._xf_fwd(f"up_blocks.{i}.attentions.{j}", layer, ro, H, W, dest)
                    layer += 1
                else:
                    self._res_fwd(name, cats[(i, j)], H, W, dest)
                x = dest
            if i < nlev - 1:
                f, _, bias = eng.up[f"up_blocks.{i}.upsamplers.0"]
                u = self.buf(f"up.{i}.us", (nb, 4 * H * W, rev[i]))
                ops.upsample2x_fwd(x.unflatten(1, (H, W)), u.view(nb, 2 * H, 2 * W, rev[i]))
                H, W = 2 * H, 2 * W
                nhid = up_block_resnet_channels(cfg, i + 1, 0)[0]
                dest = cats[(i + 1, 0)][..., :nhid]
                ops.conv3x3(u.view(nb, H, W, rev[i]), f, dest.unflatten(1, (H, W)), bias=bias, ws=self.ws)
                x = dest
        assert layer == self.n_layers and k == n_skips
        self._final_x = x
        y = self._gn("out.gn", x, eng.norm_out, cfg.norm_eps, True, H * W)
        if eng.thin_gemm:
            e8 = self.buf("out.eps8", (nb, H, W, 8), F32)
            ops.conv3x3(y.view(nb, H, W, ch[0]), eng.conv_out_g, e8, bias=eng.conv_out_gb, ws=self.ws, force_bn=64, force_split=4)
            ops.nhwc_to_nchw_thin(e8, self.eps)
        else:
            ops.conv_out_fwd(y.view(nb, H, W, ch[0]), eng.conv_out_w, eng.conv_out_b, self.eps)
        self._saved = True
        return self.eps

    def backward(self) -> torch.Tensor:
        """d_ctx[k|v][layer] = d<eps, d_eps>/d ctx for the activations of the last forward (coach.py:214, dgrad only)."""
        assert self._saved, "backward() needs a forward() on this plan first"
        self._fin_ev = None                      # no deferred dK / dV conversion pending on the side stream yet
        eng, cfg = self.eng, self.eng.cfg
        nb, h, w = self.nb, self.h, self.w
        ch = cfg.block_out_channels
        nlev = len(ch)
        rev = list(reversed(ch))
        has_attn_up = list(reversed(cfg.down_has_attn))
        cats = self._cats
        H, W = h, w
        self.stat_b.zero_()
        ops.memset(self.part_b, 0xFF)
        dy = self.buf("bwd.dy", (nb, H * W, ch[0]))
        if eng.thin_gemm:
            dcol = self.buf("bwd.eps.col", (nb * H * W, eng.conv_out_bwd_g.shape[1]))
            ops.im2col_thin(self.d_eps, dcol)
            ops.gemm(dcol, eng.conv_out_bwd_g, dy, ws=self.ws)
        else:
            ops.conv_out_bwd(self.d_eps, eng.conv_out_w, dy.view(nb, H, W, ch[0]))
        dcur = self.buf(f"bwd.up.{nlev - 1}.out", (nb, H * W, ch[0]))
        self._gn_bwd("out.gn", self._final_x, dy, eng.norm_out, cfg.norm_eps, True, H * W, dcur)
        layer = self.n_layers - 1
        dcat: Dict[Tuple[int, int], torch.Tensor] = {}
        for i in range(nlev - 1, -1, -1):
            nres = cfg.layers_per_block + 1
            if i < nlev - 1:
                # dcur is the gradient of the upsampler conv output (= hidden slice of the next block's first concat)
                _, bw, _ = eng.up[f"up_blocks.{i}.upsamplers.0"]
                du = self.buf(f"bwd.up.{i}.us", (nb, H * W, rev[i]))
                ops.conv3x3(dcur.unflatten(1, (H, W)), bw, du.view(nb, H, W, rev[i]), ws=self.ws)
                H, W = H // 2, W // 2
                dcur = self.buf(f"bwd.up.{i}.out", (nb, H * W, rev[i]))
                ops.upsample2x_bwd(du.view(nb, 2 * H, 2 * W, rev[i]), dcur.view(nb, H, W, rev[i]))
            for j in range(nres - 1, -1, -1):
                name = f"up_blocks.{i}.resnets.{j}"
                hid, sk, _ = up_block_resnet_channels(cfg, i, j)
                d = self.buf(f"bwd.cat.{i}.{j}", (nb, H * W, hid + sk))
                dcat[(i, j)] = d
                if has_attn_up[i]:
                    dro = self.buf(name + ".dout", (nb, H * W, rev[i]))
                    self._xf_bwd(f"up_blocks.{i}.attentions.{j}", layer, self.bufs[name + ".out"], H, W, dcur, dro)
                    layer -= 1
                    self._res_bwd(name, cats[(i, j)], H, W, dro, d)
                else:
                    self._res_bwd(name, cats[(i, j)], H, W, dcur, d)
                dcur = d[..., :hid]
        order = [(i, j) for i in range(nlev) for j in range(cfg.layers_per_block + 1)]
        n_skips = len(order)

        def dskip(k: int) -> torch.Tensor:
            i, j = order[n_skips - 1 - k]
            hid, sk, _ = up_block_resnet_channels(cfg, i, j)
            return dcat[(i, j)][..., hid:hid + sk]

        # mid block: dcur = d(mid.resnets.1 output)
        c = ch[-1]
        dm1 = self.buf("bwd.mid.xf", (nb, H * W, c))
        self._res_bwd("mid_block.resnets.1", self.bufs["mid.xf"], H, W, dcur, dm1)
        dm0 = self.buf("bwd.mid.r0", (nb, H * W, c))
        self._xf_bwd("mid_block.attentions.0", layer, self.bufs["mid.r0"], H, W, dm1, dm0)
        layer -= 1
        k = n_skips - 1                                   # skip produced by the last down resnet
        dx = self.buf(f"bwd.skip.{k}", (nb, H * W, c))
        self._res_bwd("mid_block.resnets.0", self._skip_home(k), H, W, dm0, dx, extra=dskip(k))
        dcur = dx                                         # total gradient of skip k
        for i in range(nlev - 1, -1, -1):
            has_attn = cfg.down_has_attn[i]
            if i < nlev - 1:
                # dcur = total gradient of the downsampler output (skip k); its input is skip k-1
                _, wb, _ = eng.down[f"down_blocks.{i}.downsamplers.0"]
                dcol = self.buf(f"bwd.down.{i}.col", (nb * H * W, 9 * ch[i]))
                ops.gemm(dcur, wb, dcol, ws=self.ws)
                H, W = 2 * H, 2 * W
                k -= 1
                dx = self.buf(f"bwd.skip.{k}", (nb, H * W, ch[i]))
                ops.col2im_s2(dcol, dx.view(nb, H, W, ch[i]), add=dskip(k).unflatten(1, (H, W)))
                dcur = dx
            for j in range(cfg.layers_per_block - 1, -1, -1):
                name = f"down_blocks.{i}.resnets.{j}"
                # dcur = total gradient of skip k = output of this layer; its input is skip k-1
                xin = self._skip_home(k - 1)
                if has_attn:
                    first = layer == 0
                    dro = None if first else self.buf(name + ".dout", (nb, H * W, ch[i]))
                    self._xf_bwd(f"down_blocks.{i}.attentions.{j}", layer, self.bufs[name + ".out"], H, W, dcur, dro,
                                 first=first)
                    layer -= 1
                    if first:
                        assert layer == -1
                        torch.cuda.current_stream().wait_stream(self.side)
                        return self.d_ctx
                    dres = dro
                else:
                    dres = dcur
                k -= 1
                dx = self.buf(f"bwd.skip.{k}", (nb, H * W, xin.shape[-1]))
                self._res_bwd(name, xin, H, W, dres, dx, extra=dskip(k) if k > 0 else None)
                dcur = dx
        torch.cuda.current_stream().wait_stream(self.side)
        return self.d_ctx

    # ---- graphs -------------------------------------------------------------------------------------
    def train_step(self) -> None:
        """forward + fp32 MSE against `target` + backward (coach.py:197-214) on the static buffers."""
        self.forward()
        ops.mse_loss(self.eps, self.target, self.loss, self.d_eps)
        self.backward()

    def _capture(self, fn):
        """Record fn's launch sequence in a CUDA graph.  Every kernel must already have run once eagerly
        (function attributes set, all buffers allocated), so nothing but launches happens here."""
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        ops.launch_count_reset()
        with torch.cuda.graph(g):
            fn()
        return g, ops.launch_count()

    def capture(self, what: str = "train"):
        """what: 'train' (forward + MSE + backward), 'fwd', 'bwd'.  Needs one eager run of the same sequence first."""
        if what not in self.graphs:
            if not self._saved or (what != "fwd" and "bwd.dy" not in self.bufs):
                s = torch.cuda.Stream()
                s.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(s):
                    self.train_step()
                torch.cuda.current_stream().wait_stream(s)
            fn = {"train": self.train_step, "fwd": self.forward, "bwd": self.backward}[what]
            self.graphs[what], self.launches[what] = self._capture(fn)
        return self.graphs[what]

    def run_forward(self) -> torch.Tensor:
        """Used by the drop-in UNet: eager on the first call, CUDA-graph replay afterwards.  Forward-only users (the
        denoise loop of sd_pipeline_call) get the forward graph after one eager pass; the backward graph is captured as
        soon as a backward has run once (both captures happen here, on the caller's thread: the autograd thread only
        ever replays)."""
        if self.eng.use_graphs and "fwd" in self.graphs:
            if "bwd" not in self.graphs and "bwd.dy" in self.bufs:
                self.capture("bwd")
            self.graphs["fwd"].replay()
            self._saved = True
            return self.eps
        out = self.forward()
        if self.eng.use_graphs:
            self.capture("fwd")
            if "bwd.dy" in self.bufs:
                self.capture("bwd")
        return out

    def run_backward(self) -> torch.Tensor:
        if "bwd" in self.graphs:
            self.graphs["bwd"].replay()
            return self.d_ctx
        return self.backward()
