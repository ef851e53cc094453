"""CLIP text-transformer encoder on the sm_100a library: the heavy part of ViewNeTI's conditioning path
(SURVEY.md 8f #1).

Reference: models/neti_clip_text_encoder.py:44-56,101-116 builds `self.encoder = CLIPEncoder(config)` (transformers,
pinned 4.27.4) and training/coach.py:276-311 runs the whole text model ONCE PER UNET LAYER (16 passes of 23 pre-LN
transformer layers over [B, 77, 1024]) with only the placeholder rows of the input embeddings differing between passes.
`CLIPEncoder` below has the call signature of transformers' class, so it replaces that one attribute; because it is
batch-agnostic the 16 per-layer passes can be stacked into ONE [16*B, 77, 1024] call.

What runs where: LayerNorm -> fused q/k/v projection -> causal attention -> out projection (+residual) -> LayerNorm ->
fc1 -> erf-GELU -> fc2 (+residual); every projection is vn_gemm (tcgen05), the norms are vn_layernorm_*, GELU and the
77-token causal attention are vn_gelu_* / vn_seq_attention_* (csrc/vn_clip.cu).  The text model is frozen
(coach.py:649-652 freezes everything but the mappers), so the backward computes data gradients only: no weight
gradients, no saved GEMM inputs; it returns d(inputs_embeds), whose placeholder rows are the mapper-output gradients.
There is no CPU fallback.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Optional, Tuple

import torch

from .. import ops

BF = torch.bfloat16
F32 = torch.float32


@dataclass(frozen=True)
class ClipEncoderConfig:
    hidden_size: int = 1024            # SD-2.1 text encoder (OpenCLIP ViT-H/14): 1024 wide, 16 heads, 23 layers used
    num_attention_heads: int = 16
    num_hidden_layers: int = 23
    intermediate_size: int = 4096
    layer_norm_eps: float = 1e-5

    @property
    def head_dim(self) -> int:
        return self.hidden_size // self.num_attention_heads


SD21_TEXT = ClipEncoderConfig()


def init_state_dict(cfg: ClipEncoderConfig = SD21_TEXT, seed: int = 0) -> Dict[str, torch.Tensor]:
    """Seeded random weights under transformers' CLIPEncoder key names at the configured shapes (no checkpoints exist
    offline; a real text encoder's `text_model.encoder.state_dict()` loads the same way).  Residual branches are damped
    by (2*layers)**-0.5 so the stream stays O(1) through all layers."""
    import math
    g = torch.Generator().manual_seed(seed)
    C, I, nl = cfg.hidden_size, cfg.intermediate_size, cfg.num_hidden_layers
    damp = (2 * nl) ** -0.5
    sd: Dict[str, torch.Tensor] = {}
    for i in range(nl):
        p = f"layers.{i}."
        for n in ("q", "k", "v"):
            sd[p + f"self_attn.{n}_proj.weight"] = torch.randn(C, C, generator=g) / math.sqrt(C)
            sd[p + f"self_attn.{n}_proj.bias"] = 0.1 * torch.randn(C, generator=g)
        sd[p + "self_attn.out_proj.weight"] = torch.randn(C, C, generator=g) * damp / math.sqrt(C)
        sd[p + "self_attn.out_proj.bias"] = 0.05 * torch.randn(C, generator=g)
        sd[p + "mlp.fc1.weight"] = torch.randn(I, C, generator=g) / math.sqrt(C)
        sd[p + "mlp.fc1.bias"] = 0.1 * torch.randn(I, generator=g)
        sd[p + "mlp.fc2.weight"] = torch.randn(C, I, generator=g) * damp / math.sqrt(I)
        sd[p + "mlp.fc2.bias"] = 0.05 * torch.randn(C, generator=g)
        for n in ("layer_norm1", "layer_norm2"):
            sd[p + n + ".weight"] = 1 + 0.1 * torch.randn(C, generator=g)
            sd[p + n + ".bias"] = 0.05 * torch.randn(C, generator=g)
    return sd


class _Layer:
    pass


class ClipEncoderEngine:
    """Frozen weights (bf16 forward operand + pre-transposed dgrad operand) and one static plan per (nseq, L)."""

    def __init__(self, state_dict: Dict[str, torch.Tensor], cfg: ClipEncoderConfig = SD21_TEXT, device="cuda"):
        self.cfg = cfg
        self.dev = torch.device(device)
        if self.dev.type != "cuda":
            raise ops._abi.VNError("ClipEncoderEngine needs a CUDA device: there is no CPU fallback")
        if cfg.head_dim != 64:
            raise ops._abi.VNError("vn_seq_attention supports head_dim 64 only")
        self.layers = []
        f32 = lambda k: state_dict[k].to(device=self.dev, dtype=F32).contiguous()      # noqa: E731

        def lin(w):
            wb = w.to(device=self.dev, dtype=BF)
            return wb.contiguous(), wb.t().contiguous()

        for i in range(cfg.num_hidden_layers):
            p = f"layers.{i}."
            l = _Layer()
            l.ln1 = (f32(p + "layer_norm1.weight"), f32(p + "layer_norm1.bias"))
            l.ln2 = (f32(p + "layer_norm2.weight"), f32(p + "layer_norm2.bias"))
            qkv = torch.cat([state_dict[p + f"self_attn.{n}_proj.weight"] for n in "qkv"], 0)
            l.qkv_f, l.qkv_b = lin(qkv)
            l.qkv_bias = torch.cat([state_dict[p + f"self_attn.{n}_proj.bias"] for n in "qkv"], 0).to(self.dev, F32).contiguous()
            l.o_f, l.o_b = lin(state_dict[p + "self_attn.out_proj.weight"])
            l.o_bias = f32(p + "self_attn.out_proj.bias")
            l.fc1_f, l.fc1_b = lin(state_dict[p + "mlp.fc1.weight"])
            l.fc1_bias = f32(p + "mlp.fc1.bias")
            l.fc2_f, l.fc2_b = lin(state_dict[p + "mlp.fc2.weight"])
            l.fc2_bias = f32(p + "mlp.fc2.bias")
            self.layers.append(l)
        self._plans: Dict[Tuple[int, int], "_ClipPlan"] = {}
        # attention core: "tc" = the tcgen05 flash-attention kernels with their causal mask (one 128 x 128 tile per
        # (sequence, head)); "seq" = the CUDA-core short-sequence kernels of csrc/vn_clip.cu (kept as cross-check)
        import os
        self.attn_impl = os.environ.get("VN_CLIP_ATTN", "tc")

    def plan(self, nseq: int, L: int, train: bool = True) -> "_ClipPlan":
        key = (nseq, L, train)
        if key not in self._plans:
            self._plans[key] = _ClipPlan(self, nseq, L, train)
        return self._plans[key]


class _ClipPlan:
    """Static buffers + launch order for [nseq, L, hidden]; forward / backward are CUDA-graph capturable."""

    def __init__(self, eng: ClipEncoderEngine, nseq: int, L: int, train: bool = True):
        cfg = eng.cfg
        self.eng, self.nseq, self.L, self.train = eng, nseq, L, train
        dev, C, I, nl = eng.dev, cfg.hidden_size, cfg.intermediate_size, cfg.num_hidden_layers
        z = lambda *s, dt=BF: torch.zeros(*s, dtype=dt, device=dev)      # noqa: E731
        rows = nseq * L
        self.rows = rows
        self.ws = ops.Workspace(rows, I, dev)
        self.x_in = z(nseq, L, C, dt=F32)            # inputs_embeds (fp32, as the reference hands them over)
        self.dy_in = z(nseq, L, C, dt=F32)
        self.dx_out = z(nseq, L, C, dt=F32)
        # train: kept per layer for the backward (both LayerNorm inputs + statistics, q/k/v, attention output, lse, fc1
        # output).  inference (no gradient wanted): every layer reuses ONE set of buffers, the stream ping-pongs between two.
        per = (lambda mk, n: [mk() for _ in range(n)]) if train else (lambda mk, n: [mk()] * n)      # noqa: E731
        # The residual stream (x: layer inputs, xm: after the attention branch) is kept in FP32: it is rounded twice per
        # layer, 46 times over the 23 layers, and in bf16 those roundings alone cost ~8e-3 of relative error in the mapper
        # gradient (measured: 1.08e-2 end to end against the 1e-2 contract, profiles/r2_parity_figures.jsonl).  Branch
        # tensors (GEMM operands) stay bf16.  x[0] is the fp32 input itself.
        if train:
            self.x = [self.x_in] + [z(nseq, L, C, dt=F32) for _ in range(nl)]
        else:
            pp = [z(nseq, L, C, dt=F32), z(nseq, L, C, dt=F32)]
            self.x = [self.x_in] + [pp[i % 2] for i in range(nl)]
        self.xm = per(lambda: z(nseq, L, C, dt=F32), nl)
        self.y_out = self.x[-1]                      # last_hidden_state (before the text model's final LayerNorm)
        self.st1 = per(lambda: z(rows, 2, dt=F32), nl)
        self.st2 = per(lambda: z(rows, 2, dt=F32), nl)
        self.qkv = per(lambda: z(nseq, L, 3 * C), nl)
        self.o = per(lambda: z(nseq, L, C), nl)
        self.lse = per(lambda: z(nseq, cfg.num_attention_heads, L, dt=F32), nl)
        self.delta = z(nseq, cfg.num_attention_heads, L, dt=F32)        # scratch of the tcgen05 attention backward
        self.h1 = per(lambda: z(nseq, L, I), nl)
        # scratch shared by all layers
        self.n = z(nseq, L, C)
        self.g = z(nseq, L, I)
        if train:
            self.dq = z(nseq, L, 3 * C)
            # gradient of the residual stream: fp32, with a bf16 copy of each tensor (the dgrad GEMMs' A operand).
            # da / da32: d(layer output), overwritten in place by d(layer input); db / db32: d(xm)
            self.da, self.db = z(nseq, L, C), z(nseq, L, C)
            self.da32, self.db32 = z(nseq, L, C, dt=F32), z(nseq, L, C, dt=F32)
        self.graphs: Dict[str, torch.cuda.CUDAGraph] = {}
        self._saved = False
        self.generation = 0           # bumped by every training forward; a backward checks that it is still its own

    # ------------------------------------------------------------------------------------------------
    def forward(self) -> torch.Tensor:
        eng, cfg = self.eng, self.eng.cfg
        C, heads, rows = cfg.hidden_size, cfg.num_attention_heads, self.rows
        scale = cfg.head_dim ** -0.5
        for i, l in enumerate(eng.layers):
            x, xm, qkv = self.x[i], self.xm[i], self.qkv[i]
            ops.layernorm_fwd_f32(x, l.ln1[0], l.ln1[1], cfg.layer_norm_eps, self.n, self.st1[i], rows)
            ops.gemm(self.n, l.qkv_f, qkv, bias=l.qkv_bias, ws=self.ws)
            if eng.attn_impl == "tc":
                ops.attention_fwd(qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:], self.o[i], self.lse[i], heads,
                                  scale=scale, causal=True)
            else:
                ops.seq_attention_fwd(qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:], self.o[i], self.lse[i], heads,
                                      scale=scale, causal=True)
            ops.gemm(self.o[i], l.o_f, xm, bias=l.o_bias, R=x, ws=self.ws)
            ops.layernorm_fwd_f32(xm, l.ln2[0], l.ln2[1], cfg.layer_norm_eps, self.n, self.st2[i], rows)
            ops.gemm(self.n, l.fc1_f, self.h1[i], bias=l.fc1_bias, ws=self.ws)
            ops.gelu_fwd(self.h1[i], self.g, rows)
            ops.gemm(self.g, l.fc2_f, self.x[i + 1], bias=l.fc2_bias, R=xm, ws=self.ws)
        self._saved = True
        return self.y_out

    def backward(self) -> torch.Tensor:
        """dx_out = d<y_out, dy_in> / d x_in for the activations of the last forward (data gradients only)."""
        assert self.train, "this plan was built for inference (no activations kept)"
        assert self._saved, "backward() needs a forward() on this plan first"
        eng, cfg = self.eng, self.eng.cfg
        C, heads, rows = cfg.hidden_size, cfg.num_attention_heads, self.rows
        scale = cfg.head_dim ** -0.5
        # (dy32, dy): gradient of the layer's output stream in fp32 and its bf16 copy
        dy32, dy = self.dy_in, self.da
        dxm32, dxm = self.db32, self.db
        ops.cast_f32_bf16(dy32, dy)
        for i in range(cfg.num_hidden_layers - 1, -1, -1):
            l = eng.layers[i]
            qkv = self.qkv[i]
            ops.gemm(dy, l.fc2_b, self.g, ws=self.ws)                        # d gelu-out
            ops.gelu_bwd(self.h1[i], self.g, self.g, rows)                   # d fc1-out (in place)
            ops.gemm(self.g, l.fc1_b, self.n, ws=self.ws)                    # d LN2-out
            ops.layernorm_bwd_f32(self.xm[i], self.n, l.ln2[0], self.st2[i], dxm32, rows, add=dy32, dx_bf16=dxm)
            ops.gemm(dxm, l.o_b, self.n, ws=self.ws)                         # d attention-out
            if eng.attn_impl == "tc":
                ops.attention_bwd(qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:], self.o[i], self.lse[i], self.n,
                                  self.delta, self.dq[..., :C], self.dq[..., C:2 * C], self.dq[..., 2 * C:], heads,
                                  scale=scale, causal=True)
            else:
                ops.seq_attention_bwd(qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:], self.o[i], self.lse[i], self.n,
                                      self.dq[..., :C], self.dq[..., C:2 * C], self.dq[..., 2 * C:], heads, scale=scale,
                                      causal=True)
            ops.gemm(self.dq, l.qkv_b, self.n, ws=self.ws)                   # d LN1-out
            # d(layer input) overwrites d(layer output): both dy32 and dy were consumed above (stream order)
            last = i == 0
            dy32 = self.dx_out if last else self.da32
            ops.layernorm_bwd_f32(self.x[i], self.n, l.ln1[0], self.st1[i], dy32, rows, add=dxm32,
                                  dx_bf16=None if last else dy)
        return self.dx_out

    # ------------------------------------------------------------------------------------------------
    def _capture(self, what: str) -> None:
        fn = self.forward if what == "fwd" else self.backward
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            fn()
        self.graphs[what] = g

    def run_forward(self, use_graphs: bool = True) -> torch.Tensor:
        if "fwd" in self.graphs:
            self.graphs["fwd"].replay()
            self._saved = True
            return self.y_out
        out = self.forward()                       # first call: eager (sets kernel attributes, allocates nothing new)
        if use_graphs and not self.train:
            self._capture("fwd")
        elif use_graphs:
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                self.backward()                    # every kernel of the backward must have run once before capture
            torch.cuda.current_stream().wait_stream(s)
            self._capture("fwd")
            self._capture("bwd")
        return out

    def run_backward(self) -> torch.Tensor:
        if "bwd" in self.graphs:
            self.graphs["bwd"].replay()
            return self.dx_out
        return self.backward()


class _ClipEncoderFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x: torch.Tensor, plan: _ClipPlan, use_graphs: bool) -> torch.Tensor:
        plan.x_in.copy_(x)
        y = plan.run_forward(use_graphs)
        plan.generation += 1
        ctx.plan, ctx.gen = plan, plan.generation
        ctx.in_dtype = x.dtype
        return y.to(x.dtype, copy=True)

    @staticmethod
    def backward(ctx, dy: torch.Tensor):
        plan = ctx.plan
        if ctx.gen != plan.generation:
            raise ops._abi.VNError("backward() after another forward of the same shape through this CLIPEncoder: the static "
                                   "activation plan was overwritten (run each backward before the next training forward - "
                                   "gradient accumulation works micro-step by micro-step, coach.py:158-218)")
        plan.dy_in.copy_(dy)
        dx = plan.run_backward()
        return dx.to(ctx.in_dtype, copy=True), None, None


class _EncoderOutput(tuple):
    """`encoder_outputs[0]`, `.last_hidden_state`, `.hidden_states`, `.attentions` as neti_clip_text_encoder.py:110-116,
    205-224 reads them."""

    def __new__(cls, last_hidden_state):
        o = super().__new__(cls, (last_hidden_state,))
        o.last_hidden_state = last_hidden_state
        o.hidden_states = None
        o.attentions = None
        return o


class CLIPEncoder(torch.nn.Module):
    """Drop-in for `transformers.models.clip.modeling_clip.CLIPEncoder` as used at models/neti_clip_text_encoder.py:53,
    101-108: `encoder(inputs_embeds=..., attention_mask=None, causal_attention_mask=..., ...)[0]`.  The causal mask is
    built in (CLIPTextTransformer always passes one); a padding `attention_mask` is not supported (the reference never
    passes one: coach.py / prompt_manager.py call the text encoder without it)."""

    def __init__(self, state_dict: Dict[str, torch.Tensor], config: ClipEncoderConfig = SD21_TEXT, device="cuda"):
        super().__init__()
        self.config = config
        self.engine = ClipEncoderEngine(state_dict, config, device)
        self.use_graphs = True

    def forward(self, inputs_embeds: torch.Tensor, attention_mask: Optional[torch.Tensor] = None,
                causal_attention_mask: Optional[torch.Tensor] = None, output_attentions: Optional[bool] = None,
                output_hidden_states: Optional[bool] = None, return_dict: Optional[bool] = None, **kwargs):
        if attention_mask is not None:
            raise ops._abi.VNError("CLIPEncoder: a padding attention_mask is not supported (the reference passes none)")
        if output_attentions or output_hidden_states:
            raise ops._abi.VNError("CLIPEncoder: attention maps / per-layer hidden states are not materialised")
        if not inputs_embeds.is_cuda:
            raise ops._abi.VNError("CLIPEncoder needs CUDA tensors (no CPU fallback)")
        nseq, L, C = inputs_embeds.shape
        assert C == self.config.hidden_size
        if not (torch.is_grad_enabled() and inputs_embeds.requires_grad):
            plan = self.engine.plan(nseq, L, train=False)          # inference: one set of layer buffers, forward graph only
            plan.x_in.copy_(inputs_embeds)
            return _EncoderOutput(plan.run_forward(self.use_graphs).to(inputs_embeds.dtype, copy=True))
        plan = self.engine.plan(nseq, L)
        y = _ClipEncoderFn.apply(inputs_embeds, plan, self.use_graphs)
        return _EncoderOutput(y)
