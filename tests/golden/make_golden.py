"""Generates tests/golden/xti_attn.pt by executing the REFERENCE's own attention processor, unmodified, from
/root/reference/models/xti_attention_processor.py (it is imported, never copied).

The reference imports `diffusers.models.cross_attention.CrossAttention` only for a type annotation; diffusers is not
installable here, so a stub module exposes the oracle's `CrossAttention` (oracle/unet_sd21.py) under that name.  The
processor then runs its real code (context selection, this_idx bookkeeping, K/V source split, head reshapes,
attention, out-projection) against that module's members.  Run in the build container only:

    python tests/golden/make_golden.py
"""
import importlib.util
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.unet_sd21 import CrossAttention  # noqa: E402

REF = "/root/reference/models/xti_attention_processor.py"


def load_reference_processor():
    for name in ("diffusers", "diffusers.models", "diffusers.models.cross_attention"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["diffusers.models.cross_attention"].CrossAttention = CrossAttention
    spec = importlib.util.spec_from_file_location("ref_xti_attention_processor", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.XTIAttenProc


def main():
    XTIAttenProc = load_reference_processor()
    g = torch.Generator().manual_seed(1234)
    heads, qdim, cdim, B, N, L = 2, 128, 192, 2, 64, 77
    cross = CrossAttention(qdim, cdim, heads)
    selfa = CrossAttention(qdim, None, heads)
    for m in (cross, selfa):
        for p in m.parameters():
            p.data = torch.randn(p.shape, generator=g) * (p.shape[-1] ** -0.5 if p.ndim > 1 else 0.05)
        m.processor = XTIAttenProc()
    hidden = torch.randn(B, N, qdim, generator=g)
    ctx = {"this_idx": 3}
    for i in range(16):
        ctx[f"CONTEXT_TENSOR_{i}"] = torch.randn(B, L, cdim, generator=g)
        ctx[f"CONTEXT_TENSOR_BYPASS_{i}"] = torch.randn(B, L, cdim, generator=g)
    used = ("this_idx", "CONTEXT_TENSOR_3", "CONTEXT_TENSOR_BYPASS_3", "CONTEXT_TENSOR_15", "CONTEXT_TENSOR_5")
    out = {"heads": heads, "hidden": hidden, "ctx": {k: v for k, v in ctx.items() if k in used},
           "cross_state": cross.state_dict(), "self_state": selfa.state_dict()}
    with torch.no_grad():
        d = dict(ctx)
        out["dict_bypass"] = cross(hidden, encoder_hidden_states=d)
        out["dict_bypass_this_idx_after"] = d["this_idx"]
        d = {k: v for k, v in ctx.items() if "BYPASS" not in k}
        d["this_idx"] = 15
        out["dict_nobypass_idx15"] = cross(hidden, encoder_hidden_states=d)
        out["dict_nobypass_this_idx_after"] = d["this_idx"]
        out["tensor_ctx"] = cross(hidden, encoder_hidden_states=ctx["CONTEXT_TENSOR_5"])
        out["self"] = selfa(hidden, encoder_hidden_states=None)
    # gradients w.r.t. the K / V contexts through the reference processor (autograd)
    d = {k: (v.clone().requires_grad_(True) if torch.is_tensor(v) else v) for k, v in ctx.items()}
    h = hidden.clone().requires_grad_(True)
    y = cross(h, encoder_hidden_states=d)
    w = torch.randn(y.shape, generator=g)
    (y * w).sum().backward()
    out["grad_w"] = w
    out["grad_ctx_k"] = d["CONTEXT_TENSOR_3"].grad
    out["grad_ctx_v"] = d["CONTEXT_TENSOR_BYPASS_3"].grad
    out["grad_hidden"] = h.grad
    path = os.path.join(ROOT, "tests", "golden", "xti_attn.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
