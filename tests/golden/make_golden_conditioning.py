"""Generates tests/golden/neti_conditioning.pt by running the REFERENCE's own conditioning path
(/root/reference/models/neti_clip_text_encoder.py + net_clip_text_embedding.py + neti_mapper.py, imported unmodified)
on the CPU of the build container, in the loop of /root/reference/training/coach.py:276-311 (one text-encoder pass per
UNet layer).

Compatibility shims (no arithmetic of the reference is touched):
  * `ipdb` is an empty stub module; `.cuda()` is identity (the reference hard-codes it in the positional encodings);
  * the reference targets transformers 4.27.4, this container has 5.5.0: `_expand_mask` (imported, only used for padding
    masks the path never passes) is supplied as a stub attribute; `_build_causal_attention_mask` (removed upstream) is
    restored with its 4.27 definition (-inf above the diagonal); the encoder attribute is wrapped so that the 4.27 keyword
    `causal_attention_mask` reaches the 5.5 `CLIPEncoder` as its additive `attention_mask`.
A small config keeps the fixture small: hidden 128, 2 heads (head_dim 64), 2 layers, MLP 256, vocab 120, 77 positions.

    python tests/golden/make_golden_conditioning.py
"""
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.modules.setdefault("ipdb", types.ModuleType("ipdb"))
torch.Tensor.cuda = lambda self, *a, **k: self
torch.nn.Module.cuda = lambda self, *a, **k: self
import transformers.models.clip.modeling_clip as hf_clip  # noqa: E402

if not hasattr(hf_clip, "_expand_mask"):
    hf_clip._expand_mask = lambda mask, dtype, tgt_len=None: (_ for _ in ()).throw(RuntimeError("padding masks are not on this path"))
sys.path.insert(0, "/root/reference")
from constants import UNET_LAYERS  # noqa: E402
from models.neti_clip_text_encoder import NeTICLIPTextTransformer  # noqa: E402
from models.neti_mapper import NeTIMapper  # noqa: E402
from utils.types import NeTIBatch, PESigmas  # noqa: E402

HIDDEN, HEADS, LAYERS, INTER, VOCAB = 128, 2, 2, 256, 120


class _EncoderCompat(torch.nn.Module):
    """4.27 call convention -> 5.5 CLIPEncoder: the causal mask is the additive attention mask."""

    def __init__(self, enc):
        super().__init__()
        self.enc = enc

    def forward(self, inputs_embeds=None, attention_mask=None, causal_attention_mask=None, output_attentions=None,
                output_hidden_states=None, return_dict=None):
        assert attention_mask is None
        return self.enc(inputs_embeds=inputs_embeds, attention_mask=causal_attention_mask)


def causal_mask_427(self, bsz, seq_len, dtype, device=None):
    mask = torch.empty(bsz, seq_len, seq_len, dtype=dtype)
    mask.fill_(torch.tensor(torch.finfo(dtype).min))
    mask.triu_(1)
    return mask.unsqueeze(1)


def main():
    torch.manual_seed(11)
    cfg = hf_clip.CLIPTextConfig(hidden_size=HIDDEN, intermediate_size=INTER, num_hidden_layers=LAYERS,
                                 num_attention_heads=HEADS, hidden_act="gelu", max_position_embeddings=77, vocab_size=VOCAB)
    cfg._attn_implementation = "eager"
    model = NeTICLIPTextTransformer(cfg).eval()
    # HF initialises biases / LayerNorms trivially: perturb them so every term of the path is exercised
    g = torch.Generator().manual_seed(12)
    with torch.no_grad():
        for n, p in model.named_parameters():
            if n.endswith(".bias"):
                p.add_(0.05 * torch.randn(p.shape, generator=g))
            elif "layer_norm" in n and n.endswith(".weight"):
                p.add_(0.1 * torch.randn(p.shape, generator=g))
            elif n.endswith("proj.weight") or "fc" in n or "embedding" in n:
                p.copy_(torch.randn(p.shape, generator=g) * (0.5 if "embedding" in n else p.shape[1] ** -0.5))
    model._build_causal_attention_mask = types.MethodType(causal_mask_427, model)
    enc_state = {k: v.detach().clone() for k, v in model.encoder.state_dict().items()}
    model.encoder = _EncoderCompat(model.encoder)

    sig = PESigmas(sigma_t=0.03, sigma_l=2.0, sigma_theta=0.5, sigma_phi=0.5, sigma_r=0.5, sigma_dtu12=0.5)
    obj_id, view_ids = 100, [101, 102, 103]
    view_tokens = ["<view_0_10_1p2>", "<view_10_40_1p2>", "<view_20_70_1p2>"]
    cases = {}
    for name, unconstrained in (("bypass_unconstrained", True), ("bypass_matched_norm", False)):
        common = dict(output_dim=HIDDEN, arch_mlp_hidden_dims=64, arch_view_net=15, arch_view_disable_tl=False,
                      use_nested_dropout=False, pe_sigmas=sig, output_bypass=True, bypass_unconstrained=unconstrained,
                      output_bypass_alpha=0.2)
        torch.manual_seed(123)
        mo = NeTIMapper(embedding_type="object", norm_scale=torch.tensor(0.3714), placeholder_object_token="<statue>", **common)
        torch.manual_seed(321)
        mv = NeTIMapper(embedding_type="view", norm_scale=torch.tensor(0.4102), placeholder_view_tokens=list(view_tokens),
                        placeholder_view_token_ids=list(view_ids), **common)
        model.embeddings.set_mapper({obj_id: mo}, mv, device="cpu")
        B = 2
        input_ids = torch.randint(1, 90, (B, 77), generator=g)
        input_ids[:, 0] = 98                      # bos-like
        input_ids[0, 5], input_ids[1, 9] = obj_id, obj_id
        input_ids[0, 3], input_ids[1, 12] = view_ids[2], view_ids[0]
        timesteps = torch.tensor([17, 731])
        ph_obj = torch.tensor([obj_id, obj_id])
        ph_view = torch.tensor([view_ids[2], view_ids[0]])
        hs, hs_bypass, mo_out, mv_out = [], [], [], []
        params = [p for m in (mo, mv) for n, p in m.named_parameters() if n != "encoder.w"]
        for layer_idx, _ in enumerate(UNET_LAYERS):          # coach.py:289-305
            batch = NeTIBatch(input_ids=input_ids.clone(), input_ids_placeholder_object=ph_obj, input_ids_placeholder_view=ph_view,
                              timesteps=timesteps, unet_layers=torch.tensor(layer_idx).repeat(B))
            o, ob = model(batch=batch)
            hs.append(o[0])
            hs_bypass.append(ob[0])
            with torch.no_grad():                # the mapper outputs of this pass (inputs of the CPU oracle test)
                a = mo(timestep=timesteps.float(), unet_layer=batch.unet_layers.float(), input_ids_placeholder_view=None,
                       truncation_idx=None)
                c = mv(timestep=timesteps.float(), unet_layer=batch.unet_layers.float(), input_ids_placeholder_view=ph_view,
                       truncation_idx=None)
                mo_out.append(torch.stack([a.word_embedding, a.bypass_output]))
                mv_out.append(torch.stack([c.word_embedding, c.bypass_output]))
        hs, hs_bypass = torch.stack(hs), torch.stack(hs_bypass)            # [16, B, 77, C]
        gg = torch.Generator().manual_seed(100 + len(cases))       # the test regenerates these cotangents from the seed
        gk = torch.randn(hs.shape, generator=gg)
        gv = torch.randn(hs.shape, generator=gg)
        loss = (hs * gk).sum() + (hs_bypass * gv).sum()
        grads = torch.autograd.grad(loss, params)
        keep = list(range(16)) if not cases else [0, 7, 15]         # second case: three layers keep the fixture small
        names = [f"{mn}.{n}" for mn, m in (("object", mo), ("view", mv)) for n, p in m.named_parameters() if n != "encoder.w"]
        cases[name] = {
            "object_state": {k: v.detach().clone() for k, v in mo.state_dict().items()}, "object_w": mo.encoder.w.detach().clone(),
            "view_state": {k: v.detach().clone() for k, v in mv.state_dict().items()}, "view_w": mv.encoder.w.detach().clone(),
            "input_ids": input_ids, "timesteps": timesteps, "ph_obj": ph_obj, "ph_view": ph_view,
            "layers_kept": keep, "hs": hs.detach()[keep].half(), "hs_bypass": hs_bypass.detach()[keep].half(),
            "cotangent_seed": 100 + len(cases), "bypass_unconstrained": unconstrained, "output_bypass_alpha": 0.2,
            "mapper_object_out": torch.stack(mo_out), "mapper_view_out": torch.stack(mv_out),      # [16, 2 (word|bypass), B, C]
            "grads": dict(zip(names, [x.detach() for x in grads])),
        }
    out = {"config": dict(hidden=HIDDEN, heads=HEADS, layers=LAYERS, inter=INTER, vocab=VOCAB),
           "encoder_state": enc_state,
           "token_embedding": model.embeddings.token_embedding.weight.detach().clone(),
           "position_embedding": model.embeddings.position_embedding.weight.detach().clone(),
           "final_ln": (model.final_layer_norm.weight.detach().clone(), model.final_layer_norm.bias.detach().clone()),
           "obj_id": obj_id, "view_ids": view_ids, "view_tokens": view_tokens, "cases": cases}
    path = os.path.join(ROOT, "tests", "golden", "neti_conditioning.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
