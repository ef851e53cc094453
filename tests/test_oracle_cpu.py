"""-m "not gpu": the oracle against the golden vectors produced by the reference's own code
(tests/golden/make_golden.py executes /root/reference/models/xti_attention_processor.py), plus structural pins."""
import os

import pytest
import torch

from oracle.unet_sd21 import CrossAttention, UNetOracle, XTIAttenProcOracle, train_step_oracle
from tests.unet_parity import ctx_to, make_inputs
from view_neti_b200.sd21 import SD21, TINY, cross_attn_layer_names, init_state_dict, num_params, param_table

GOLD = os.path.join(os.path.dirname(__file__), "golden", "xti_attn.pt")


@pytest.fixture(scope="module")
def gold():
    return torch.load(GOLD)


def _attn(state, qdim, cdim, heads):
    m = CrossAttention(qdim, cdim, heads)
    m.load_state_dict(state)
    m.processor = XTIAttenProcOracle()
    return m


def test_oracle_processor_matches_reference_outputs(gold):
    cross = _attn(gold["cross_state"], 128, 192, gold["heads"])
    selfa = _attn(gold["self_state"], 128, None, gold["heads"])
    h = gold["hidden"]
    with torch.no_grad():
        d = dict(gold["ctx"])
        y = cross(h, encoder_hidden_states=d)
        assert torch.equal(y, gold["dict_bypass"])
        assert d["this_idx"] == gold["dict_bypass_this_idx_after"] == 4
        d = {k: v for k, v in gold["ctx"].items() if "BYPASS" not in k}
        d["this_idx"] = 15
        assert torch.equal(cross(h, encoder_hidden_states=d), gold["dict_nobypass_idx15"])
        assert d["this_idx"] == gold["dict_nobypass_this_idx_after"] == 0          # wraps modulo 16
        assert torch.equal(cross(h, encoder_hidden_states=gold["ctx"]["CONTEXT_TENSOR_5"]), gold["tensor_ctx"])
        assert torch.equal(selfa(h, encoder_hidden_states=None), gold["self"])


def test_oracle_processor_matches_reference_gradients(gold):
    cross = _attn(gold["cross_state"], 128, 192, gold["heads"])
    d = {k: (v.clone().requires_grad_(True) if torch.is_tensor(v) else v) for k, v in gold["ctx"].items()}
    h = gold["hidden"].clone().requires_grad_(True)
    y = cross(h, encoder_hidden_states=d)
    (y * gold["grad_w"]).sum().backward()
    assert torch.allclose(d["CONTEXT_TENSOR_3"].grad, gold["grad_ctx_k"], rtol=0, atol=1e-6)
    assert torch.allclose(d["CONTEXT_TENSOR_BYPASS_3"].grad, gold["grad_ctx_v"], rtol=0, atol=1e-6)
    assert torch.allclose(h.grad, gold["grad_hidden"], rtol=0, atol=1e-6)


def test_same_k_and_v_context_is_plain_cross_attention(gold):
    """XTI with CONTEXT_TENSOR_BYPASS_i == CONTEXT_TENSOR_i is vanilla cross-attention (SURVEY 8c invariant i)."""
    cross = _attn(gold["cross_state"], 128, 192, gold["heads"])
    c = gold["ctx"]["CONTEXT_TENSOR_3"]
    with torch.no_grad():
        a = cross(gold["hidden"], encoder_hidden_states={"this_idx": 3, "CONTEXT_TENSOR_3": c, "CONTEXT_TENSOR_BYPASS_3": c})
        b = cross(gold["hidden"], encoder_hidden_states=c)
    assert torch.equal(a, b)


def test_sd21_topology_pins():
    assert num_params(SD21) == 865_910_724                       # SD-2.1 UNet (diffusers reports 865.91 M)
    names = cross_attn_layer_names(SD21)
    assert len(names) == 16 == SD21.num_cross_layers
    assert names[0] == "down_blocks.0.attentions.0" and names[6] == "mid_block.attentions.0" \
        and names[-1] == "up_blocks.3.attentions.2"
    keys = {k for k, _, _ in param_table(SD21)}
    for k in ("conv_in.weight", "time_embedding.linear_2.bias", "down_blocks.1.attentions.0.transformer_blocks.0.attn2.to_k.weight",
              "up_blocks.1.upsamplers.0.conv.weight", "up_blocks.3.resnets.2.conv_shortcut.weight", "conv_norm_out.bias"):
        assert k in keys
    shp = {k: s for k, s, _ in param_table(SD21)}
    assert shp["down_blocks.0.attentions.0.transformer_blocks.0.attn2.to_k.weight"] == (320, 1024)
    assert shp["up_blocks.0.resnets.0.conv1.weight"] == (1280, 2560, 3, 3)
    assert shp["up_blocks.3.resnets.0.conv1.weight"] == (320, 960, 3, 3)


def test_oracle_unet_state_dict_keys_and_this_idx_roundtrip():
    sd = init_state_dict(TINY, 0)
    unet = UNetOracle(TINY)
    missing, unexpected = unet.load_state_dict(sd, strict=True)
    assert not missing and not unexpected
    lat, t, tgt, ctx = make_inputs(TINY, 1, 8, 8)
    c = ctx_to(ctx, "cpu")
    c["this_idx"] = 5
    eps, loss, grads = train_step_oracle(unet, lat, t, tgt, c)
    assert c["this_idx"] == 5                                     # 16 increments modulo 16 (SURVEY 8c invariant iii)
    assert eps.shape == lat.shape and torch.isfinite(eps).all() and all(g is not None for g in grads)


def test_plain_tensor_context_equals_dict_of_identical_entries():
    """SURVEY 8c invariant ii."""
    sd = init_state_dict(TINY, 0)
    unet = UNetOracle(TINY)
    unet.load_state_dict(sd)
    lat, t, _, ctx = make_inputs(TINY, 1, 8, 8)
    c0 = ctx["CONTEXT_TENSOR_0"]
    with torch.no_grad():
        a = unet(lat, t, c0).sample
        b = unet(lat, t, {"this_idx": 0, **{f"CONTEXT_TENSOR_{i}": c0 for i in range(16)}}).sample
    assert torch.equal(a, b)


def test_context_gradient_matches_finite_differences():
    """SURVEY 8c invariant iv: d loss / d ctx of the oracle agrees with central differences (float64)."""
    torch.manual_seed(0)
    sd = {k: v.double() for k, v in init_state_dict(TINY, 0).items()}
    unet = UNetOracle(TINY).double()
    unet.load_state_dict(sd)
    for m in unet.modules():                      # `.float()` logits would inject fp32 noise into the differences
        if isinstance(m, CrossAttention):
            m.upcast_attention = False
    lat, t, tgt, ctx = make_inputs(TINY, 1, 8, 8)
    c = {k: (v.double().requires_grad_(True) if torch.is_tensor(v) else v) for k, v in ctx.items()}
    eps, loss, grads = train_step_oracle(unet, lat.double(), t, tgt.double(), c)
    keys = [k for k, v in c.items() if torch.is_tensor(v)]
    g = dict(zip(keys, grads))
    for key, idx in (("CONTEXT_TENSOR_7", (0, 3, 5)), ("CONTEXT_TENSOR_BYPASS_0", (0, 10, 100))):
        h = 1e-4
        vals = []
        for sgn in (+1, -1):
            c2 = {k: (v.detach().clone() if torch.is_tensor(v) else v) for k, v in c.items()}
            c2[key][idx] += sgn * h
            with torch.no_grad():
                e = unet(lat.double(), t, c2).sample
            vals.append(torch.nn.functional.mse_loss(e, tgt.double()).item())
        fd = (vals[0] - vals[1]) / (2 * h)
        assert abs(fd - g[key][idx].item()) <= 1e-6 + 1e-3 * abs(fd), (key, fd, g[key][idx].item())


def test_clip_encoder_oracle_matches_transformers_clipencoder():
    """The conditioning path's arithmetic lives in transformers' CLIPEncoder (reference neti_clip_text_encoder.py:53);
    the restatement in oracle/clip_encoder.py is pinned to that class itself: outputs and input gradients agree to fp32
    round-off, and the causal mask makes position i independent of later tokens."""
    import torch
    from oracle.clip_encoder import encoder_forward, hf_encoder, init_state_dict
    hidden, heads, layers, inter = 128, 2, 3, 512
    sd = init_state_dict(hidden, heads, layers, inter, seed=3)
    run = hf_encoder(sd, hidden, heads, layers, inter)
    g = torch.Generator().manual_seed(4)
    x = torch.randn(3, 77, hidden, generator=g)
    dy = torch.randn(3, 77, hidden, generator=g)
    x1 = x.clone().requires_grad_(True)
    y1 = run(x1)
    y1.backward(dy)
    x2 = x.clone().requires_grad_(True)
    y2 = encoder_forward(sd, x2, heads, layers)
    y2.backward(dy)
    assert float((y1 - y2).detach().abs().max()) < 2e-5 * float(y1.detach().abs().max())
    assert float((x1.grad - x2.grad).abs().max()) < 2e-5 * float(x1.grad.abs().max())
    xp = x.clone()
    xp[:, 40:] += 1.0
    yp = encoder_forward(sd, xp, heads, layers)
    assert float((yp[:, :40] - y2[:, :40]).abs().max()) == 0.0 and float((yp[:, 40:] - y2[:, 40:]).abs().max()) > 1e-3


def test_conditioning_oracle_matches_reference_golden():
    """tests/golden/neti_conditioning.pt = outputs of the reference's OWN conditioning path (its text transformer, text
    embeddings and mappers imported unmodified, one pass per UNet layer as in coach.py:276-311).  The batched restatement
    (oracle/neti_conditioning.py + oracle/clip_encoder.py) must reproduce them for both bypass modes."""
    import os
    import torch
    from oracle.clip_encoder import encoder_forward
    from oracle.neti_conditioning import conditioning_forward
    G = torch.load(os.path.join(os.path.dirname(__file__), "golden", "neti_conditioning.pt"), weights_only=False)
    cfg = G["config"]
    enc = lambda x: encoder_forward(G["encoder_state"], x, cfg["heads"], cfg["layers"])      # noqa: E731
    for name, c in G["cases"].items():
        hs = conditioning_forward(c["input_ids"], c["ph_obj"], c["ph_view"], G["token_embedding"], G["position_embedding"],
                                  G["final_ln"], enc, c["mapper_object_out"], c["mapper_view_out"],
                                  c["bypass_unconstrained"], c["output_bypass_alpha"])
        for j, layer in enumerate(c["layers_kept"]):
            for key, gold in ((f"CONTEXT_TENSOR_{layer}", c["hs"][j]), (f"CONTEXT_TENSOR_BYPASS_{layer}", c["hs_bypass"][j])):
                err = float((hs[key] - gold.float()).abs().max())
                assert err < 4e-3, (name, key, err)          # fixture is stored in fp16 (|values| ~ 1-4)


@pytest.mark.parametrize("kind", ["object", "view"])
def test_mapper_oracle_matches_reference_golden(kind):
    """oracle/neti_mapper.py against outputs and parameter gradients of the reference's own NeTIMapper
    (tests/golden/neti_mapper.pt, generated by tests/golden/make_golden_mapper.py)."""
    from oracle.neti_mapper import encode_inputs, mapper_forward
    G = torch.load(os.path.join(os.path.dirname(__file__), "golden", "neti_mapper.pt"))
    g = G[kind]
    state = {k: v.clone().requires_grad_(True) for k, v in g["state"].items() if k != "encoder.w"}
    vp = vmin = vmax = None
    if kind == "view":
        table = {i: [float(s.replace("p", ".")) for s in tok[6:-1].split("_")] for tok, i in zip(g["tokens"], g["ids"])}
        allp = torch.tensor(list(table.values()))
        vp = torch.tensor([table[int(i)][:2] for i in g["input_ids"]])                  # (theta, phi) degrees of freedom
        vmin, vmax = allp.min(0).values.tolist(), allp.max(0).values.tolist()
    x = encode_inputs(G["t"], G["l"], vp, vmin, vmax)
    word, bypass = mapper_forward(state, g["w"], x, g["norm_scale"])
    assert torch.allclose(word, g["word"], rtol=0, atol=2e-6) and torch.allclose(bypass, g["bypass"], rtol=0, atol=2e-6)
    ((word * g["gw"]).sum() + (bypass * g["gb"]).sum()).backward()
    for k, v in state.items():
        assert torch.allclose(v.grad, g["grads"][k], rtol=1e-4, atol=1e-5), k


def test_end_to_end_oracle_chain_runs_on_cpu():
    """Oracle half of tests/e2e_parity.py (mappers -> conditioning -> CLIP -> UNet -> MSE -> mapper gradients) on the narrow
    net: finite, non-zero gradients for all 2 x 10 mapper tensors (the GPU comparison is tests/test_clip_gpu.py)."""
    from tests.e2e_parity import run_e2e
    if torch.cuda.is_available():
        pytest.skip("dry run of the oracle half is for CPU-only boxes")
    r = run_e2e(TINY, 2, 2, 256, 2, 16, 16)
    assert r["n"] == 2 * (64 * 64 * 2 + 64 * 6 + 2 * TINY.cross_attention_dim * 65) and r["mapper_grad_norm"] > 0
