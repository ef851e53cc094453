"""-m gpu: the CLIP text-transformer encoder of the conditioning path through the C-ABI against the
CPU oracle (oracle/clip_encoder.py, itself pinned to transformers' CLIPEncoder in tests/test_oracle_cpu.py)."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
BF = torch.bfloat16


def rel(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    assert torch.isfinite(a).all()
    return float((a - b).norm() / (b.norm() + 1e-12))


def test_gelu_fwd_bwd_matches_torch():
    from view_neti_b200 import ops
    g = torch.Generator().manual_seed(1)
    h = (2.0 * torch.randn(1232, 4096, generator=g)).cuda().to(BF)
    dy = torch.randn(1232, 4096, generator=g).cuda().to(BF)
    y, dh = torch.empty_like(h), torch.empty_like(h)
    ops.gelu_fwd(h, y, h.shape[0])
    ops.gelu_bwd(h, dy, dh, h.shape[0])
    hr = h.float().requires_grad_(True)
    yr = F.gelu(hr)
    yr.backward(dy.float())
    assert rel(y, yr) < 4e-3 and rel(dh, hr.grad) < 4e-3


@pytest.mark.parametrize("nseq,heads,L,causal", [(16, 16, 77, True), (3, 4, 77, False), (2, 2, 32, True), (1, 1, 128, True)])
def test_seq_attention_matches_torch(nseq, heads, L, causal):
    from view_neti_b200 import ops
    C = heads * 64
    g = torch.Generator().manual_seed(2)
    qkv = torch.randn(nseq, L, 3 * C, generator=g).cuda().to(BF)      # q / k / v as column slices of one buffer
    q, k, v = qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:]
    d_o = torch.randn(nseq, L, C, generator=g).cuda().to(BF)
    o = torch.empty(nseq, L, C, dtype=BF, device="cuda")
    lse = torch.empty(nseq, heads, L, device="cuda")
    dqkv = torch.zeros_like(qkv)
    ops.seq_attention_fwd(q, k, v, o, lse, heads, scale=0.125, causal=causal)
    ops.seq_attention_bwd(q, k, v, o, lse, d_o, dqkv[..., :C], dqkv[..., C:2 * C], dqkv[..., 2 * C:], heads, scale=0.125,
                          causal=causal)
    qr, kr, vr = (t.float().view(nseq, L, heads, 64).transpose(1, 2).detach().requires_grad_(True) for t in (q, k, v))
    s = (qr @ kr.transpose(-1, -2)) * 0.125
    if causal:
        s = s + torch.full((L, L), float("-inf"), device="cuda").triu(1)
    orf = torch.softmax(s, -1) @ vr
    orf.backward(d_o.float().view(nseq, L, heads, 64).transpose(1, 2))
    back = lambda t: t.transpose(1, 2).reshape(nseq, L, C)      # noqa: E731
    assert rel(o, back(orf)) < 5e-3
    assert rel(lse, torch.logsumexp(s, -1)) < 1e-4
    assert rel(dqkv[..., :C], back(qr.grad)) < 8e-3
    assert rel(dqkv[..., C:2 * C], back(kr.grad)) < 8e-3
    assert rel(dqkv[..., 2 * C:], back(vr.grad)) < 8e-3


@pytest.mark.parametrize("hidden,heads,layers,inter,nseq", [(1024, 16, 3, 4096, 16), (256, 4, 2, 1024, 5)])
def test_clip_encoder_matches_oracle(hidden, heads, layers, inter, nseq):
    """SD-2.1 text-encoder width (1024 x 16 heads, 4096 MLP) at 16 stacked per-UNet-layer passes of 77 tokens (B = 1),
    3 of the 23 layers so the fp32 CPU oracle stays quick: last_hidden_state and d(inputs_embeds)."""
    from oracle.clip_encoder import encoder_forward, init_state_dict
    from view_neti_b200.models.clip_encoder import CLIPEncoder, ClipEncoderConfig
    sd = init_state_dict(hidden, heads, layers, inter, seed=7)
    cfg = ClipEncoderConfig(hidden_size=hidden, num_attention_heads=heads, num_hidden_layers=layers, intermediate_size=inter)
    enc = CLIPEncoder(sd, cfg, "cuda")
    g = torch.Generator().manual_seed(8)
    x = torch.randn(nseq, 77, hidden, generator=g)
    dy = torch.randn(nseq, 77, hidden, generator=g)
    xo = x.clone().requires_grad_(True)
    yo = encoder_forward(sd, xo, heads, layers)
    yo.backward(dy)
    for rep in range(3):                       # eager, then two CUDA-graph replays
        xg = x.cuda().requires_grad_(True)
        out = enc(inputs_embeds=xg, attention_mask=None, causal_attention_mask=None)
        y = out[0]
        y.backward(dy.cuda())
        assert out.last_hidden_state is y
        e_y, e_dx = rel(y, yo), rel(xg.grad, xo.grad)
        assert e_y < 1e-2, (rep, e_y)
        assert e_dx < 2e-2, (rep, e_dx)
    # only the rows of a changed token and the rows after it move (causal mask), exactly as in the oracle test
    x2 = x.clone()
    x2[:, 50:] += 0.5
    y2 = enc(inputs_embeds=x2.cuda())[0]
    assert float((y2[:, :50] - y.detach()[:, :50]).abs().max()) == 0.0


def test_clip_encoder_rejects_unsupported_calls():
    from oracle.clip_encoder import init_state_dict
    from view_neti_b200._abi import VNError
    from view_neti_b200.models.clip_encoder import CLIPEncoder, ClipEncoderConfig
    cfg = ClipEncoderConfig(hidden_size=128, num_attention_heads=2, num_hidden_layers=1, intermediate_size=256)
    enc = CLIPEncoder(init_state_dict(128, 2, 1, 256), cfg, "cuda")
    with pytest.raises(VNError):
        enc(inputs_embeds=torch.zeros(1, 77, 128))                                  # CPU tensor: no fallback
    with pytest.raises(VNError):
        enc(inputs_embeds=torch.zeros(1, 77, 128, device="cuda"), attention_mask=torch.ones(1, 77, device="cuda"))


@pytest.mark.parametrize("case", ["bypass_unconstrained", "bypass_matched_norm"])
def test_batched_conditioning_matches_reference_golden(case):
    """The whole conditioning path (CUDA mappers -> embedding overwrite -> CUDA CLIP encoder -> bypass injection -> final
    LayerNorm), batched over the 16 UNet layers, against the outputs AND mapper-parameter gradients of the reference's own
    modules run layer by layer (tests/golden/make_golden_conditioning.py)."""
    import os
    from view_neti_b200.models.clip_encoder import CLIPEncoder, ClipEncoderConfig
    from view_neti_b200.models.neti_conditioning import NeTIConditioning
    from view_neti_b200.models.neti_mapper import NeTIMapper
    from view_neti_b200.utils.types import PESigmas
    G = torch.load(os.path.join(os.path.dirname(__file__), "golden", "neti_conditioning.pt"), weights_only=False)
    cfg, c = G["config"], G["cases"][case]
    enc = CLIPEncoder(G["encoder_state"], ClipEncoderConfig(hidden_size=cfg["hidden"], num_attention_heads=cfg["heads"],
                                                            num_hidden_layers=cfg["layers"], intermediate_size=cfg["inter"]), "cuda")
    sig = PESigmas(sigma_t=0.03, sigma_l=2.0, sigma_theta=0.5, sigma_phi=0.5, sigma_r=0.5, sigma_dtu12=0.5)
    kw = dict(output_dim=cfg["hidden"], arch_mlp_hidden_dims=64, arch_view_net=15, arch_view_disable_tl=False,
              use_nested_dropout=False, pe_sigmas=sig, output_bypass=True, bypass_unconstrained=c["bypass_unconstrained"],
              output_bypass_alpha=c["output_bypass_alpha"])
    mo = NeTIMapper(embedding_type="object", norm_scale=torch.tensor(0.3714), placeholder_object_token="<statue>", **kw)
    mv = NeTIMapper(embedding_type="view", norm_scale=torch.tensor(0.4102), placeholder_view_tokens=list(G["view_tokens"]),
                    placeholder_view_token_ids=list(G["view_ids"]), **kw)
    for m, key in ((mo, "object"), (mv, "view")):
        assert torch.equal(m.encoder_w, c[key + "_w"])
        m.load_state_dict({k: v for k, v in c[key + "_state"].items() if k != "encoder.w"}, strict=True)
        m.cuda()
    cond = NeTIConditioning(G["token_embedding"], G["position_embedding"], G["final_ln"], enc, {G["obj_id"]: mo}, mv)
    assert len(list(cond.parameters())) == 20          # both mappers are visible to the optimizer / the all-reduce
    hs = cond(input_ids=c["input_ids"], timesteps=c["timesteps"], input_ids_placeholder_object=c["ph_obj"],
              input_ids_placeholder_view=c["ph_view"])
    assert hs["this_idx"] == 0
    for j, layer in enumerate(c["layers_kept"]):
        assert rel(hs[f"CONTEXT_TENSOR_{layer}"], c["hs"][j]) < 1.5e-2, layer
        assert rel(hs[f"CONTEXT_TENSOR_BYPASS_{layer}"], c["hs_bypass"][j]) < 1.5e-2, layer
    if len(c["layers_kept"]) == 16:
        gg = torch.Generator().manual_seed(c["cotangent_seed"])
        shape = (16, *c["hs"].shape[1:])
        gk, gv = torch.randn(shape, generator=gg).cuda(), torch.randn(shape, generator=gg).cuda()
        loss = sum((hs[f"CONTEXT_TENSOR_{i}"] * gk[i]).sum() + (hs[f"CONTEXT_TENSOR_BYPASS_{i}"] * gv[i]).sum() for i in range(16))
        loss.backward()
        flat = torch.cat([p.grad.reshape(-1) for mname, m in (("object", mo), ("view", mv)) for n, p in m.named_parameters()])
        gold = torch.cat([c["grads"][f"{mname}.{n}"].reshape(-1) for mname, m in (("object", mo), ("view", mv))
                          for n, p in m.named_parameters()])
        assert rel(flat, gold) < 3e-2          # relative L2 of the flat mapper gradient (north_star: < 1e-2 at fp16 tolerance)


def test_complete_train_steps_through_coach():
    """Coach.train_step with the real conditioning stack (CUDA mappers -> batched CLIP encoder -> XTI dict) feeding the
    CUDA UNet (TINY widths, cross_attention_dim 128): gradients reach both mappers through the UNet's dgrad backward and
    the encoder's dgrad backward, AdamW moves them, and the loss on a fixed sample goes down (coach.py:165-218)."""
    from view_neti_b200.models.clip_encoder import ClipEncoderConfig
    from view_neti_b200.sd21 import TINY, init_state_dict
    from view_neti_b200.training.coach import Coach
    from view_neti_b200.training.synthetic import build_conditioning, synthetic_prompt
    from view_neti_b200.unet import UNet2DConditionModel
    cfg = ClipEncoderConfig(hidden_size=TINY.cross_attention_dim, num_attention_heads=2, num_hidden_layers=2, intermediate_size=256)
    cond = build_conditioning("cuda", cfg, seed=3)
    unet = UNet2DConditionModel(init_state_dict(TINY, 0), TINY, "cuda")
    coach = Coach(cfg=None, unet=unet, conditioning=cond, optimizer=torch.optim.AdamW(cond.parameters(), lr=2e-3),
                  generator=torch.Generator(device="cuda").manual_seed(5))
    batch = synthetic_prompt(2, "cuda")
    latents = torch.randn(2, 4, 16, 16, generator=torch.Generator().manual_seed(6)).cuda()
    before = [p.detach().clone() for p in cond.parameters()]

    def fixed_loss():
        with torch.no_grad():
            g = torch.Generator(device="cuda").manual_seed(9)
            noise = torch.randn(latents.shape, generator=g, device="cuda")
            t = torch.tensor([300, 700], device="cuda")
            hs = cond(timesteps=t, **batch)
            pred = unet(coach.noise_scheduler.add_noise(latents, noise, t), t, hs).sample
            target = noise if coach.noise_scheduler.config.prediction_type == "epsilon" else \
                coach.noise_scheduler.get_velocity(latents, noise, t)
            return float(F.mse_loss(pred.float(), target.float()))

    l0 = fixed_loss()
    losses = [float(coach.train_step(latents, batch)) for _ in range(30)]
    l1 = fixed_loss()
    assert all(math.isfinite(v) for v in losses)
    moved = [float((p.detach() - b).abs().max()) for p, b in zip(cond.parameters(), before)]
    assert len(moved) == 20 and min(moved) > 0, moved          # every tensor of both mappers received gradient
    assert l1 < l0, (l0, l1)


def test_prompt_manager_batches_timesteps_and_matches_reference_golden():
    """PromptManager.embed_prompt (reference prompt_manager.py:44-101): all timesteps of a prompt in batched inference
    passes.  The entry of timestep 17 for prompt 0 / 731 for prompt 1 must equal the reference's own outputs for those
    (prompt, timestep) pairs in tests/golden/neti_conditioning.pt; the other entries must equal the training-mode path."""
    import os
    from view_neti_b200.models.clip_encoder import CLIPEncoder, ClipEncoderConfig
    from view_neti_b200.models.neti_conditioning import NeTIConditioning
    from view_neti_b200.models.neti_mapper import NeTIMapper
    from view_neti_b200.prompt_manager import PromptManager
    from view_neti_b200.utils.types import PESigmas
    G = torch.load(os.path.join(os.path.dirname(__file__), "golden", "neti_conditioning.pt"), weights_only=False)
    cfg, c = G["config"], G["cases"]["bypass_unconstrained"]
    enc = CLIPEncoder(G["encoder_state"], ClipEncoderConfig(hidden_size=cfg["hidden"], num_attention_heads=cfg["heads"],
                                                            num_hidden_layers=cfg["layers"], intermediate_size=cfg["inter"]), "cuda")
    sig = PESigmas(sigma_t=0.03, sigma_l=2.0, sigma_theta=0.5, sigma_phi=0.5, sigma_r=0.5, sigma_dtu12=0.5)
    kw = dict(output_dim=cfg["hidden"], arch_mlp_hidden_dims=64, arch_view_net=15, arch_view_disable_tl=False,
              use_nested_dropout=False, pe_sigmas=sig, output_bypass=True, bypass_unconstrained=True, output_bypass_alpha=0.2)
    mo = NeTIMapper(embedding_type="object", norm_scale=torch.tensor(0.3714), placeholder_object_token="<statue>", **kw)
    mv = NeTIMapper(embedding_type="view", norm_scale=torch.tensor(0.4102), placeholder_view_tokens=list(G["view_tokens"]),
                    placeholder_view_token_ids=list(G["view_ids"]), **kw)
    for m, key in ((mo, "object"), (mv, "view")):
        m.load_state_dict({k: v for k, v in c[key + "_state"].items() if k != "encoder.w"}, strict=True)
        m.cuda()
    cond = NeTIConditioning(G["token_embedding"], G["position_embedding"], G["final_ln"], enc, {G["obj_id"]: mo}, mv)
    timesteps = [999, 17, 500, 731, 20]
    pm = PromptManager(tokenizer=None, text_encoder=cond, timesteps=timesteps, placeholder_view_token_ids=G["view_ids"],
                       placeholder_object_token_ids=[G["obj_id"]], chunk=2)
    for prompt, t_gold in ((0, 17), (1, 731)):
        embeds = pm.embed_prompt(c["input_ids"][prompt:prompt + 1], num_images_per_prompt=3)
        assert len(embeds) == len(timesteps) and embeds[0]["this_idx"] == 0
        e = embeds[timesteps.index(t_gold)]
        assert e["CONTEXT_TENSOR_0"].shape == (3, 77, cfg["hidden"])
        for layer in (0, 7, 15):
            assert rel(e[f"CONTEXT_TENSOR_{layer}"][0], c["hs"][layer][prompt]) < 1.5e-2
            assert rel(e[f"CONTEXT_TENSOR_BYPASS_{layer}"][2], c["hs_bypass"][layer][prompt]) < 1.5e-2
        # batched inference pass == training-mode path, timestep by timestep
        t_other = 500
        ref = cond(input_ids=c["input_ids"][prompt:prompt + 1].cuda(), timesteps=torch.tensor([t_other]).cuda(),
                   input_ids_placeholder_object=c["ph_obj"][prompt:prompt + 1], input_ids_placeholder_view=c["ph_view"][prompt:prompt + 1])
        got = embeds[timesteps.index(t_other)]
        assert rel(got["CONTEXT_TENSOR_9"][1], ref["CONTEXT_TENSOR_9"][0]) < 1e-5
        assert rel(got["CONTEXT_TENSOR_BYPASS_9"][1], ref["CONTEXT_TENSOR_BYPASS_9"][0]) < 1e-5


def test_object_only_conditioning_is_mode0_textual_inversion():
    """BASELINE config 0 (mode 0: object-only TI, no view mapper): prompts without a view token
    (input_ids_placeholder_view == -1, net_clip_text_embedding.py:105-106) still yield the 16 + 16 context tensors, only the
    object mapper receives gradient, and `original_ti` returns the first layer's plain tensor (coach.py:307-309)."""
    from view_neti_b200.models.clip_encoder import CLIPEncoder, ClipEncoderConfig, init_state_dict
    from view_neti_b200.models.neti_conditioning import NeTIConditioning
    from view_neti_b200.models.neti_mapper import NeTIMapper
    from view_neti_b200.utils.types import PESigmas
    cfg = ClipEncoderConfig(hidden_size=128, num_attention_heads=2, num_hidden_layers=2, intermediate_size=256)
    enc = CLIPEncoder(init_state_dict(cfg, 1), cfg, "cuda")
    g = torch.Generator().manual_seed(2)
    mo = NeTIMapper(embedding_type="object", output_dim=128, arch_mlp_hidden_dims=64, arch_view_net=15, arch_view_disable_tl=False,
                    use_nested_dropout=False, pe_sigmas=PESigmas(sigma_t=0.03, sigma_l=2.0), output_bypass=True,
                    bypass_unconstrained=True, output_bypass_alpha=0.2, norm_scale=torch.tensor(0.37),
                    placeholder_object_token="<teapot>").cuda()
    cond = NeTIConditioning(torch.randn(200, 128, generator=g) * 0.1, torch.randn(77, 128, generator=g) * 0.05,
                            (torch.ones(128), torch.zeros(128)), enc, {150: mo}, None)
    ids = torch.randint(1, 100, (2, 77), generator=g)
    ids[:, 4] = 150
    kw = dict(input_ids=ids.cuda(), timesteps=torch.tensor([10, 900]).cuda(), input_ids_placeholder_object=torch.tensor([150, 150]),
              input_ids_placeholder_view=torch.tensor([-1, -1]))
    hs = cond(**kw)
    assert sorted(k for k in hs if k != "this_idx") == sorted([f"CONTEXT_TENSOR_{i}" for i in range(16)] +
                                                              [f"CONTEXT_TENSOR_BYPASS_{i}" for i in range(16)])
    assert hs["CONTEXT_TENSOR_3"].shape == (2, 77, 128)
    sum((v ** 2).sum() for k, v in hs.items() if k != "this_idx").backward()
    assert all(p.grad is not None and float(p.grad.abs().max()) > 0 for p in mo.parameters())
    # rows before the placeholder do not depend on the mapper (causal mask): identical across the 16 layers
    a, b = hs["CONTEXT_TENSOR_0"].detach(), hs["CONTEXT_TENSOR_11"].detach()
    assert float((a[:, :4] - b[:, :4]).abs().max()) == 0.0
    assert float((a[:, 4:] - b[:, 4:]).abs().max()) > 0.0
    first = cond(original_ti=True, **kw)
    assert torch.is_tensor(first) and first.shape == (2, 77, 128)
    assert rel(first, hs["CONTEXT_TENSOR_0"]) < 1e-6


def _record(name, r):
    import json
    import os
    out = os.path.join(os.path.dirname(os.path.dirname(__file__)), "gpurun_out")
    try:
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, "parity_figures.jsonl"), "a") as f:
            f.write(json.dumps({"test": name, **r}) + "\n")
    except OSError:
        pass


def test_loss_to_mapper_gradient_parity_sd21_512px():
    """The north-star contract end to end at the headline shape: MSE loss -> SD-2.1-width UNet dgrad (64x64 latents, B = 1)
    -> FULL 23-layer CLIP-H-width text-encoder dgrad (16 stacked passes) -> the 283 392 parameters of M_o and M_v, against
    the fp32 CPU oracles on identical inputs and weights (coach.py:186-214, 276-311).  eps-MSE < 1e-3 and relative L2 error
    of the flat mapper gradient < 1e-2 (BASELINE.json north_star)."""
    from tests.e2e_parity import run_e2e
    from view_neti_b200.sd21 import SD21
    r = run_e2e(SD21, 23, 16, 4096, 1, 64, 64)
    _record("e2e_mapper_grad_sd21_512px", r)
    assert r["eps_mse"] < 1e-3 and r["loss_rel"] < 2e-3, (r["eps_mse"], r["loss_rel"])
    assert r["mapper_grad_flat_rel"] < 1e-2, r


def test_loss_to_mapper_gradient_parity_tiny_matched_norm_bypass():
    """Same chain on the narrow net, batch 2, with the norm-matched bypass mode (neti_clip_text_encoder.py:139-142)."""
    from tests.e2e_parity import run_e2e
    from view_neti_b200.sd21 import TINY
    r = run_e2e(TINY, 2, 2, 256, 2, 16, 16, bypass_unconstrained=False)
    _record("e2e_mapper_grad_tiny", r)
    assert r["eps_mse"] < 1e-3 and r["mapper_grad_flat_rel"] < 3e-2, r       # narrow random-weight net: noisier gradients


@pytest.mark.parametrize("mode,cam", [(2, "dtu-12d"), (3, "spherical"), (0, "spherical")])
def test_coach_from_runconfig_trains(mode, cam, tmp_path):
    """`Coach(cfg).train()` as reference scripts/train.py:23-24 runs it: every component built from the RunConfig (seeded
    narrow models, synthetic data in the reference's item format), micro-batch 2 x accumulation 2, checkpoints written in
    the reference's layout and read back."""
    from view_neti_b200.checkpoint_handler import CheckpointHandler
    from view_neti_b200.training.coach import Coach
    from view_neti_b200.training.config import from_dict
    objs = [f"<o{i}>" for i in range(3)]
    cfg = from_dict({
        "learnable_mode": mode, "seed": 3,
        "log": {"exp_dir": str(tmp_path), "save_steps": 2},
        "data": {"train_data_dir": "synthetic", "placeholder_object_token": "<statue>", "dataloader_num_workers": 0,
                 "camera_representation": cam, "resolution": 128, "repeats": 16,
                 **({"placeholder_object_tokens": objs, "super_category_object_tokens": ["object"] * 3,
                     "train_data_subsets": ["a", "b", "c"]} if mode == 3 else {})},
        "model": {"pretrained_model_name_or_path": "synthetic-tiny", "word_embedding_dim": 128, "arch_mlp_hidden_dims": 64,
                  "use_nested_dropout": False, "arch_view_net": 15, "arch_view_disable_tl": False,
                  "normalize_view_mapper_output": True, "bypass_unconstrained_object": True, "bypass_unconstrained_view": True},
        "optim": {"max_train_steps": 3, "train_batch_size": 2, "gradient_accumulation_steps": 2, "learning_rate": 1e-3,
                  "scale_lr": True, "seed": 1},
    })
    coach = Coach(cfg)
    n_obj = 3 if mode == 3 else 1
    assert len(coach.text_encoder.text_model.embeddings.mapper_object_lookup) == n_obj
    assert (coach.text_encoder.text_model.embeddings.mapper_view is None) == (mode == 0)
    assert coach.optimizer.defaults["lr"] == pytest.approx(1e-3 * 2 * 2)                 # scale_lr (coach.py:728-733)
    before = [p.detach().clone() for p in coach.conditioning.parameters()]
    losses = coach.train()
    assert coach.global_step == 3 and coach.micro_step == 6 and len(losses) == 6
    assert all(math.isfinite(float(l)) for l in losses)
    moved = [not torch.equal(p.detach(), b) for p, b in zip(coach.conditioning.parameters(), before)]
    assert any(moved)
    import os
    files = sorted(os.listdir(tmp_path))
    assert "mapper-steps-2_object.pt" in files and "mapper-final_object.pt" in files and "learned_embeds-final.bin" in files
    assert ("mapper-final_view.pt" in files) == (mode != 0)
    _, loaded = CheckpointHandler.load_mapper(tmp_path / "mapper-final_object.pt", "object",
                                              placeholder_object_tokens=coach.placeholder_object_tokens,
                                              placeholder_object_token_ids=coach.placeholder_object_token_ids)
    assert sorted(loaded) == sorted(coach.placeholder_object_token_ids)
    tid = coach.placeholder_object_token_ids[0]
    ours = coach.text_encoder.text_model.embeddings.mapper_object_lookup[tid]
    assert torch.equal(loaded[tid].output_layer[0].weight.cpu(), ours.output_layer[0].weight.detach().cpu())
