"""-m "not gpu": host logic of the training driver and the text-model shell - config loading, tokenizer surgery,
gradient accumulation, mode 3 (several object mappers, one active per step, same object on every rank), the 2-tuple
call shapes of NeTICLIPTextModel.  The models are CPU stubs: nothing here computes on the product's CUDA path."""
import os
from types import SimpleNamespace

import pytest
import torch

from view_neti_b200.training.coach import Coach
from view_neti_b200.training.config import RunConfig, from_dict, load_config, parse_args, to_dict
from view_neti_b200.training.synthetic import SyntheticTIDataset, SyntheticTokenizer
from view_neti_b200.utils.types import NeTIBatch, PESigmas

YAML = os.path.join(os.path.dirname(__file__), "golden", "train_m3_synthetic.yaml")


# ---- config ---------------------------------------------------------------------------------------------------------
def test_config_loads_yaml_with_dotted_overrides():
    cfg = load_config(YAML, ["--optim.max_train_steps", "7", "--log.overwrite_ok", "--model.arch_view_net=15"])
    assert isinstance(cfg, RunConfig) and cfg.learnable_mode == 3
    assert cfg.optim.max_train_steps == 7 and cfg.log.overwrite_ok is True and cfg.optim.train_batch_size == 3
    assert cfg.optim.gradient_accumulation_steps == 3                                      # the reference's 3 x 3 default
    # config.py:142-178: dict -> PESigmas, theta / r filled from the phi key, experiment keys applied
    assert cfg.model.pe_sigmas == PESigmas(sigma_t=0.03, sigma_l=2.0, sigma_theta=2.0, sigma_phi=2.0, sigma_r=2.0, sigma_dtu12=0.5)
    assert len(cfg.data.placeholder_object_tokens) == 14
    d = to_dict(cfg)
    assert d["model"]["word_embedding_dim"] == 1024 and d["log"]["exp_dir"] == "results"
    assert parse_args(["--config_path", YAML, "--seed", "5"]).seed == 5


def test_config_validation_follows_the_reference():
    with pytest.raises(ValueError):                              # config.py:269-271
        from_dict({"optim": {"train_batch_size": 4}})
    with pytest.raises(AssertionError):                          # mode 3 needs the super-category list
        from_dict({"learnable_mode": 3, "data": {"dataloader_num_workers": 0}})
    with pytest.raises(AssertionError):
        from_dict({"data": {"placeholder_object_tokens": ["<a>", "<a>"]}})
    with pytest.raises(ValueError):
        from_dict({"optim": {"no_such_field": 1}})
    with pytest.raises(AssertionError):                          # modes 4 / 5 need a pretrained view mapper
        from_dict({"learnable_mode": 5})


# ---- CPU stand-ins --------------------------------------------------------------------------------------------------
class _IdentityEncoder:
    engine = SimpleNamespace(dev=torch.device("cpu"))

    def __call__(self, inputs_embeds=None, **_):
        return (inputs_embeds * 1.0,)


def _text_model(vocab=49408, C=16):
    from view_neti_b200.models.neti_clip_text_encoder import NeTICLIPTextModel
    g = torch.Generator().manual_seed(0)
    return NeTICLIPTextModel.from_parts(torch.randn(vocab, C, generator=g), torch.randn(77, C, generator=g) * 0.1,
                                        (torch.ones(C), torch.zeros(C)), _IdentityEncoder())


class _StubMapper(torch.nn.Module):
    """Host-side stand-in for a NeTIMapper (the real one runs on the CUDA library only)."""

    def __init__(self, C, seed):
        super().__init__()
        g = torch.Generator().manual_seed(seed)
        self.w = torch.nn.Parameter(torch.randn(2, C, generator=g))

    def forward(self, timestep, unet_layer, input_ids_placeholder_view, truncation_idx=None):
        from view_neti_b200.utils.types import MapperOutput
        s = (1 + timestep / 1000 + unet_layer / 16).unsqueeze(1)
        return MapperOutput(word_embedding=s * self.w[0], bypass_output=s * self.w[1], bypass_unconstrained=False,
                            output_bypass_alpha=0.2)


class _StubUNet(torch.nn.Module):
    """eps = mean over the 32 context tensors, broadcast to the latent shape: linear in the contexts, frozen."""
    device = torch.device("cpu")

    def forward(self, sample, timestep, ehs):
        ctx = [v for k, v in ehs.items() if k != "this_idx"]
        s = torch.stack([c.mean(dim=(1, 2)) for c in ctx]).mean(0)                         # [B]
        return SimpleNamespace(sample=sample * 0.1 + s.view(-1, 1, 1, 1))


# ---- the text-model shell -------------------------------------------------------------------------------------------
def test_text_model_answers_the_references_call_shapes():
    tm = _text_model()
    tok = SyntheticTokenizer()
    # sd_pipeline_call.py:36-41: negative prompt, plain CLIP text model, ALWAYS a 2-tuple
    neg = tok("", padding="max_length", max_length=tok.model_max_length, truncation=True, return_tensors="pt")
    negative_prompt_embeds, _ = tm(input_ids=neg.input_ids, attention_mask=None)
    assert _ is None and negative_prompt_embeds[0].shape == (1, 77, 16)
    assert negative_prompt_embeds.last_hidden_state is negative_prompt_embeds[0]
    assert negative_prompt_embeds.pooler_output.shape == (1, 16)
    # coach.py:289-305: one UNet layer per call through batch=NeTIBatch, mappers installed with set_mapper
    tok.add_tokens(["<view_0_0_1p2>", "<obj>"])
    tm.resize_token_embeddings(len(tok))
    vid, oid = tok.convert_tokens_to_ids(["<view_0_0_1p2>", "<obj>"])
    tm.text_model.embeddings.set_mapper({oid: _StubMapper(16, 1)}, _StubMapper(16, 2), device="cpu")
    assert set(tm.text_model.embeddings.mapper_object_lookup) == {oid} and tm.text_model.embeddings.mapper_view is not None
    ids = tok(["<view_0_0_1p2> . A photo of a <obj>"] * 2, return_tensors="pt").input_ids
    b = NeTIBatch(input_ids=ids, input_ids_placeholder_object=torch.tensor([oid, oid]),
                  input_ids_placeholder_view=torch.tensor([vid, vid]), timesteps=torch.tensor([10, 900]),
                  unet_layers=torch.tensor([5, 5]))
    out, out_bypass = tm(batch=b)
    assert out[0].shape == (2, 77, 16) and out_bypass is not None and out_bypass[0].shape == (2, 77, 16)
    pos = int((ids[0] == oid).nonzero()[0])
    assert not torch.allclose(out[0][:, pos], out_bypass[0][:, pos])                       # bypass injected at the placeholder
    other = [i for i in range(77) if i not in (pos, int((ids[0] == vid).nonzero()[0]))]
    assert torch.allclose(out[0][:, other], out_bypass[0][:, other])                        # ... and nowhere else
    # the stacked path used by Coach.get_text_conditioning equals the per-layer calls
    hs = tm.conditioning(input_ids=ids, timesteps=b.timesteps, input_ids_placeholder_object=b.input_ids_placeholder_object,
                         input_ids_placeholder_view=b.input_ids_placeholder_view)
    assert torch.allclose(hs["CONTEXT_TENSOR_5"], out[0]) and torch.allclose(hs["CONTEXT_TENSOR_BYPASS_5"], out_bypass[0])
    with pytest.raises(ValueError):
        tm()
    tm.text_model.encoder.requires_grad_(False)                                            # coach.py:650-653
    tm.text_model.final_layer_norm.requires_grad_(False)
    tm.text_model.embeddings.position_embedding.requires_grad_(False)


def test_add_concept_token_to_tokenizer_static():
    cfg = load_config(YAML)
    tok, tm = SyntheticTokenizer(), _text_model()
    views, objs = ["<view_0_0_1p2>", "<view_0_90_1p2>"], ["<skull>", "<toy>"]
    emb, all_ids, view_ids, obj_ids = Coach._add_concept_token_to_tokenizer_static(cfg, views, objs, tok, tm)
    assert len(tok) == 49408 + 4 and emb.shape[0] == len(tok) and all_ids == view_ids + obj_ids
    sup_v = tok.encode(cfg.data.super_category_view_token, add_special_tokens=False)[0]
    sup_o = tok.encode(cfg.data.super_category_object_token, add_special_tokens=False)[0]
    assert torch.equal(emb[view_ids[1]], emb[sup_v]) and torch.equal(emb[obj_ids[0]], emb[sup_o])
    assert cfg.model.target_norm_view == pytest.approx(float(emb[sup_v].norm()))
    assert cfg.model.target_norm_object == pytest.approx(float(emb[sup_o].norm()))
    with pytest.raises(ValueError):                              # nothing new to add
        Coach._add_concept_token_to_tokenizer_static(cfg, views, objs, tok, tm)


# ---- Coach: accumulation and mode 3 ---------------------------------------------------------------------------------
class _StubConditioning(torch.nn.Module):
    """mapper_object_lookup (ModuleDict) + mapper_view, emitting the XTI dict from the ACTIVE object mapper only."""

    def __init__(self, object_ids, C=8):
        super().__init__()
        self.mapper_object_lookup = torch.nn.ModuleDict({str(i): _StubMapper(C, 10 + k) for k, i in enumerate(object_ids)})
        self.mapper_view = _StubMapper(C, 3)

    def forward(self, input_ids=None, timesteps=None, input_ids_placeholder_object=None, input_ids_placeholder_view=None, **_):
        ph = [int(v) for v in input_ids_placeholder_object]
        assert all(v == ph[0] for v in ph), "one object per batch"
        B = timesteps.shape[0]
        out = {"this_idx": 0}
        for l in range(16):
            lay = torch.full((B,), float(l))
            o = self.mapper_object_lookup[str(ph[0])](timesteps.float(), lay, None)
            v = self.mapper_view(timesteps.float(), lay, None)
            out[f"CONTEXT_TENSOR_{l}"] = (o.word_embedding + v.word_embedding).unsqueeze(1).expand(B, 77, -1)
            out[f"CONTEXT_TENSOR_BYPASS_{l}"] = (o.bypass_output + v.bypass_output).unsqueeze(1).expand(B, 77, -1)
        return out


def _cfg(accum=1, mode=2, lr=0.05):
    return SimpleNamespace(learnable_mode=mode, model=SimpleNamespace(original_ti=False),
                           optim=SimpleNamespace(learning_rate=lr, gradient_accumulation_steps=accum, scale_lr=False,
                                                 adam_beta1=0.8, adam_beta2=0.9, adam_weight_decay=0.0, adam_epsilon=1e-6,
                                                 max_train_steps=None, train_batch_size=1))


def test_gradient_accumulation_window_matches_one_big_batch():
    """coach.py:158 / accelerate.accumulate: k micro-steps, loss / k, ONE optimizer step - the accumulated gradient is the
    mean over the window, parameters move only on the window's last micro-step."""
    ids = [49408]
    batch = {"input_ids": None, "input_ids_placeholder_object": torch.tensor([49408]), "input_ids_placeholder_view": torch.tensor([-1])}

    def run(accum, n_micro):
        torch.manual_seed(0)
        cond = _StubConditioning(ids)
        coach = Coach(_cfg(accum), unet=_StubUNet(), conditioning=cond, generator=torch.Generator().manual_seed(4))
        assert coach.optimizer.defaults["betas"] == (0.8, 0.9) and coach.optimizer.defaults["eps"] == 1e-6   # cfg.optim is honoured
        snaps, lat = [], torch.ones(1, 4, 4, 4)
        for _ in range(n_micro):
            coach.train_step(lat, batch)
            snaps.append(torch.cat([p.detach().reshape(-1).clone() for p in cond.parameters()]))
        return coach, snaps

    coach, snaps = run(3, 6)
    assert coach.global_step == 2 and coach.micro_step == 6
    p0 = torch.cat([p.detach().reshape(-1) for p in _StubConditioning(ids).parameters()])
    assert torch.equal(snaps[0], p0) and torch.equal(snaps[1], p0) and not torch.equal(snaps[2], p0)   # moved on micro-step 3 only
    assert torch.equal(snaps[3], snaps[2]) and torch.equal(snaps[4], snaps[2]) and not torch.equal(snaps[5], snaps[2])


def test_inactive_object_mappers_stay_bit_identical():
    """Mode 3, one process: the step's object mapper and M_v move, the other object mappers keep .grad None and do not
    change by a single bit (no weight decay, no stale momentum) - reference coach.py:736-757 after zero_grad."""
    ids = [49408, 49409, 49410]
    cond = _StubConditioning(ids)
    cfg = _cfg(1, mode=3)
    cfg.optim.adam_weight_decay = 0.5
    coach = Coach(cfg, unet=_StubUNet(), conditioning=cond, generator=torch.Generator().manual_seed(1))
    before = {k: m.w.detach().clone() for k, m in cond.mapper_object_lookup.items()}
    v0 = cond.mapper_view.w.detach().clone()
    for active in (49409, 49409, 49410):
        coach.train_step(torch.ones(2, 4, 4, 4), {"input_ids_placeholder_object": torch.tensor([active, active]),
                                                  "input_ids_placeholder_view": torch.tensor([-1, -1])})
    assert torch.equal(cond.mapper_object_lookup["49408"].w, before["49408"])
    assert not torch.equal(cond.mapper_object_lookup["49409"].w, before["49409"])
    assert not torch.equal(cond.mapper_view.w, v0)


def _mode3_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    tok = SyntheticTokenizer()
    ds = SyntheticTIDataset(3, tok, placeholder_object_tokens=[f"<o{i}>" for i in range(14)], size=(16, 16), seed=100 + rank)
    tok.add_tokens(ds.placeholder_tokens)
    obj_ids = tok.convert_tokens_to_ids(ds.placeholder_object_tokens)
    torch.manual_seed(rank)                                       # different local init: the constructor must broadcast rank 0's
    cond = _StubConditioning(obj_ids)
    for p in cond.parameters():
        p.data.add_(float(rank))
    coach = Coach(_cfg(1, mode=3), unet=_StubUNet(), conditioning=cond, generator=torch.Generator().manual_seed(7 + rank),
                  train_dataset=ds)
    start = {k: m.w.detach().clone() for k, m in cond.mapper_object_lookup.items()}
    drawn = []
    for step in range(6):
        idx = coach.reset_sampled_object()                        # rank 0 draws, everybody follows (coach.py:155-156 + SURVEY 8e)
        drawn.append(idx)
        item = ds[step + rank]
        batch = {"input_ids": item["input_ids"][None], "input_ids_placeholder_object": item["input_ids_placeholder_object"][None],
                 "input_ids_placeholder_view": torch.tensor([-1])}
        coach.train_step(torch.full((1, 4, 4, 4), 1.0 + rank), batch)          # different data per rank
    used = {str(obj_ids[i]) for i in drawn}
    unchanged = all(torch.equal(m.w, start[k]) for k, m in cond.mapper_object_lookup.items() if k not in used)
    moved = all(not torch.equal(cond.mapper_object_lookup[k].w, start[k]) for k in used)
    flat = torch.cat([p.detach().reshape(-1) for p in cond.parameters()])
    q.put((rank, drawn, unchanged, moved, flat.tolist()))        # plain lists: a tensor in the queue shares memory with a
    #                                                              process that may be gone by the time the parent reads it
    dist.destroy_process_group()


def test_mode3_same_object_on_all_ranks_gloo_world2():
    """BASELINE config 4 host logic: 14 object mappers resident, ONE active per step and the SAME one on every rank (drawn on
    rank 0, broadcast), all-reduce of M_v + that M_o only; parameters stay identical across ranks."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_mode3_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=180) for _ in procs], key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, d0, u0, m0, f0), (_, d1, u1, m1, f1) = res
    assert d0 == d1 and len(set(d0)) > 1              # same object sequence on both ranks, more than one object visited
    assert u0 and u1 and m0 and m1                    # inactive mappers untouched, active ones trained
    assert f0 == f1                                   # bit-identical parameters across ranks after 6 steps
