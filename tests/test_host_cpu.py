"""-m "not gpu": the C-ABI library loads and exports every symbol the header declares; host-side logic (context
protocol, schedulers, flat-gradient all-reduce over gloo with world_size 2)."""
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_loads_and_exports_every_declared_symbol():
    from view_neti_b200 import _abi
    from view_neti_b200.build import build
    build()
    lib = _abi.load()
    hdr = open(os.path.join(ROOT, "include", "viewneti.h")).read()
    declared = set(re.findall(r"\b(vn_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"vn_stream_t"}
    assert len(declared) >= 30
    assert declared == set(_abi.SIGNATURES), (declared ^ set(_abi.SIGNATURES))
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.vn_version() == 1
    assert isinstance(lib.vn_last_error(), bytes)


def test_product_path_fails_loudly_without_cuda():
    from view_neti_b200._abi import VNError
    from view_neti_b200.sd21 import TINY, init_state_dict
    from view_neti_b200.unet import UNet2DConditionModel
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises((VNError, RuntimeError, AssertionError)):
        UNet2DConditionModel(init_state_dict(TINY, 0), TINY, "cpu")


def test_context_protocol_static_binding():
    """Layer l reads entry (this_idx + l) % 16; K from CONTEXT_TENSOR, V from BYPASS when present (xti...:14-26)."""
    from view_neti_b200.sd21 import TINY
    from view_neti_b200.unet import UNet2DConditionModel
    m = UNet2DConditionModel.__new__(UNet2DConditionModel)
    m.cfg = TINY
    d = {"this_idx": 3}
    for i in range(16):
        d[f"CONTEXT_TENSOR_{i}"] = torch.full((1, 2, 2), float(i))
        if i % 2 == 0:
            d[f"CONTEXT_TENSOR_BYPASS_{i}"] = torch.full((1, 2, 2), 100.0 + i)
    c = UNet2DConditionModel._contexts(m, d)
    assert len(c) == 32 and d["this_idx"] == 3
    for l in range(16):
        i = (3 + l) % 16
        assert float(c[l][0, 0, 0]) == i
        assert float(c[16 + l][0, 0, 0]) == (100.0 + i if i % 2 == 0 else i)
    t = torch.zeros(1, 2, 2)
    assert all(x is t for x in UNet2DConditionModel._contexts(m, t))


def test_schedulers_match_oracle():
    import numpy as np
    from oracle import schedulers as O
    from view_neti_b200.schedulers import DDIMScheduler, DDPMScheduler, alphas_cumprod
    assert np.allclose(alphas_cumprod().double().numpy(), O.alphas_cumprod(), rtol=2e-6)
    s = DDIMScheduler("v_prediction")
    s.set_timesteps(50)
    assert s.timesteps.tolist() == O.ddim_timesteps(50).tolist()
    assert s.timesteps[0] == 981 and s.timesteps[-1] == 1
    g = torch.Generator().manual_seed(0)
    x, e = torch.randn(2, 4, 8, 8, generator=g), torch.randn(2, 4, 8, 8, generator=g)
    for pt in ("epsilon", "v_prediction"):
        s = DDIMScheduler(pt)
        s.set_timesteps(50)
        for t in (981, 501, 1):
            ours = s.step(e, t, x).prev_sample.numpy()
            ref = O.ddim_step(e.double().numpy(), t, x.double().numpy(), 50, pt)
            assert np.allclose(ours, ref, rtol=1e-4, atol=1e-5)
    d = DDPMScheduler()
    t = torch.tensor([0, 999])
    assert np.allclose(d.add_noise(x, e, t).numpy(), O.add_noise(x.double().numpy(), e.double().numpy(), t.numpy()), atol=1e-5)
    assert np.allclose(d.get_velocity(x, e, t).numpy(), O.get_velocity(x.double().numpy(), e.double().numpy(), t.numpy()), atol=1e-5)


def test_dpm_solver_pp_scheduler():
    """DPM-Solver++(2M) (reference training/validate.py:568, inference_dtu.py:304).  Diffusers is absent, so the oracle
    is pinned by identities of the published algorithm, then the host scheduler (and the five scalars the fused
    kernel receives) is held to the oracle."""
    import numpy as np
    from oracle import schedulers as O
    from view_neti_b200.schedulers import DDIMScheduler, DPMSolverMultistepScheduler
    assert O.dpmpp_timesteps(50).tolist()[:3] == [999, 979, 959] and O.dpmpp_timesteps(50)[-1] == 20
    assert O.dpmpp_timesteps(30)[-1] == 33 and len(O.dpmpp_timesteps(30)) == 30
    g = torch.Generator().manual_seed(0)
    x = torch.randn(1, 4, 8, 8, generator=g).double().numpy()
    outs = [torch.randn(1, 4, 8, 8, generator=g).double().numpy() for _ in range(50)]
    acp = O.alphas_cumprod()
    for pt in ("epsilon", "v_prediction"):
        # identity 1: a single first-order DPM-Solver++ step is the DDIM step between the same two timesteps
        ts = O.dpmpp_timesteps(10)
        one = O.dpmpp_2m_sample(lambda x_, t, i: outs[i], x, 1, pt)            # one step: t = 999 -> 0
        a_t, a_p = acp[999], acp[0]
        if pt == "epsilon":
            x0, eps = (x - np.sqrt(1 - a_t) * outs[0]) / np.sqrt(a_t), outs[0]
        else:
            x0, eps = np.sqrt(a_t) * x - np.sqrt(1 - a_t) * outs[0], np.sqrt(a_t) * outs[0] + np.sqrt(1 - a_t) * x
        assert np.allclose(one, np.sqrt(a_p) * x0 + np.sqrt(1 - a_p) * eps, rtol=1e-10, atol=1e-12)
        # identity 2: with a constant DATA prediction the second-order correction vanishes and every step moves the
        # sample along the straight DDIM line towards that x0: x_t = alpha_t x0 + sigma_t/sigma_s (x_s - alpha_s x0)
        target = outs[7]

        def model_const_x0(x_, t, i, pt=pt):
            a, s = np.sqrt(acp[t]), np.sqrt(1 - acp[t])
            return (x_ - a * target) / s if pt == "epsilon" else (a * x_ - target) / s

        got = O.dpmpp_2m_sample(model_const_x0, x, 20, pt)
        a0, s0 = np.sqrt(acp[999]), np.sqrt(1 - acp[999])
        af, sf = np.sqrt(acp[0]), np.sqrt(1 - acp[0])
        assert np.allclose(got, af * target + sf / s0 * (x - a0 * target), rtol=1e-9, atol=1e-10)
        # the host scheduler and its kernel coefficients against the oracle loop, orders 2M with and without the
        # first-order final step (N < 15)
        for steps in (50, 30, 12, 2):
            sch = DPMSolverMultistepScheduler.from_config(DDIMScheduler(pt).config)
            assert sch.config.prediction_type == pt and sch.config.solver_order == 2
            sch.set_timesteps(steps)
            assert sch.timesteps.tolist() == O.dpmpp_timesteps(steps).tolist()
            xs = torch.from_numpy(x)
            xk, x0p = x.copy(), np.zeros_like(x)
            for i, t in enumerate(sch.timesteps):
                xs = sch.step(torch.from_numpy(outs[i]), t, xs).prev_sample
                p_, q_, A, B0, B1 = sch.kernel_coefficients(i)
                x0k = p_ * xk + q_ * outs[i]
                xk, x0p = A * xk + B0 * x0k + B1 * x0p, x0k
                if i == 0 or (i == steps - 1 and steps < 15):
                    assert B1 == 0.0
            ref = O.dpmpp_2m_sample(lambda x_, t, i: outs[i], x, steps, pt)
            # the host table of alphas_cumprod is float32 (as diffusers builds it), the oracle's float64
            assert np.allclose(xs.numpy(), ref, rtol=2e-5, atol=2e-5) and np.allclose(xk, xs.numpy(), rtol=1e-12, atol=1e-12)
    sch = DPMSolverMultistepScheduler()
    sch.set_timesteps(5)
    with pytest.raises(ValueError):
        sch.step(torch.zeros(1), 5, torch.zeros(1))


def _worker(rank, world, port, q):
    import torch.distributed as dist
    from view_neti_b200.training.dist import FlatGradAllReducer
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    a, b, c, d = (torch.nn.Parameter(torch.zeros(3, 4)), torch.nn.Parameter(torch.zeros(5)), torch.nn.Parameter(torch.zeros(2)),
                  torch.nn.Parameter(torch.zeros(3)))
    a.grad = torch.full((3, 4), float(rank + 1))
    b.grad = torch.arange(5.0) * (rank + 1)          # c.grad stays None on every rank -> must stay None (AdamW skips it)
    if rank == 1:
        d.grad = torch.full((3,), 4.0)               # used on one rank only -> every rank gets the mean (the other counts as 0)
    red = FlatGradAllReducer([a, b, c, d])
    flat = red.allreduce_()
    # the caller names the parameters a step used (mode 3: M_v + the object mapper rank 0 drew): only those travel
    e, f = torch.nn.Parameter(torch.zeros(4)), torch.nn.Parameter(torch.zeros(4))
    e.grad = torch.full((4,), float(2 * rank))
    red2 = FlatGradAllReducer([e, f])
    part = red2.allreduce_(active=[e])
    # construction-time broadcast (DDP semantics): ranks start from rank 0's values whatever their local init was
    w = torch.nn.Parameter(torch.full((4,), float(10 + rank)))
    FlatGradAllReducer([w]).broadcast_parameters_(0)
    # plain lists, not tensors: a tensor in the queue shares memory with a process that may be gone when the parent reads it
    q.put((rank, a.grad.tolist(), b.grad.tolist(), c.grad, d.grad.tolist(), flat.numel(), w.detach().tolist(), e.grad.tolist(),
           f.grad, part.numel()))
    dist.destroy_process_group()


def test_flat_gradient_allreduce_gloo_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, a, b, c, d, n, w, e, f, npart in res:
        a, b, d, w, e = (torch.tensor(v) for v in (a, b, d, w, e))
        assert torch.equal(w, torch.full((4,), 10.0))
        assert n == 12 + 5 + 2 + 3
        assert torch.allclose(a, torch.full((3, 4), 1.5))            # mean of 1 and 2
        assert torch.allclose(b, torch.arange(5.0) * 1.5)
        assert c is None                                              # unused on every rank: no gradient appears
        assert torch.allclose(d, torch.full((3,), 2.0))               # mean of (missing = 0) and 4
        assert torch.allclose(e, torch.full((4,), 1.0)) and f is None and npart == 4


def test_reducer_is_a_no_op_in_one_process_and_leaves_unused_mappers_alone():
    """with one process nothing is reduced and no zero gradients are invented, so AdamW neither decays
    nor moves a mapper that took no part in the step (reference coach.py:736-757 after zero_grad)."""
    from view_neti_b200.training.dist import FlatGradAllReducer
    used, unused = torch.nn.Linear(4, 4), torch.nn.Linear(4, 4)
    params = list(used.parameters()) + list(unused.parameters())
    opt = torch.optim.AdamW(params, lr=0.1, weight_decay=0.5)
    before = [p.detach().clone() for p in unused.parameters()]
    red = FlatGradAllReducer(params)
    for _ in range(3):
        used(torch.ones(2, 4)).sum().backward()
        assert red.allreduce_() is None
        assert all(p.grad is None for p in unused.parameters())
        opt.step()
        opt.zero_grad(set_to_none=True)
    assert all(torch.equal(p.detach(), b) for p, b in zip(unused.parameters(), before))      # bit-identical


def test_synthetic_conditioning_emits_xti_dict():
    from view_neti_b200.training.coach import SyntheticConditioning
    c = SyntheticConditioning(dim=64, rank=4)
    d = c(timesteps=torch.zeros(3, dtype=torch.long))
    assert d["this_idx"] == 0 and len([k for k in d if k.startswith("CONTEXT_TENSOR")]) == 32
    assert d["CONTEXT_TENSOR_BYPASS_15"].shape == (3, 77, 64) and d["CONTEXT_TENSOR_0"].requires_grad


def test_neti_mapper_structure_matches_reference():
    """Parameter names / count and the Fourier matrix of NeTIMapper against the reference-generated golden
    (tests/golden/make_golden_mapper.py); the arithmetic itself is CUDA-only and tested under -m gpu."""
    from view_neti_b200._abi import VNError
    from view_neti_b200.models.neti_mapper import NeTIMapper
    from view_neti_b200.utils.types import PESigmas
    gold = torch.load(os.path.join(ROOT, "tests", "golden", "neti_mapper.pt"))
    sig = PESigmas(sigma_t=0.03, sigma_l=2.0, sigma_theta=0.5, sigma_phi=0.5, sigma_r=0.5, sigma_dtu12=0.5)
    kw = dict(arch_mlp_hidden_dims=64, arch_view_net=15, arch_view_disable_tl=False, use_nested_dropout=False,
              pe_sigmas=sig, output_bypass=True)
    mo = NeTIMapper(embedding_type="object", output_dim=1024, **kw)
    assert sum(p.numel() for p in mo.parameters()) == 141_696
    assert set(mo.state_dict()) == {k for k in gold["object"]["state"] if k != "encoder.w"}
    mv = NeTIMapper(embedding_type="view", output_dim=256, placeholder_view_tokens=gold["view"]["tokens"],
                    placeholder_view_token_ids=gold["view"]["ids"], **kw)
    assert mv.deg_freedom == "theta-phi" and torch.equal(mv.encoder_w, gold["view"]["w"])
    assert torch.equal(NeTIMapper(embedding_type="object", output_dim=256, **kw).encoder_w, gold["object"]["w"])
    x = mv._encode_inputs(torch.tensor([0., 999.]), torch.tensor([0., 15.]), torch.tensor([49408, 49411]))
    assert torch.allclose(x, torch.tensor([[-1., -1., -1., -1.], [0.998, 0.875, 1., 1.]]))
    with pytest.raises(VNError):
        mv(torch.tensor([0.]), torch.tensor([0.]), torch.tensor([49408]))           # CPU: fails loudly, no fallback


def test_gemm_tuning_table_is_well_formed():
    """view_neti_b200/gemm_tuning.json (written by scripts/gemm_autotune.py on B200) only holds choices vn_gemm accepts:
    BN in {64,128,160,192,256}, split-K cluster in {1,2,4,8} with BN/split a multiple of 8, BN 160 never unsplit,
    CTA pairs (bit 8) only without split-K and for BN >= 128."""
    from view_neti_b200 import ops
    assert len(ops.TUNING) > 50
    for key, (bn, code) in ops.TUNING.items():
        mode, M, N, K, H, W = (int(v) for v in key.split(":"))
        assert mode in (0, 1) and M > 0 and N % 8 == 0 and K % 64 == 0, key
        split, pair = code & 15, (code >> 8) & 3
        assert bn in (64, 128, 160, 192, 256), (key, bn)
        assert split in (1, 2, 4, 8) and (bn // split) % 8 == 0, (key, bn, split)
        assert not (bn == 160 and split == 1), key
        if pair == 1:
            assert split == 1 and bn in (128, 192, 256), (key, bn, code)
            if mode == 0:                      # (conv m-tiles depend on the spatial tiling; vn_gemm falls back if odd)
                assert ((M + 127) // 128) % 2 == 0, key


def test_prompt_manager_host_logic():
    """PromptManager's host side (reference prompt_manager.py:52-71): tokenizer call convention and which placeholder
    ids a prompt holds (-1 when none, exactly one of a kind otherwise); the embedding itself is CUDA-only."""
    from types import SimpleNamespace
    from view_neti_b200.prompt_manager import PromptManager

    class Tok:
        model_max_length = 77

        def __call__(self, text, padding=None, max_length=None, return_tensors=None):
            assert padding == "max_length" and max_length == 77 and return_tensors == "pt"
            ids = torch.full((1, 77), 7)
            ids[0, 2], ids[0, 5] = 49410, 49408
            return SimpleNamespace(input_ids=ids)

    cond = SimpleNamespace(n_layers=16)
    pm = PromptManager(Tok(), cond, timesteps=[999, 500], placeholder_view_token_ids=[49409, 49410],
                       placeholder_object_token_ids=[49408])
    ids = pm._tokenize("<view_10_40_1p2>. a photo of a <statue>")
    assert ids.shape == (1, 77)
    assert pm._placeholder(ids, pm.placeholder_view_token_ids, "p").tolist() == [49410]
    assert pm._placeholder(ids, pm.placeholder_object_token_ids, "p").tolist() == [49408]
    assert pm._placeholder(ids, [55555], "p").tolist() == [-1] and pm._placeholder(ids, None, "p").tolist() == [-1]
    two = ids.clone()
    two[0, 9] = 49409
    with pytest.raises(AssertionError):
        pm._placeholder(two, pm.placeholder_view_token_ids, "p")
    assert torch.equal(pm._tokenize(torch.arange(77)), torch.arange(77).view(1, 77))
    from view_neti_b200 import constants
    assert len(constants.SD_INFERENCE_TIMESTEPS) == 50 and constants.SD_INFERENCE_TIMESTEPS[:3] == [999, 979, 959]
    assert constants.SD_INFERENCE_TIMESTEPS[24:27] == [519, 500, 480] and constants.SD_INFERENCE_TIMESTEPS[-1] == 20


def test_checkpoint_handler_reads_reference_format_and_round_trips(tmp_path):
    """tests/golden/mapper_ckpt_{object,view}.pt are in the layout the reference writes, built from the reference's own
    NeTIMapper / RunConfig objects (tests/golden/make_golden_checkpoint.py) - they pickle reference classes.  They must
    load here WITHOUT the reference on sys.path, give back the exact weights, and survive a save / load round trip."""
    import sys
    from view_neti_b200.checkpoint_handler import CheckpointHandler
    assert not any(p.rstrip("/").endswith("reference") for p in sys.path)
    gdir = os.path.join(ROOT, "tests", "golden")
    raw = torch.load(os.path.join(gdir, "mapper_ckpt_object.pt"), map_location="cpu", weights_only=False,
                     pickle_module=__import__("view_neti_b200.checkpoint_handler", fromlist=["_PickleModule"])._PickleModule)
    cfg, objs = CheckpointHandler.load_mapper(os.path.join(gdir, "mapper_ckpt_object.pt"), "object",
                                              placeholder_object_tokens=["<statue>", "<teapot>"],
                                              placeholder_object_token_ids=[49408, 49409])
    assert cfg["model"]["arch_view_net"] == 15 and sorted(objs) == [49408, 49409]
    for tid, m in objs.items():
        ref_sd = raw["mappers"][tid]["state_dict"]
        assert m.placeholder_object_token == raw["mappers"][tid]["placeholder_object_token"]
        for k, v in m.state_dict().items():
            assert torch.equal(v, ref_sd[k]), k
        assert torch.equal(m.encoder_w, ref_sd["encoder.w"]) and abs(float(m.norm_scale) - 0.3714) < 1e-6
        assert m.bypass_unconstrained and m.output_bypass and sum(p.numel() for p in m.parameters()) == 64 * 64 * 2 + 64 * 4 + 64 * 2 + 256 * 64 + 256
    toks, ids = ["<view_0_10_1p2>", "<view_10_40_1p2>", "<view_20_70_1p2>"], [49410, 49411, 49412]
    cfg, mv = CheckpointHandler.load_mapper(os.path.join(gdir, "mapper_ckpt_view.pt"), "view", placeholder_view_tokens=toks,
                                            placeholder_view_token_ids=ids)
    assert mv.deg_freedom == "theta-phi" and mv.embedding_type == "view"
    # round trip through our writer (file names as the reference derives them, checkpoint_handler.py:73-76,93-96)
    h = CheckpointHandler(cfg=cfg, placeholder_view_tokens=toks, placeholder_view_token_ids=ids,
                          placeholder_object_tokens=["<statue>", "<teapot>"], placeholder_object_token_ids=[49408, 49409],
                          save_root=tmp_path)
    h.save_mapper(objs, mv, "mapper-steps-250.pt")
    assert sorted(os.listdir(tmp_path)) == ["mapper-steps-250_object.pt", "mapper-steps-250_view.pt"]
    _, objs2 = CheckpointHandler.load_mapper(tmp_path / "mapper-steps-250_object.pt", "object",
                                             placeholder_object_tokens=["<statue>", "<teapot>"], placeholder_object_token_ids=[49408, 49409])
    _, mv2 = CheckpointHandler.load_mapper(tmp_path / "mapper-steps-250_view.pt", "view", placeholder_view_tokens=toks,
                                           placeholder_view_token_ids=ids)
    for a, b in ((objs[49408], objs2[49408]), (objs[49409], objs2[49409]), (mv, mv2)):
        assert all(torch.equal(x, y) for x, y in zip(a.state_dict().values(), b.state_dict().values()))
    emb = torch.randn(49420, 16)
    h.save_learned_embeds(emb, "learned_embeds-steps-250.bin")
    tokens, rows = CheckpointHandler.load_learned_embeds(tmp_path / "learned_embeds-steps-250.bin")
    assert tokens == toks + ["<statue>", "<teapot>"] and torch.equal(rows, emb[ids + [49408, 49409]])
