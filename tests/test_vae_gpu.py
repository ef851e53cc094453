"""GPU parity of the VAE path (reference training/coach.py:165-169, sd_pipeline_call.py:115) against the
fp32 CPU oracle (oracle/vae.py), through the drop-in AutoencoderKL surface and the C-ABI kernels underneath.

Tolerance: activations are stored in bf16 between kernels (fp32 accumulation inside), which puts ~1.1e-2 relative L2 on
the outputs (measured with the CPU emulation of the same launch sequence, tests/test_vae_cpu.py); the bar is 2.5e-2.
The two data-movement kernels (im2col, softmax rounding aside) are checked bit-exactly / to bf16 rounding.
Also here: the rest of the image-side tail of inference - the denoise loop under DPM-Solver++(2M), the scheduler the
reference's inference scripts install, and the pipeline's decode to numpy / PIL."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
TOL = 2.5e-2


def rel(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    if not torch.isfinite(a).all():
        return float("inf")
    return float((a - b).norm() / (b.norm() + 1e-20))


@pytest.mark.parametrize("rows,cols", [(64, 64), (256, 256), (300, 1024), (4096, 4096), (130, 6912), (9, 8192)])
def test_softmax_rows(rows, cols):
    from view_neti_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(rows + cols)
    S = torch.randn(rows, cols, device="cuda", generator=g) * 30
    P = torch.empty(rows, cols, dtype=torch.bfloat16, device="cuda")
    scale = 1.0 / math.sqrt(512)
    ops.softmax_rows(S, P, scale)
    ref = torch.softmax(S.double() * scale, dim=-1)
    err = (P.double() - ref).abs()
    assert bool((err <= ref * 2.0 ** -8 + 1e-7).all()), float((err / (ref + 1e-7)).max())
    assert float((P.double().sum(-1) - 1).abs().max()) < 2e-3
    # a strided destination / source (row stride > cols)
    Sb = torch.randn(rows, cols + 64, device="cuda", generator=g)
    Pb = torch.zeros(rows, cols + 64, dtype=torch.bfloat16, device="cuda")
    ops.softmax_rows(Sb[:, :cols], Pb[:, :cols], 1.0)
    refb = torch.softmax(Sb[:, :cols].double(), dim=-1)
    assert bool(((Pb[:, :cols].double() - refb).abs() <= refb * 2.0 ** -8 + 1e-7).all())
    assert float(Pb[:, cols:].abs().sum()) == 0.0


@pytest.mark.parametrize("nb,H,W,C", [(1, 8, 8, 64), (2, 16, 24, 128), (1, 64, 64, 256), (1, 6, 10, 64)])
def test_im2col_s2_pad0_bit_exact(nb, H, W, C):
    from view_neti_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn(nb, H, W, C, device="cuda", generator=g).to(torch.bfloat16)
    Ho, Wo = H // 2, W // 2
    col = torch.full((nb * Ho * Wo, 9 * C), 7.0, dtype=torch.bfloat16, device="cuda")
    ops.im2col_s2_pad0(x, col)
    xp = F.pad(x.float().permute(0, 3, 1, 2), (0, 1, 0, 1))
    u = F.unfold(xp, 3, stride=2).view(nb, C, 9, Ho * Wo).permute(0, 3, 2, 1).reshape(nb * Ho * Wo, 9 * C)
    assert torch.equal(col.float(), u)


@pytest.mark.parametrize("nb,Ct,H,W", [(1, 3, 16, 16), (2, 4, 8, 24), (1, 3, 512, 512), (3, 7, 5, 9)])
def test_im2col_thin_bit_exact(nb, Ct, H, W):
    from view_neti_b200 import ops
    x = torch.randn(nb, Ct, H, W, device="cuda", generator=torch.Generator(device="cuda").manual_seed(1))
    col = torch.full((nb * H * W, 64), 3.0, dtype=torch.bfloat16, device="cuda")
    ops.im2col_thin(x, col)
    u = F.unfold(x, 3, padding=1).view(nb, Ct, 9, H * W).permute(0, 3, 2, 1).reshape(nb * H * W, 9 * Ct)
    assert torch.equal(col[:, : 9 * Ct], u.to(torch.bfloat16)) and float(col[:, 9 * Ct:].abs().sum()) == 0.0


@pytest.mark.parametrize("nb,H,W,C,Ct", [(1, 64, 64, 512, 8), (1, 128, 128, 128, 3), (2, 24, 40, 64, 3)])
def test_conv_to_few_channels_on_the_gemm(nb, H, W, C, Ct):
    """conv_out of the VAE as the implicit 3x3 GEMM with N padded to 8 and an fp32 destination."""
    from view_neti_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(2)
    x = torch.randn(nb, H, W, C, device="cuda", generator=g).to(torch.bfloat16)
    w = (torch.randn(Ct, C, 3, 3, device="cuda", generator=g) / math.sqrt(9 * C)).to(torch.bfloat16)
    b = torch.randn(8, device="cuda", generator=g)
    wk = torch.zeros(8, 9 * C, dtype=torch.bfloat16, device="cuda")
    wk[:Ct] = w.permute(0, 2, 3, 1).reshape(Ct, -1)
    d = torch.full((nb, H, W, 8), 9.0, device="cuda")
    ops.conv3x3(x, wk, d, bias=b, ws=ops.Workspace(8192, 8192, "cuda"), force_bn=64, force_split=1)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.float(), b[:Ct], padding=1).permute(0, 2, 3, 1)
    assert rel(d[..., :Ct], ref) < 1e-5
    assert rel(d[..., Ct:], b[Ct:].expand(nb, H, W, 8 - Ct)) < 1e-6 if Ct < 8 else True


def test_gemm_with_computed_strided_operands():
    """The attention products of the VAE: A and B are column slices of one [hw, 2C] buffer written by the launch
    before, fp32 output; V^T straight out of a GEMM with the weight as the A operand."""
    from view_neti_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(5)
    hw, C = 1024, 512
    t = torch.randn(hw, C, device="cuda", generator=g).to(torch.bfloat16)
    Wqk = (torch.randn(2 * C, C, device="cuda", generator=g) / math.sqrt(C)).to(torch.bfloat16)
    Wv = (torch.randn(C, C, device="cuda", generator=g) / math.sqrt(C)).to(torch.bfloat16)
    ws = ops.Workspace(8192, 8192, "cuda")
    qk = torch.empty(hw, 2 * C, dtype=torch.bfloat16, device="cuda")
    S = torch.empty(hw, hw, device="cuda")
    vt = torch.empty(C, hw, dtype=torch.bfloat16, device="cuda")
    for _ in range(3):                                   # back to back: the early B fetch would race without b_dynamic
        ops.gemm(t, Wqk, qk, ws=ws)
        ops.gemm(qk[:, :C], qk[:, C:], S, ws=ws, b_dynamic=True)
        ops.gemm(Wv, t, vt, ws=ws, b_dynamic=True)
    qkr = (t.float() @ Wqk.float().t())
    assert rel(qk, qkr) < 5e-3
    assert rel(S, qk[:, :C].float() @ qk[:, C:].float().t()) < 1e-5
    assert rel(vt, (t.float() @ Wv.float().t()).t()) < 5e-3


@pytest.fixture(scope="module")
def sd21_vae():
    from view_neti_b200.models.vae import SD21_VAE, AutoencoderKL, init_state_dict
    sd = init_state_dict(SD21_VAE, 0)
    return sd, AutoencoderKL(sd, SD21_VAE, "cuda")


def _check(cfg, sd, vae, nb, H, W, seed=1):
    from oracle import vae as ovae
    g = torch.Generator().manual_seed(seed)
    img = torch.rand(nb, 3, H, W, generator=g) * 2 - 1
    z = torch.randn(nb, 4, H // 8, W // 8, generator=g)
    mean_o, logvar_o = ovae.encode_moments(sd, cfg, img)
    dec_o = ovae.decode(sd, cfg, z)
    dist = vae.encode(img.cuda()).latent_dist
    dec = vae.decode(z.cuda()).sample
    r = (rel(dist.mean, mean_o), rel(dist.logvar, logvar_o), rel(dec, dec_o))
    assert max(r) < TOL, r
    return r


@pytest.mark.parametrize("nb,H,W", [(1, 64, 64), (2, 64, 64), (3, 64, 128)])
def test_tiny_vae_matches_oracle(nb, H, W):
    from view_neti_b200.models.vae import TINY_VAE, AutoencoderKL, init_state_dict
    sd = init_state_dict(TINY_VAE, 0)
    _check(TINY_VAE, sd, AutoencoderKL(sd, TINY_VAE, "cuda"), nb, H, W)


@pytest.mark.parametrize("nb,H,W", [(1, 128, 128), (2, 64, 192), (1, 512, 512), (1, 384, 512)])
def test_sd21_vae_matches_oracle(sd21_vae, nb, H, W):
    """BASELINE sizes included: 512 x 512 (the metric's resolution) and the DTU default 512 x 384 (dataset.py:711-712)."""
    from view_neti_b200.models.vae import SD21_VAE
    sd, vae = sd21_vae
    _check(SD21_VAE, sd, vae, nb, H, W)


def test_sd21_vae_properties_at_full_size(sd21_vae):
    """Size-independent properties at 512 x 512: bitwise repeatable, a batch is its images (to rounding), and the drop-in call shapes
    of coach.py:165-169 / sd_pipeline_call.py:115."""
    from oracle import vae as ovae
    from view_neti_b200.models.vae import SD21_VAE, decode_latents
    sd, vae = sd21_vae
    g = torch.Generator().manual_seed(3)
    img = (torch.rand(2, 3, 512, 512, generator=g) * 2 - 1).cuda()
    d1 = vae.encode(img[:1]).latent_dist
    m1, l1 = d1.mean.clone(), d1.logvar.clone()
    d1b = vae.encode(img[:1]).latent_dist
    assert torch.equal(d1b.mean, m1) and torch.equal(d1b.logvar, l1)
    d2 = vae.encode(img).latent_dist
    assert d2.mean.shape == (2, 4, 64, 64)
    # a batch is its images - up to rounding: GEMM tiling / split-K (summation order) is chosen per problem size
    assert rel(d2.mean[:1], m1) < 1e-2 and rel(d2.logvar[:1], l1) < 1e-2, (rel(d2.mean[:1], m1), rel(d2.logvar[:1], l1))
    gen = torch.Generator(device="cuda").manual_seed(0)
    latents = d2.sample(gen) * vae.config.scaling_factor              # coach.py:167-169
    noise = torch.randn(d2.mean.shape, generator=torch.Generator(device="cuda").manual_seed(0), device="cuda")
    assert torch.equal(latents, (d2.mean + d2.std * noise) * 0.18215)
    image = decode_latents(vae, latents[:1])                           # sd_pipeline_call.py:115
    assert image.shape == (1, 512, 512, 3) and image.dtype.name == "float32" and image.min() >= 0 and image.max() <= 1
    want = ovae.decode_latents(sd, SD21_VAE, latents[:1].cpu()).numpy()
    assert abs(image - want).max() < 6e-2 and abs(image - want).mean() < 5e-3


def test_vae_rejects_bad_shapes(sd21_vae):
    from view_neti_b200._abi import VNError
    _, vae = sd21_vae
    with pytest.raises(VNError):
        vae.encode(torch.zeros(1, 3, 60, 64, device="cuda"))
    with pytest.raises(VNError):
        vae.encode(torch.zeros(1, 3, 32, 32, device="cuda"))          # 4 x 4 latents: attention needs hw % 64 == 0
    with pytest.raises(VNError):
        vae.decode(torch.zeros(1, 3, 8, 8, device="cuda"))


def test_coach_step_from_pixel_values_equals_step_from_its_latents():
    """coach.py:165-218 with the VAE in the loop: a step fed `pixel_values` is the step fed the latents the VAE
    samples for them (same generator stream); dict batches drive Coach.train as the reference's loader does."""
    from types import SimpleNamespace
    from view_neti_b200.models.vae import TINY_VAE, AutoencoderKL, init_state_dict as vae_sd
    from view_neti_b200.schedulers import DDPMScheduler
    from view_neti_b200.sd21 import TINY, init_state_dict
    from view_neti_b200.training.coach import Coach, SyntheticConditioning
    from view_neti_b200.unet import UNet2DConditionModel
    unet = UNet2DConditionModel(init_state_dict(TINY, 0), TINY, "cuda")
    vae = AutoencoderKL(vae_sd(TINY_VAE, 0), TINY_VAE, "cuda")
    img = (torch.rand(2, 3, 128, 128, generator=torch.Generator().manual_seed(2)) * 2 - 1).cuda()
    cfg = SimpleNamespace(optim=SimpleNamespace(learning_rate=1e-2, max_train_steps=2))

    def coach():
        cond = SyntheticConditioning(dim=TINY.cross_attention_dim, rank=4).cuda()
        return Coach(cfg, unet, cond, DDPMScheduler("v_prediction"), generator=torch.Generator(device="cuda").manual_seed(7),
                     vae=vae), cond

    a, cond_a = coach()
    loss_a = a.train_step(batch={"pixel_values": img})
    b, cond_b = coach()
    lat = vae.encode(img).latent_dist.sample(b.generator) * 0.18215
    assert lat.shape == (2, 4, 16, 16) and not lat.requires_grad
    loss_b = b.train_step(lat)
    assert abs(float(loss_a) - float(loss_b)) <= 1e-5 * abs(float(loss_b)), (float(loss_a), float(loss_b))
    assert rel(cond_a.base, cond_b.base) < 1e-4
    c, cond_c = coach()
    losses = c.train([{"pixel_values": img}] * 4)
    assert len(losses) == 2 and abs(float(losses[0]) - float(loss_a)) <= 1e-5 * abs(float(loss_a))
    no_vae = Coach(cfg, unet, SyntheticConditioning(dim=TINY.cross_attention_dim, rank=4).cuda(), DDPMScheduler("v_prediction"))
    with pytest.raises(ValueError):
        no_vae.train_step(batch={"pixel_values": img})


def test_sd_pipeline_call_decodes_through_the_vae():
    """sd_pipeline_call.py:104-129: output_type "np" / "pil" run pipeline.decode_latents on the final latents."""
    import numpy as np
    from oracle import vae as ovae
    from view_neti_b200.models.vae import TINY_VAE, AutoencoderKL, init_state_dict as vae_sd
    from view_neti_b200.schedulers import DDIMScheduler
    from view_neti_b200.sd21 import TINY, init_state_dict
    from view_neti_b200.sd_pipeline_call import ViewNeTIPipeline, sd_pipeline_call
    from view_neti_b200.unet import UNet2DConditionModel
    unet = UNet2DConditionModel(init_state_dict(TINY, 0), TINY, "cuda")
    sd = vae_sd(TINY_VAE, 0)
    g = torch.Generator().manual_seed(4)
    neg = torch.randn(1, 77, TINY.cross_attention_dim, generator=g).cuda()
    embeds = [torch.randn(1, 77, TINY.cross_attention_dim, generator=g).cuda() for _ in range(2)]
    x0 = torch.randn(1, 4, 16, 16, generator=g).cuda()
    pipe = ViewNeTIPipeline(unet, DDIMScheduler("v_prediction"), negative_prompt_embeds=neg, vae=AutoencoderKL(sd, TINY_VAE, "cuda"))
    kw = dict(height=128, width=128, num_inference_steps=2, guidance_scale=3.0, latents=x0)
    lat = sd_pipeline_call(pipe, embeds, output_type="latent", **kw).images
    img = sd_pipeline_call(pipe, embeds, output_type="np", **kw).images
    assert isinstance(img, np.ndarray) and img.shape == (1, 128, 128, 3) and img.dtype == np.float32
    want = ovae.decode_latents(sd, TINY_VAE, lat.float().cpu()).numpy()
    assert abs(img - want).max() < 6e-2 and abs(img - want).mean() < 5e-3
    pil = sd_pipeline_call(pipe, embeds, output_type="pil", **kw).images
    assert len(pil) == 1 and pil[0].size == (128, 128)
    assert abs(np.asarray(pil[0]).astype(np.float32) / 255 - img[0]).max() <= 0.5 / 255 + 1e-6
    bare = ViewNeTIPipeline(unet, DDIMScheduler("v_prediction"), negative_prompt_embeds=neg)
    from view_neti_b200._abi import VNError
    with pytest.raises(VNError):
        sd_pipeline_call(bare, embeds, output_type="np", **kw)


@pytest.mark.parametrize("steps,pt", [(6, "v_prediction"), (16, "epsilon")])
def test_sd_pipeline_call_with_dpm_solver_matches_oracle_loop(steps, pt):
    """sd_pipeline_call.py:71-101 under the scheduler the reference's inference scripts install (validate.py:568,
    inference_dtu.py:304): batched CFG + fused DPM-Solver++(2M) step against the oracle's two-pass loop; 6 steps end with
    the first-order final step (N < 15), 16 steps do not."""
    from oracle import schedulers as O
    from oracle.unet_sd21 import UNetOracle
    from view_neti_b200.schedulers import DDIMScheduler, DPMSolverMultistepScheduler
    from view_neti_b200.sd21 import TINY, init_state_dict
    from view_neti_b200.sd_pipeline_call import ViewNeTIPipeline, sd_pipeline_call
    from view_neti_b200.unet import UNet2DConditionModel
    gs = 5.0
    g = torch.Generator().manual_seed(21)
    x0 = torch.randn(1, 4, 16, 16, generator=g)
    neg = torch.randn(1, 77, TINY.cross_attention_dim, generator=g)
    embeds = []
    for _ in range(steps):
        d = {"this_idx": 0}
        for i in range(16):
            d[f"CONTEXT_TENSOR_{i}"] = torch.randn(1, 77, TINY.cross_attention_dim, generator=g)
            d[f"CONTEXT_TENSOR_BYPASS_{i}"] = torch.randn(1, 77, TINY.cross_attention_dim, generator=g)
        embeds.append(d)
    sd = init_state_dict(TINY, 0)
    oracle = UNetOracle(TINY)
    oracle.load_state_dict(sd)

    def model(x, t, i):
        xt = torch.from_numpy(x).float()
        with torch.no_grad():
            u = oracle(xt, t, neg).sample
            c = oracle(xt, t, dict(embeds[i])).sample
        return (u + gs * (c - u)).double().numpy()

    want = O.dpmpp_2m_sample(model, x0.double().numpy(), steps, pt)
    unet = UNet2DConditionModel(sd, TINY, "cuda")
    pipe = ViewNeTIPipeline(unet, DPMSolverMultistepScheduler.from_config(DDIMScheduler(pt).config),
                            negative_prompt_embeds=neg.cuda())
    cuda_embeds = [{k: (v.cuda() if torch.is_tensor(v) else v) for k, v in d.items()} for d in embeds]
    seen = []
    out = sd_pipeline_call(pipe, cuda_embeds, height=128, width=128, num_inference_steps=steps, guidance_scale=gs,
                           latents=x0.cuda(), output_type="latent", callback=lambda i, t, l: seen.append(t))
    assert seen == O.dpmpp_timesteps(steps).tolist()
    assert rel(out.images, torch.from_numpy(want)) < 5e-2          # CFG amplifies the bf16 error of the UNet output
