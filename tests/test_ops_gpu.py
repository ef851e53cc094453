"""-m gpu: every kernel of libviewneti_sm100a.so (called through the C-ABI) against a plain PyTorch fp32
reference of the same op on identical bf16-rounded inputs, at the SD-2.1 layer shapes."""
import pytest

from tests import opchecks

CASES = opchecks.all_checks()


@pytest.mark.gpu
@pytest.mark.parametrize("idx", range(len(CASES)), ids=[f"{i}-{c[0].__name__}" for i, c in enumerate(CASES)])
def test_op(idx):
    fn, kw = CASES[idx]
    for label, err, tol in fn(**kw):
        assert err <= tol, f"{label}: rel err {err:.3e} > {tol:.1e}"


@pytest.mark.gpu
@pytest.mark.parametrize("nb,hw,Cc", [(2, 6912, 960), (1, 4096, 320), (1, 256, 2560), (1, 64, 1280)])
def test_groupnorm_fused_repeats_bitwise(nb, hw, Cc):
    """The one-launch GroupNorm exchanges statistics between all CTAs of its grid through global memory; 200
    back-to-back repetitions (forward and backward) must reproduce the first result bit for bit - a lost or
    mis-ordered partial, or a shared-memory race, would show up as a flipped bit or a trap."""
    import torch
    from view_neti_b200 import ops
    dev = "cuda"
    g = torch.Generator(device="cpu").manual_seed(5)
    x = (torch.randn(nb, hw, Cc, generator=g) * 1.5 + 0.3).to(dev).to(torch.bfloat16)
    dy = torch.randn(nb, hw, Cc, generator=g).to(dev).to(torch.bfloat16)
    gamma, beta = (1 + 0.1 * torch.randn(Cc, generator=g)).to(dev), (0.1 * torch.randn(Cc, generator=g)).to(dev)
    stats = torch.zeros(nb, 32, 2, device=dev, dtype=torch.float64)
    red = torch.zeros(nb, 32, 2, device=dev, dtype=torch.float64)
    part = torch.empty(2, ops.groupnorm_partial_floats(nb), device=dev)
    y, dx = torch.empty_like(x), torch.empty_like(x)
    ref = None
    for it in range(200):
        ops.memset(part, 0xFF)
        ops.groupnorm_fwd(x, gamma, beta, 1e-5, True, y, nb, hw, 32, stats, part[0])
        ops.groupnorm_bwd_fused(x, dy, stats, red, part[1], gamma, beta, 1e-5, True, dx, nb, hw, 32)
        if ref is None:
            ref = (y.clone(), dx.clone(), stats.clone(), red.clone())
        elif it % 20 == 19 or it < 5:
            assert torch.equal(y, ref[0]) and torch.equal(dx, ref[1]), f"repetition {it} differs"
            assert torch.equal(stats, ref[2]) and torch.equal(red, ref[3]), f"statistics of repetition {it} differ"
    torch.cuda.synchronize()
