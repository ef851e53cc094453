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
