"""GPU parity: FunctionCorrelation through the C ABI against golden vectors and the oracle."""
import pytest
import torch

from conftest import load_golden
from oracle import correlation_ref

pytestmark = pytest.mark.gpu
TOL = 1e-4  # fp32 dot products of N(0,1) data, different summation order than the reference


@pytest.mark.parametrize("case", ["correlation_c8", "correlation_c32", "correlation_c196"])
def test_golden(case):
    from motif_b200.correlation import FunctionCorrelation

    g = load_golden(case)
    out = FunctionCorrelation(g["first"].cuda(), g["second"].cuda())
    assert out.shape == g["out"].shape
    assert (out.cpu() - g["out"]).abs().max().item() < TOL


# PWC-Net pyramid levels of a 720x1280 pair, the two smallest at full size, plus ragged sizes
@pytest.mark.parametrize("shape", [(1, 196, 12, 20), (1, 128, 24, 40), (2, 96, 13, 27), (1, 64, 33, 70), (1, 32, 7, 5), (1, 3, 1, 1)])
def test_vs_oracle(shape):
    from motif_b200.correlation import FunctionCorrelation, ModuleCorrelation

    gen = torch.Generator().manual_seed(sum(shape))
    f1 = torch.randn(*shape, generator=gen)
    f2 = torch.randn(*shape, generator=gen)
    ref = correlation_ref.function_correlation(f1, f2)
    out = FunctionCorrelation(f1.cuda(), f2.cuda())
    assert (out.cpu() - ref).abs().max().item() < TOL
    assert torch.equal(ModuleCorrelation()(f1.cuda(), f2.cuda()), out)


def test_full_size_properties():
    from motif_b200.correlation import FunctionCorrelation

    torch.manual_seed(0)
    f1 = torch.randn(1, 32, 192, 320, device="cuda")
    f2 = torch.randn(1, 32, 192, 320, device="cuda")
    out = FunctionCorrelation(f1, f2)
    # centre displacement is the channel mean of the product; symmetry corr(a,b)[d](p) == corr(b,a)[-d](p+d)
    assert (out[:, 40] - (f1 * f2).mean(1)).abs().max().item() < 1e-5
    swapped = FunctionCorrelation(f2, f1)
    d = 9 * 2 + 7  # dy=-2, dx=+3
    dm = 9 * 6 + 1  # dy=+2, dx=-3
    a = out[0, d, 2:, :-3]
    b = swapped[0, dm, :-2, 3:]
    assert (a - b).abs().max().item() < 1e-5
    # linear in the first argument
    assert (FunctionCorrelation(3 * f1, f2) - 3 * out).abs().max().item() < 1e-4


def test_contiguity_assert_like_reference():
    from motif_b200.correlation import FunctionCorrelation

    x = torch.randn(1, 4, 8, 8, device="cuda")
    with pytest.raises(AssertionError):
        FunctionCorrelation(x.permute(0, 1, 3, 2), x)
