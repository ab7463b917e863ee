"""GPU parity: modulated deformable convolution (models/modules/DCNv2) through motif_dcn_v2_fwd."""
import pytest
import torch

from oracle import dcn_v2_ref

pytestmark = pytest.mark.gpu


def _case(B, Cin, Cout, H, W, dg, sigma, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, Cin, H, W, generator=g)
    off = torch.randn(B, dg * 18, H, W, generator=g) * sigma
    m = torch.sigmoid(torch.randn(B, dg * 9, H, W, generator=g))
    w = torch.randn(Cout, Cin, 3, 3, generator=g) / (Cin * 9) ** 0.5
    b = torch.randn(Cout, generator=g)
    return x, off, m, w, b


@pytest.mark.parametrize("shape,sigma", [((2, 16, 12, 9, 11, 4), 2.5), ((1, 64, 64, 45, 80, 8), 1.0), ((1, 64, 64, 20, 28, 8), 30.0),
                                          ((1, 8, 70, 5, 7, 1), 0.0), ((3, 24, 3, 1, 1, 3), 0.7)])
def test_vs_oracle(shape, sigma):
    """The model's configuration (64 -> 64, 8 groups), far out-of-frame offsets, zero offsets (= a plain 3x3 convolution
    times the mask), more than one output-channel block, a single pixel."""
    from motif_b200.dcn_v2 import dcn_v2_conv

    B, Cin, Cout, H, W, dg = shape
    x, off, m, w, b = _case(B, Cin, Cout, H, W, dg, sigma, seed=sum(shape))
    ref = dcn_v2_ref.dcn_v2_conv(x, off, m, w, b, 1, 1, 1, dg)
    with torch.no_grad():
        out = dcn_v2_conv(x.cuda(), off.cuda(), m.cuda(), w.cuda(), b.cuda(), 1, 1, 1, dg).cpu()
    assert out.shape == ref.shape
    assert (out - ref).abs().max().item() < 2e-5 * (1.0 + ref.abs().max().item())


def test_zero_offsets_equal_plain_convolution():
    from motif_b200.dcn_v2 import dcn_v2_conv

    x, off, m, w, b = _case(1, 64, 64, 24, 40, 8, 0.0, seed=9)
    m = torch.ones_like(m)
    with torch.no_grad():
        out = dcn_v2_conv(x.cuda(), off.cuda(), m.cuda(), w.cuda(), b.cuda(), 1, 1, 1, 8).cpu()
    ref = torch.nn.functional.conv2d(x, w, b, 1, 1)
    assert (out - ref).abs().max().item() < 2e-5


def test_adobe_lr_size_vs_torchvision_on_device():
    """180x320 (the Adobe LR size the encoder runs at): against torchvision's CUDA kernel, the implementation the oracle's
    shims run in place of the reference's unbuildable extension."""
    tv = pytest.importorskip("torchvision")
    from motif_b200.dcn_v2 import dcn_v2_conv

    x, off, m, w, b = [t.cuda() for t in _case(1, 64, 64, 180, 320, 8, 3.0, seed=4)]
    with torch.no_grad():
        out = dcn_v2_conv(x, off, m, w, b, 1, 1, 1, 8)
        ref = tv.ops.deform_conv2d(x, off, w, b, 1, 1, 1, m)
    assert (out - ref).abs().max().item() < 5e-5


def test_argument_checks_and_install():
    import sys
    import types

    from motif_b200 import dcn_v2

    x, off, m, w, b = _case(1, 8, 8, 4, 4, 1, 1.0, seed=1)
    with pytest.raises(NotImplementedError):
        dcn_v2.dcn_v2_conv(x, off, m, w, b, 1, 1, 1, 1)  # CPU tensors
    with torch.no_grad(), pytest.raises(NotImplementedError):
        dcn_v2.dcn_v2_conv(x.cuda(), off.cuda(), m.cuda(), w.cuda(), b.cuda(), 2, 1, 1, 1)  # stride 2
    with torch.no_grad(), pytest.raises(ValueError):
        dcn_v2.dcn_v2_conv(x.cuda(), off[:, :16].cuda(), m.cuda(), w.cuda(), b.cuda(), 1, 1, 1, 1)
    fake = types.ModuleType("models.modules.DCNv2.dcn_v2")
    fake.DCNv2 = object
    fake.dcn_v2_conv = None
    sys.modules["models.modules.DCNv2.dcn_v2"] = fake
    try:
        assert dcn_v2.install() >= 1 and fake.dcn_v2_conv is dcn_v2.dcn_v2_conv
    finally:
        del sys.modules["models.modules.DCNv2.dcn_v2"]
