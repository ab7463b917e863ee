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


@pytest.mark.parametrize("scale_x,scale_w,scale_m", [(1.0, 1.0, 1.0), (3.0e-5, 1.0, 1.0), (2.0e4, 1.0e-3, 1.0), (1.0, 50.0, 6.0)])
def test_tensor_core_path_ranges_and_ragged_tiles(scale_x, scale_w, scale_m):
    """The tcgen05 path (64 -> 64, 8 groups): two images whose 442 pixels do not fill the 128-pixel tiles (one tile straddles
    the batch boundary), no bias, and operand magnitudes far from 1 -- the power-of-two operand scaling must keep the two-piece
    fp16 split at fp32 accuracy whatever the ranges (relative gate)."""
    from motif_b200.dcn_v2 import dcn_v2_conv

    x, off, m, w, b = _case(2, 64, 64, 13, 17, 8, 2.0, seed=77)
    x, w, m = x * scale_x, w * scale_w, m * scale_m
    ref = dcn_v2_ref.dcn_v2_conv(x.double(), off.double(), m.double(), w.double(), torch.zeros(64, dtype=torch.float64), 1, 1, 1, 8).float()
    with torch.no_grad():
        out = dcn_v2_conv(x.cuda(), off.cuda(), m.cuda(), w.cuda(), None, 1, 1, 1, 8).cpu()
    assert torch.isfinite(out).all()
    err = (out - ref).abs().max().item() / ref.abs().max().item()
    assert err < 1e-5, err


def test_tensor_core_path_matches_cuda_core_path():
    """Same inputs through both kernels of motif_dcn_v2_fwd (the fused CUDA-core kernel serves every other shape)."""
    import subprocess
    import sys

    code = (
        "import torch, sys; sys.path.insert(0, '.'); from motif_b200.dcn_v2 import dcn_v2_conv\n"
        "g = torch.Generator().manual_seed(5)\n"
        "x = torch.randn(1, 64, 45, 80, generator=g); off = torch.randn(1, 144, 45, 80, generator=g) * 1.5\n"
        "m = torch.sigmoid(torch.randn(1, 72, 45, 80, generator=g)); w = torch.randn(64, 64, 3, 3, generator=g) / 24; b = torch.randn(64, generator=g)\n"
        "with torch.no_grad(): out = dcn_v2_conv(x.cuda(), off.cuda(), m.cuda(), w.cuda(), b.cuda(), 1, 1, 1, 8).cpu()\n"
        "torch.save(out, sys.argv[1])\n")
    import os
    import tempfile

    outs = []
    for simt in ("0", "1"):
        with tempfile.NamedTemporaryFile(suffix=".pt") as f:
            subprocess.run([sys.executable, "-c", code, f.name], check=True, env=dict(os.environ, MOTIF_DCN_SIMT=simt), cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
            outs.append(torch.load(f.name))
    assert (outs[0] - outs[1]).abs().max().item() < 2e-5 * (1.0 + outs[1].abs().max().item())
    assert not torch.equal(outs[0], outs[1])  # two different kernels did run


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
