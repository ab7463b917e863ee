"""GPU parity: RAFT's correlation lookup (models/core/corr.py:8-87) through motif_raft_corr_lookup / motif_b200.alt_cuda_corr."""
import pytest
import torch
import torch.nn.functional as F

from conftest import load_golden
from oracle import raft_corr_ref

pytestmark = pytest.mark.gpu


def _alternate_block(fmap1, fmap2, coords, num_levels, radius):
    """AlternateCorrBlock (corr.py:59-87) with alt_cuda_corr.forward answered by this repository."""
    from motif_b200 import alt_cuda_corr

    pyramid = [fmap2]
    for _ in range(num_levels - 1):
        pyramid.append(F.avg_pool2d(pyramid[-1], 2, stride=2))
    coords = coords.permute(0, 2, 3, 1)
    B, H, W, _ = coords.shape
    f1 = fmap1.permute(0, 2, 3, 1).contiguous()
    out = []
    for i in range(num_levels):
        f2 = pyramid[i].permute(0, 2, 3, 1).contiguous()
        corr, = alt_cuda_corr.forward(f1, f2, (coords / 2 ** i).reshape(B, 1, H, W, 2).contiguous(), radius)
        out.append(corr.squeeze(1))
    corr = torch.stack(out, dim=1).reshape(B, -1, H, W)
    return corr / torch.sqrt(torch.tensor(fmap1.shape[1]).float())


def test_lookup_vs_reference_corr_block_golden():
    g = load_golden("raft_corr")
    r = int(g["radius"][0])
    out = _alternate_block(g["fmap1"].cuda(), g["fmap2"].cuda(), g["coords"].cuda(), 4, r).cpu()
    assert out.shape == g["out"].shape
    assert (out - g["out"]).abs().max().item() < 2e-5  # values up to 3; fp32 summation order over C


@pytest.mark.parametrize("shape,radius", [((1, 128, 90, 160), 3), ((2, 256, 23, 31), 4), ((1, 6, 9, 7), 0), ((1, 30, 8, 8), 2)])
def test_lookup_vs_oracle(shape, radius):
    """RAFT-small at the HR size of the Adobe workload (720x1280 / 8), the full model's C = 256 / r = 4, odd channel
    counts (scalar path), r = 0."""
    from motif_b200 import alt_cuda_corr

    B, C, H, W = shape
    gen = torch.Generator().manual_seed(3)
    f1 = torch.randn(B, H, W, C, generator=gen)
    H2, W2 = max(H // 2, 2), max(W // 2, 2)
    f2 = torch.randn(B, H2, W2, C, generator=gen)
    ys, xs = torch.meshgrid(torch.arange(H).float(), torch.arange(W).float(), indexing="ij")
    coords = (torch.stack([xs, ys], -1)[None].repeat(B, 1, 1, 1) / 2 + torch.randn(B, H, W, 2, generator=gen) * 2.0).reshape(B, 1, H, W, 2)
    coords[0, 0, 0, 0] = torch.tensor([1.0, 1.0])
    coords[0, 0, H - 1, W - 1] = torch.tensor([-40.0, 1e9])
    ref = raft_corr_ref.alt_forward(f1, f2, coords, radius)
    out, = alt_cuda_corr.forward(f1.cuda(), f2.cuda(), coords.cuda(), radius)
    assert out.shape == ref.shape
    scale = ref.abs().max().item() + 1.0
    assert (out.cpu() - ref).abs().max().item() < 2e-6 * scale * (C ** 0.5)


def test_non_finite_coordinates_and_argument_checks():
    from motif_b200 import alt_cuda_corr

    f = torch.randn(1, 4, 4, 8).cuda()
    c = torch.full((1, 1, 4, 4, 2), float("nan")).cuda()
    out, = alt_cuda_corr.forward(f, f, c, 1)
    assert torch.equal(out, torch.zeros_like(out))
    with pytest.raises(NotImplementedError):
        alt_cuda_corr.forward(f.cpu(), f.cpu(), c.cpu(), 1)
    with pytest.raises(ValueError):
        alt_cuda_corr.forward(f, f, c[:, :, :2], 1)


def test_install_rebinds_the_reference_import():
    import sys
    import types

    from motif_b200 import alt_cuda_corr

    fake = types.ModuleType("models.core.corr")
    fake.AlternateCorrBlock = object
    fake.alt_cuda_corr = None
    sys.modules["models.core.corr"] = fake
    try:
        alt_cuda_corr.install()
        assert sys.modules["alt_cuda_corr"] is alt_cuda_corr and fake.alt_cuda_corr is alt_cuda_corr
    finally:
        del sys.modules["models.core.corr"]
        sys.modules.pop("alt_cuda_corr", None)
