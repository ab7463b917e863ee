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


@pytest.mark.parametrize("C,levels", [(128, 4), (256, 3), (128, 1)])
def test_pyramid_lookup_is_bit_equal_to_the_level_by_level_path(C, levels):
    """``motif_raft_corr_lookup_pyramid`` (every level of AlternateCorrBlock.__call__ in one launch: coords / 2^l, the stacking and
    the division by sqrt(C) inside the kernel) against the same block assembled from per-level ``alt_cuda_corr.forward`` calls and
    torch ops as corr.py:69-87 writes it; also with coordinates far outside the map and non-finite ones."""
    from motif_b200 import alt_cuda_corr

    g = torch.Generator().manual_seed(C + levels)
    B, H, W = 2, 20, 28
    f1, f2 = torch.randn(B, C, H, W, generator=g).cuda(), torch.randn(B, C, H, W, generator=g).cuda()
    ys, xs = torch.meshgrid(torch.arange(H).float(), torch.arange(W).float(), indexing="ij")
    coords = torch.stack([xs, ys])[None].repeat(B, 1, 1, 1) + torch.randn(B, 2, H, W, generator=g) * 6.0
    coords[0, :, 0, 0] = torch.tensor([1.0e7, -3.0])
    coords[1, :, 3, 5] = torch.tensor([float("nan"), 2.0])
    coords = coords.cuda()
    want = _alternate_block(f1, f2, coords, levels, 3)
    pyr = [f2]
    for _ in range(levels - 1):
        pyr.append(F.avg_pool2d(pyr[-1], 2, stride=2))
    got = alt_cuda_corr.forward_pyramid(f1.permute(0, 2, 3, 1).contiguous(), [p.permute(0, 2, 3, 1).contiguous() for p in pyr],
                                        coords.permute(0, 2, 3, 1).contiguous(), 3, normalize=True)
    assert got.shape == want.shape == (B, levels * 49, H, W)
    assert torch.equal(got, want)


def test_lookup_block_equals_the_oracle_block():
    """``raft_schedule.LookupBlock`` (AlternateCorrBlock with the layout changes hoisted out of the iteration loop) against the
    oracle's restatement of ``CorrBlock`` (pinned by the reference's own class, tests/golden/raft_corr.npz); called twice with
    different coordinates, as the RAFT iteration does."""
    from motif_b200.raft_schedule import LookupBlock

    g = torch.Generator().manual_seed(11)
    B, C, H, W = 2, 128, 24, 40
    f1, f2 = torch.randn(B, C, H, W, generator=g), torch.randn(B, C, H, W, generator=g)
    ys, xs = torch.meshgrid(torch.arange(H).float(), torch.arange(W).float(), indexing="ij")
    block = LookupBlock(f1.cuda(), f2.cuda(), radius=3)
    for k in range(2):
        coords = torch.stack([xs, ys])[None].repeat(B, 1, 1, 1) + torch.randn(B, 2, H, W, generator=g) * (2.0 + 3.0 * k)
        ref = raft_corr_ref.corr_block_lookup(f1, f2, coords, 4, 3)
        out = block(coords.cuda()).cpu()
        assert out.shape == ref.shape == (B, 4 * 49, H, W)
        assert (out - ref).abs().max().item() < 2e-4 * (1.0 + ref.abs().max().item())


def test_flow_two_pairs_runs_raft_sub_modules_on_cuda():
    """The schedule's glue on the device with a RAFT-shaped stand-in (the reference checkout is absent on the GPU box; bit-equality
    with the unmodified ``RAFT.forward`` is pinned on the CPU, tests/test_host_logic.py): the stand-in's own four-pair forward
    (written like raft.py:86-144, lookups through the oracle's block) and ``four_pair_flows`` must agree on the pairs 01 / 10."""
    import sys
    import types

    import torch.nn as nn
    import torch.nn.functional as F

    from motif_b200 import raft_schedule

    class Update(nn.Module):
        def __init__(self):
            super().__init__()
            self.c = nn.Conv2d(4 * 49 + 2 + 16 + 16, 16 + 2, 3, padding=1)

        def forward(self, net, inp, corr, flow):
            y = self.c(torch.cat([net, inp, corr, flow], 1))
            return torch.tanh(y[:, :16]), None, 0.5 * torch.tanh(y[:, 16:])

    class Enc(nn.Module):
        def __init__(self, o):
            super().__init__()
            self.c = nn.Conv2d(3, o, 8, stride=8)
            self.n = nn.InstanceNorm2d(o)

        def forward(self, x):
            is_list = isinstance(x, (list, tuple))
            if is_list:
                bd = x[0].shape[0]
                x = torch.cat(x, 0)
            y = self.n(self.c(x))
            return torch.split(y, [bd, bd], 0) if is_list else y

    mod = types.ModuleType("fake_raft_module")

    class FakeRaft(nn.Module):
        def __init__(self):
            super().__init__()
            self.args = types.SimpleNamespace(mixed_precision=False, alternate_corr=True, corr_radius=3)
            self.hidden_dim, self.context_dim = 16, 16
            self.fnet, self.cnet, self.update_block = Enc(128), Enc(32), Update()

        def initialize_flow(self, img):
            n, _, h, w = img.shape
            ys, xs = torch.meshgrid(torch.arange(h // 8, device=img.device).float(), torch.arange(w // 8, device=img.device).float(), indexing="ij")
            c = torch.stack([xs, ys])[None].repeat(n, 1, 1, 1)
            return c, c.clone()

        def upsample_flow(self, flow, mask):
            raise AssertionError("no mask in this stand-in")

        def forward(self, image1, image2, iters=12):
            image1, image2 = 2 * (image1 / 255.0) - 1.0, 2 * (image2 / 255.0) - 1.0
            fmap1, fmap2 = self.fnet([image1.contiguous(), image2.contiguous()])
            cnet = self.cnet(image1)
            net, inp = torch.tanh(cnet[:, :16]), torch.relu(cnet[:, 16:])
            coords0, coords1 = self.initialize_flow(image1)
            preds = []
            for _ in range(iters):
                corr = raft_corr_ref.corr_block_lookup(fmap1.cpu(), fmap2.cpu(), coords1.cpu(), 4, 3).to(image1.device)
                net, _m, delta = self.update_block(net, inp, corr, coords1 - coords0)
                coords1 = coords1 + delta
                preds.append(mod.upflow8(coords1 - coords0))
            return preds

    FakeRaft.__module__ = "fake_raft_module"
    mod.autocast = torch.cuda.amp.autocast
    mod.upflow8 = lambda flow: 8 * F.interpolate(flow, size=(8 * flow.shape[2], 8 * flow.shape[3]), mode="bilinear", align_corners=True)
    sys.modules["fake_raft_module"] = mod
    try:
        torch.manual_seed(2)
        raft = FakeRaft().cuda().eval()
        assert raft_schedule.is_raft(raft)
        low = torch.rand(2, 3, 24, 40)  # 24x40 features: the coarsest of the four levels is 3x5 (CorrBlock divides by H - 1)
        fr0, fr1 = [t.cuda() for t in F.interpolate(low, size=(192, 320), mode="bilinear", align_corners=False).split(1)]
        with torch.no_grad():
            four = raft(torch.cat([fr0, fr0, fr1, fr1]) * 255.0, torch.cat([fr0, fr1, fr0, fr1]) * 255.0, iters=3)[-1]
            sched = raft_schedule.four_pair_flows(raft, fr0, fr1, 3)
        assert sched.shape == four.shape == (4, 2, 192, 320) and torch.isfinite(four).all()
        assert not sched[0].any() and not sched[3].any()
        assert (sched[1:3] - four[1:3]).abs().max().item() < 1e-3 * (1.0 + four.abs().max().item())
    finally:
        del sys.modules["fake_raft_module"]
