"""GPU parity: the reliability-map front end (Ours.py:562-578, 613-637) through the C ABI (motif_flow_front)."""
import pytest
import torch

from conftest import load_golden
from oracle import flow_front_ref

pytestmark = pytest.mark.gpu

TOL = 2e-6  # fp32 re-association of the bilinear weights and of the 3x3 window sum; values are O(1)


@pytest.mark.parametrize("case", ["front_raft", "front_smooth_b2"])
def test_flow_front_vs_reference_golden(case):
    from motif_b200.flow_front import flow_front

    g = load_golden(case)
    x = g["x"]
    B, H, W = x.shape[0], x.shape[-2], x.shape[-1]
    flow = flow_front_ref.lr_flow_from_hr(g["flow_hr"], B, H, W)
    out = flow_front(x[:, 0].cuda(), x[:, 1].cuda(), flow.cuda(), g["g_filter"].cuda())
    assert out.shape == g["flow_process_in"].shape
    d = (out.cpu() - g["flow_process_in"]).abs()
    assert d.max().item() < TOL, [d[:, c].max().item() for c in range(14)]
    # channels that are pure data movement are bit-exact: flow / 20 and the durations
    for c in (0, 1, 5, 6, 7, 8, 12, 13):
        assert torch.equal(out[:, c].cpu(), g["flow_process_in"][:, c]), c


@pytest.mark.parametrize("shape,sigma", [((1, 180, 320), 3.0), ((2, 45, 80), 40.0), ((1, 2, 2), 1.0), ((1, 7, 5), 0.0)])
def test_flow_front_vs_oracle(shape, sigma):
    """Adobe-sized LR pair, far out-of-frame flows (border clamping), the smallest legal size and zero flow."""
    from motif_b200.flow_front import flow_front

    B, H, W = shape
    gen = torch.Generator().manual_seed(11)
    fr0, fr1 = torch.rand(B, 3, H, W, generator=gen), torch.rand(B, 3, H, W, generator=gen)
    flow = torch.randn(4 * B, 2, H, W, generator=gen) * sigma
    flow[:B] = 0.0
    flow[3 * B:] = 0.0
    flow[B, :, 0, 0] = torch.tensor([float(W), -float(H)])  # exactly out of frame
    flow[B, :, H - 1, W - 1] = torch.tensor([0.5, 0.5])
    gf = torch.tensor([[1, 2, 1], [2, 4, 2], [1, 2, 1]], dtype=torch.float32) / 16.0
    ref = flow_front_ref.flow_front(fr0, fr1, flow, gf)
    out = flow_front(fr0.cuda(), fr1.cuda(), flow.cuda(), gf.cuda()).cpu()
    scale = max(1.0, sigma)  # psi_flow and psi_var grow with the flow magnitude
    assert (out - ref).abs().max().item() < TOL * scale * 4


def test_flow_front_argument_checks():
    from motif_b200.flow_front import flow_front

    fr = torch.rand(1, 3, 8, 8)
    with pytest.raises(NotImplementedError):
        flow_front(fr, fr, torch.zeros(4, 2, 8, 8), torch.ones(3, 3))
    with pytest.raises(ValueError):
        flow_front(fr.cuda(), fr.cuda(), torch.zeros(3, 2, 8, 8).cuda(), torch.ones(3, 3).cuda())
