"""GPU parity: the evaluation loop's output path (test.py:187-235) in one kernel against its restatement and float64."""
import pytest
import torch

from oracle import metrics_ref

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,b,h,w,hp,wp", [(3, 1, 37, 53, 40, 56), (7, 1, 720, 1280, 720, 1280), (2, 2, 16, 20, 16, 20)])
def test_frame_metrics_match_the_reference_lines(n, b, h, w, hp, wp):
    from motif_b200 import metrics

    g = torch.Generator().manual_seed(n * 100 + h)
    fake = torch.rand(n, b, 3, hp, wp, generator=g)
    real = (fake[:, :, :, :h, :w].reshape(n * b, 3, h, w) + 0.05 * torch.randn(n * b, 3, h, w, generator=g)).clamp(0, 1)
    r_loss, r_mse = metrics_ref.frame_metrics(fake, real)
    loss, mse = metrics.frame_metrics(fake.cuda(), real.cuda())
    assert abs(loss.item() - r_loss) < 1e-6 * max(r_loss, 1e-3)
    assert torch.allclose(mse.cpu().float(), r_mse, rtol=2e-5, atol=1e-9)
    # independent float64 evaluation of the same formula
    f64 = fake[:, :, :, :h, :w].reshape(n * b, 3, h, w).double()
    y = lambda t: (((t[:, 0] * 255 * 65.481 + t[:, 1] * 255 * 128.553 + t[:, 2] * 255 * 24.966) / 255 + 16) / 255)
    m64 = ((y(real.double()) - y(f64)) ** 2).flatten(1).mean(1)
    assert torch.allclose(mse.cpu(), m64, rtol=1e-4)
    s = metrics.psnr_summary(mse)
    p = 10 * torch.log10(1 / m64)
    assert abs(s["anchor"] - p[0].item()) < 1e-3 and abs(s["center"] - p[len(p) // 2].item()) < 1e-3
    if len(p) > 2:
        assert abs(s["psnr"] - (p[0].item() + p[1:-1].mean().item() * (len(p) - 2)) / (len(p) - 1)) < 1e-3


def test_frame_metrics_refuse_cpu_and_bad_shapes():
    from motif_b200 import metrics

    with pytest.raises(NotImplementedError):
        metrics.frame_metrics(torch.zeros(1, 1, 3, 4, 4), torch.zeros(1, 3, 4, 4))
    with pytest.raises(ValueError):
        metrics.frame_metrics(torch.zeros(1, 1, 3, 4, 4).cuda(), torch.zeros(1, 3, 8, 4).cuda())
