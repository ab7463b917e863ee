"""CPU restatement of the evaluation loop's output path (TEST INFRASTRUCTURE ONLY): ``test.py:187-235`` -- crop to the
ground-truth size, L1 loss, BT.601 luma, per-frame MSE -- with the same torch operators in the same order.  The lines are
inline in the reference's ``main()`` loop (nothing importable), so parity is pinned by restating them and by an independent
float64 evaluation in the tests, not by a reference fixture."""
import torch


def frame_metrics(fake_H: torch.Tensor, real_H: torch.Tensor):
    n, b = fake_H.shape[:2]
    H, W = real_H.shape[-2:]
    real_H = real_H.clone().float()
    fake_H = fake_H[:, :, :, 0:H, 0:W].reshape(b * n, 3, H, W).clone().float()   # test.py:196
    loss = abs(real_H - fake_H).mean().item()                                      # :201
    real_H *= 255.0                                                                # :212-217
    fake_H *= 255.0
    real_H = (real_H[:, 0] * 65.481 + real_H[:, 1] * 128.553 + real_H[:, 2] * 24.966) / 255.0 + 16.0
    fake_H = (fake_H[:, 0] * 65.481 + fake_H[:, 1] * 128.553 + fake_H[:, 2] * 24.966) / 255.0 + 16.0
    real_H /= 255.0
    fake_H /= 255.0
    mse = (real_H - fake_H) ** 2                                                   # :222-223
    mse = torch.mean(mse.contiguous().view(b * n, -1), dim=1)
    return loss, mse
