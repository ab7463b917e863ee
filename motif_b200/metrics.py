"""Output path of the reference's evaluation loop (``test.py:187-235``) on the device.

``test.py`` crops the decoded frames to the ground-truth size, then forms the L1 loss, the BT.601 luma of both tensors
and the per-frame MSE / PSNR with a dozen eager kernels and full-size temporaries.  ``frame_metrics`` does the crop, the
luma and both reductions in one pass (``motif_frame_metrics``) and returns the same numbers; ``psnr_summary`` applies
``test.py:226-231`` to them.
"""
from __future__ import annotations

import math

import torch

from . import _lib


def frame_metrics(fake_H: torch.Tensor, real_H: torch.Tensor):
    """``fake_H [N, B, 3, HHp, WWp]`` (decoder output, possibly padded, ``test.py:192-199``) and ``real_H [B*N, 3, H, W]``
    (``model.real_H[:, 1:-1].reshape(b * n, 3, H, W)``, ``test.py:189``).  Frames are paired in flattened order exactly as
    ``test.py:196`` does (``fake_H[..., :H, :W].reshape(b * n, 3, H, W)``).  Returns ``(loss, mse [B*N])``:
    ``loss = mean |real - fake|`` over RGB (``test.py:201``), ``mse`` = per-frame mean squared luma difference (``:222-223``)."""
    lib = _lib.load()
    _lib.require_cuda_f32("fake_H", fake_H, 5)
    _lib.require_cuda_f32("real_H", real_H, 4)
    n, b, c, hp, wp = fake_H.shape
    f, c2, h, w = real_H.shape
    if c != 3 or c2 != 3 or f != n * b or hp < h or wp < w:
        raise ValueError(f"frame_metrics: shapes do not match: fake {tuple(fake_H.shape)}, real {tuple(real_H.shape)}")
    out = torch.empty(f, 2, dtype=torch.float64, device=fake_H.device)
    with torch.cuda.device(fake_H.device):
        rc = lib.motif_frame_metrics(fake_H.contiguous().data_ptr(), real_H.contiguous().data_ptr(), out.data_ptr(), f, hp, wp, h, w,
                                     _lib.current_stream_ptr(fake_H.device))
    _lib.check(rc, "motif_frame_metrics")
    loss = out[:, 0].sum() / (f * 3 * h * w)
    mse = out[:, 1] / (h * w)
    return loss, mse


def psnr_summary(mse: torch.Tensor):
    """``test.py:224-231``: PSNR of the first frame, mean PSNR of the inner frames, of the centre frame, and the weighted mean."""
    m = mse.detach().double().cpu()
    n = len(m)
    p = 10.0 * torch.log10(1.0 / m)
    anchor = p[0].item()
    inter = p[1:-1].mean().item() if n > 2 else float("nan")
    center = p[n // 2].item()
    overall = (anchor * 1 + inter * (n - 2)) / (n - 1) if n > 2 else anchor
    return {"anchor": anchor, "inter": inter, "center": center, "psnr": overall, "all": p.tolist()}
