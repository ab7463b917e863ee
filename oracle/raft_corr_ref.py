"""CPU restatement of RAFT's correlation lookup as the reference runs it (TEST INFRASTRUCTURE ONLY).

The shipped model builds RAFT-small with ``alternate_corr=True`` (``models/modules/Ours.py:417-430``), whose lookup calls the
binary-only CUDA module ``alt_cuda_corr`` (``models/core/corr.py:5, 82``; the ``.so`` is not in the checkout, its source
is princeton-vl/RAFT's ``alt_cuda_corr``, no version pinned).  The in-repo definition of the same quantity is ``CorrBlock``
(``models/core/corr.py:8-56``): all-pairs correlation, ``avg_pool2d`` pyramid, ``bilinear_sampler`` lookup
(``models/core/utils/utils.py:57-70``: ``grid_sample(align_corners=True)``, zero padding) in a (2r+1)^2 window whose FIRST
index moves x (``delta = stack(meshgrid(dy, dx))`` is added to (x, y) coordinates, ``corr.py:36-41``).  Average pooling
commutes with the (linear) correlation, so pooling ``fmap2`` -- what ``AlternateCorrBlock`` does, ``corr.py:64-67`` -- gives
the same pyramid.  Pinned by ``tests/golden/raft_corr.npz``: the output of the reference's own ``CorrBlock`` class.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def _bilinear_sampler(img, coords):
    H, W = img.shape[-2:]
    xgrid, ygrid = coords.split([1, 1], dim=-1)
    xgrid = 2 * xgrid / (W - 1) - 1
    ygrid = 2 * ygrid / (H - 1) - 1
    return F.grid_sample(img, torch.cat([xgrid, ygrid], dim=-1), align_corners=True)


def corr_block_lookup(fmap1: torch.Tensor, fmap2: torch.Tensor, coords: torch.Tensor, num_levels: int = 4, radius: int = 3) -> torch.Tensor:
    """``CorrBlock(fmap1, fmap2, num_levels, radius)(coords)``: fmaps ``[B,C,H,W]``, coords ``[B,2,H,W]`` (x, y) ->
    ``[B, num_levels * (2r+1)^2, H, W]`` (already divided by ``sqrt(C)``)."""
    B, C, H, W = fmap1.shape
    corr = torch.matmul(fmap1.view(B, C, H * W).transpose(1, 2), fmap2.view(B, C, H * W)).view(B, H, W, 1, H, W)
    corr = corr / torch.sqrt(torch.tensor(C).float())
    corr = corr.reshape(B * H * W, 1, H, W)
    pyramid = [corr]
    for _ in range(num_levels - 1):
        corr = F.avg_pool2d(corr, 2, stride=2)
        pyramid.append(corr)
    r = radius
    coords = coords.permute(0, 2, 3, 1)
    out = []
    for i in range(num_levels):
        dx = torch.linspace(-r, r, 2 * r + 1)
        dy = torch.linspace(-r, r, 2 * r + 1)
        delta = torch.stack(torch.meshgrid(dy, dx, indexing="ij"), dim=-1)
        centroid = coords.reshape(B * H * W, 1, 1, 2) / 2 ** i
        out.append(_bilinear_sampler(pyramid[i], centroid + delta.view(1, 2 * r + 1, 2 * r + 1, 2)).view(B, H, W, -1))
    return torch.cat(out, dim=-1).permute(0, 3, 1, 2).contiguous().float()


def alt_forward(fmap1_nhwc: torch.Tensor, fmap2_nhwc: torch.Tensor, coords: torch.Tensor, r: int) -> torch.Tensor:
    """What ``alt_cuda_corr.forward(fmap1 [B,H,W,C], fmap2 [B,H2,W2,C], coords [B,1,H,W,2], r)[0]`` must return for
    ``AlternateCorrBlock`` (``corr.py:69-87``) to equal ``CorrBlock``: ``[B, 1, (2r+1)^2, H, W]``, NOT yet divided by
    ``sqrt(C)`` (the caller divides, ``corr.py:87``).  Direct evaluation: one level, zero padding outside ``fmap2``."""
    B, H, W, C = fmap1_nhwc.shape
    H2, W2 = fmap2_nhwc.shape[1:3]
    corr = torch.einsum("bhwc,byxc->bhwyx", fmap1_nhwc, fmap2_nhwc).reshape(B * H * W, 1, H2, W2)
    dx = torch.linspace(-r, r, 2 * r + 1)
    delta = torch.stack(torch.meshgrid(dx, dx, indexing="ij"), dim=-1)
    centroid = coords.reshape(B * H * W, 1, 1, 2)
    s = _bilinear_sampler(corr, centroid + delta.view(1, 2 * r + 1, 2 * r + 1, 2)).view(B, H, W, -1)
    return s.permute(0, 3, 1, 2).unsqueeze(1).contiguous()


def alternate_corr_block_lookup(fmap1, fmap2, coords, num_levels=4, radius=3):
    """``AlternateCorrBlock(fmap1, fmap2, num_levels, radius)(coords)`` (``corr.py:59-87``) on top of ``alt_forward``."""
    pyramid = [(fmap1, fmap2)]
    for _ in range(num_levels - 1):  # the reference pools once more than it uses (corr.py:65-68)
        fmap1 = F.avg_pool2d(fmap1, 2, stride=2)
        fmap2 = F.avg_pool2d(fmap2, 2, stride=2)
        pyramid.append((fmap1, fmap2))
    coords = coords.permute(0, 2, 3, 1)
    B, H, W, _ = coords.shape
    dim = pyramid[0][0].shape[1]
    out = []
    for i in range(num_levels):
        f1 = pyramid[0][0].permute(0, 2, 3, 1).contiguous()
        f2 = pyramid[i][1].permute(0, 2, 3, 1).contiguous()
        out.append(alt_forward(f1, f2, (coords / 2 ** i).reshape(B, 1, H, W, 2).contiguous(), radius).squeeze(1))
    corr = torch.stack(out, dim=1).reshape(B, -1, H, W)
    return corr / torch.sqrt(torch.tensor(dim).float())
