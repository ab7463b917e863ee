"""CPU restatement of the reliability-map front end of ``LunaTokis.forward`` (TEST INFRASTRUCTURE ONLY).

``models/modules/Ours.py:562-578`` (psi_photo, psi_flow, psi_var through ``BackWarp``, ``:892-923``, and the 3x3
gaussian ``conv3d``) and ``:613-637`` (the concatenation that feeds ``flow_process``; ``trans=False``,
``input_Z=True`` as shipped).  Pinned by ``tests/golden/front_*.npz``: the input of ``flow_process`` captured by a
forward pre-hook while the unmodified reference forward ran (``oracle/make_golden.py``).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def lr_flow_from_hr(flow_hr: torch.Tensor, B: int, H: int, W: int) -> torch.Tensor:
    """``Ours.py:549-556``: bilinear down-sampling of the four HR flows, rescaling, pairs 00 and 11 set to zero."""
    HH = flow_hr.shape[-2]
    flow = F.interpolate(flow_hr, size=(H, W), mode="bilinear", align_corners=False) * (H / HH)
    flow = flow.reshape(4, B, 2, H, W).clone()
    flow[0] *= 0.0
    flow[3] *= 0.0
    return flow.reshape(4 * B, 2, H, W)


def back_warp(img: torch.Tensor, flow: torch.Tensor) -> torch.Tensor:
    """``BackWarp(clip=True).forward`` (``Ours.py:900-923``): note the normalisation by ``w`` (not ``w - 1``) fed to a
    ``grid_sample(align_corners=True, padding_mode='border')``."""
    b, _, h, w = flow.shape
    grid_y, grid_x = torch.meshgrid(torch.arange(h), torch.arange(w), indexing="ij")
    x = grid_x.unsqueeze(0).expand(b, h, w).float() + flow[:, 0]
    y = grid_y.unsqueeze(0).expand(b, h, w).float() + flow[:, 1]
    x = (x / w) * 2 - 1
    y = (y / h) * 2 - 1
    return F.grid_sample(img, torch.stack((x, y), dim=-1), mode="bilinear", align_corners=True, padding_mode="border")


def flow_front(fr0: torch.Tensor, fr1: torch.Tensor, flow: torch.Tensor, g_filter: torch.Tensor) -> torch.Tensor:
    """fr0, fr1 ``[B,3,H,W]``; flow ``[4B,2,H,W]`` (pairs 00, 01, 10, 11; LR pixels); g_filter ``[1,1,1,3,3]``.
    Returns the ``flow_process`` input ``[2B,14,H,W]``."""
    B, _, H, W = fr0.shape
    warped = back_warp(torch.cat([fr0, fr1, fr0, fr1], 0), flow)  # Ours.py:567
    psi_photo = F.l1_loss(input=torch.cat([fr0, fr0, fr1, fr1], 0), target=warped, reduction="none").mean(1)
    f4 = flow.reshape(4, B, 2, H, W)
    warped = back_warp(-torch.cat([f4[0], f4[2], f4[1], f4[3]], 0), flow)  # Ours.py:570-575
    psi_flow = F.l1_loss(input=flow, target=warped, reduction="none").mean(1)
    f = flow.reshape(4 * B, -1, H, W)
    sq_mean, mean_sq = torch.split(  # Ours.py:577-581
        F.conv3d(F.pad(torch.cat([f ** 2, f], 1), (1, 1, 1, 1), mode="reflect").unsqueeze(1), g_filter.reshape(1, 1, 1, 3, 3)).squeeze(1), 2, dim=1)
    psi_var = (sq_mean - mean_sq ** 2).clip(1e-9, None).sqrt().mean(1)
    psies = torch.stack([psi_photo, psi_flow / 10.0, psi_var], dim=1)
    durations = torch.tensor([[0, 0], [0, 8], [8, 0], [8, 8]], dtype=torch.float32).unsqueeze(1)  # Ours.py:615-621
    return torch.cat(  # Ours.py:625-631
        (
            (flow / 20.0).reshape(2, 2, B, -1, H, W).permute(0, 2, 1, 3, 4, 5).reshape(2 * B, 2, -1, H, W),
            psies.reshape(2, 2, B, -1, H, W).permute(0, 2, 1, 3, 4, 5).reshape(2 * B, 2, -1, H, W),
            durations.reshape(2, 4, 1, 1).unsqueeze(1).repeat(1, B, 1, H, W).reshape(2 * B, 2, 2, H, W) / 8.0,
        ),
        dim=2,
    ).reshape(2 * B, -1, H, W)
