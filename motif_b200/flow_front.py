"""Reliability maps and flow-encoder input of ``LunaTokis.forward`` in one kernel (``Ours.py:562-578, 613-637``).

``flow_front(fr0, fr1, flow, g_filter)`` returns the tensor the reference hands to ``self.flow_process``:
``[2B, 14, H, W]`` = per reference frame ``r`` and frame pair ``j``: ``[flow / 20, psi_photo, psi_flow / 10, psi_var,
durations / 8]`` (``trans=False``, ``input_Z=True`` as shipped).  ``flow`` is the LR flow of the four pairs 00, 01, 10, 11
(``[4B, 2, H, W]``, pairs 00 and 11 already zeroed as ``Ours.py:553-555`` does).  CUDA fp32 only, like the operators.
"""
from __future__ import annotations

import torch

from . import _lib


def flow_front(fr0: torch.Tensor, fr1: torch.Tensor, flow: torch.Tensor, g_filter: torch.Tensor) -> torch.Tensor:
    lib = _lib.load()
    fr0, fr1, flow = fr0.contiguous(), fr1.contiguous(), flow.contiguous()  # the reference slices fr0 / fr1 out of a permuted clip
    _lib.require_cuda_f32("fr0", fr0, 4)
    _lib.require_cuda_f32("fr1", fr1, 4)
    _lib.require_cuda_f32("flow", flow, 4)
    B, C, H, W = fr0.shape
    if C != 3 or fr1.shape != fr0.shape or flow.shape != (4 * B, 2, H, W):
        raise ValueError(f"flow_front: fr0 {tuple(fr0.shape)}, fr1 {tuple(fr1.shape)}, flow {tuple(flow.shape)} do not fit [B,3,H,W] / [4B,2,H,W]")
    gf = g_filter.detach().to(device=fr0.device, dtype=torch.float32).reshape(-1).contiguous()
    if gf.numel() != 9:
        raise ValueError("flow_front: g_filter must hold 3x3 values (LunaTokis.g_filter)")
    out = torch.empty(2 * B, 14, H, W, dtype=torch.float32, device=fr0.device)
    with torch.cuda.device(fr0.device):
        rc = lib.motif_flow_front(fr0.data_ptr(), fr1.data_ptr(), flow.data_ptr(), gf.data_ptr(), out.data_ptr(), B, H, W, _lib.current_stream_ptr(fr0.device))
    _lib.check(rc, "motif_flow_front")
    return out
