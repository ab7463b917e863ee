"""Count splat, drop-in for the reference's ``models/softsplat_count_cp.py``.

``FunctionSoftsplat(tenInput, tenFlow)`` ignores the values of ``tenInput`` (the reference
replaces it by ones, ``softsplat_count_cp.py:163-165``) and returns, detached, the number of
source pixels whose 2x2 footprint covers each destination pixel (unweighted, ``:39-50``) as
``[N, 1, H, W]`` fp32.  Exact (integer-valued).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _lib


def FunctionSoftsplat(tenInput, tenFlow):
    lib = _lib.load()
    _lib.require_cuda_f32("tenInput", tenInput, 4)
    _lib.require_cuda_f32("tenFlow", tenFlow, 4)
    n, _, h, w = tenInput.shape
    assert tenFlow.shape[1] == 2
    assert tenFlow.shape[2] == h and tenFlow.shape[3] == w and tenFlow.shape[0] == n
    tenFlow = tenFlow.detach().contiguous()
    out = torch.empty((n, 1, h, w), dtype=torch.float32, device=tenFlow.device)
    with torch.cuda.device(tenFlow.device):
        rc = lib.motif_splat_count_fwd(tenFlow.data_ptr(), out.data_ptr(), n, h, w, _lib.current_stream_ptr(tenFlow.device))
    _lib.check(rc, "motif_splat_count_fwd")
    return out


class Softsplat_Count(nn.Module):
    def forward(self, img, flow):
        return FunctionSoftsplat(img, flow)
