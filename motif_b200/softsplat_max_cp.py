"""Max splat, drop-in for the reference's ``models/softsplat_max_cp.py``.

``FunctionSoftsplat(tenInput, tenFlow)`` returns ``max(1.0, max over contributions of
in * bilinear weight)`` per destination cell: the reference initialises its output to ONES
(``softsplat_max_cp.py:254``) and reduces with ``atomicMaxFloat`` (``:13-18``).  Deterministic.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _lib


class _FunctionSoftsplat(torch.autograd.Function):
    @staticmethod
    def forward(ctx, tenInput, tenFlow):
        lib = _lib.load()
        _lib.require_cuda_f32("tenInput", tenInput, 4)
        _lib.require_cuda_f32("tenFlow", tenFlow, 4)
        n, c, h, w = tenInput.shape
        assert tenFlow.shape[1] == 2
        assert tenFlow.shape[2] == h and tenFlow.shape[3] == w and tenFlow.shape[0] == n
        tenInput = tenInput.contiguous()
        tenFlow = tenFlow.contiguous()
        out = torch.empty_like(tenInput)
        with torch.cuda.device(tenInput.device):
            rc = lib.motif_splat_max_fwd(tenInput.data_ptr(), tenFlow.data_ptr(), out.data_ptr(), n, c, h, w,
                                         _lib.current_stream_ptr(tenInput.device))
        _lib.check(rc, "motif_splat_max_fwd")
        return out

    @staticmethod
    def backward(ctx, gradOutput):
        raise NotImplementedError("motif_b200 implements the inference (forward) path only")


def FunctionSoftsplat(tenInput, tenFlow):
    return _FunctionSoftsplat.apply(tenInput, tenFlow)


class Softsplat_Max(nn.Module):
    def forward(self, img, flow):
        return FunctionSoftsplat(img, flow)
