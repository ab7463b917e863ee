"""Forward splatting, drop-in for the reference's ``models/softsplat_cp.py``.

``FunctionSoftsplat(tenInput, tenFlow, tenMetric, strType)`` keeps the reference contract
(``softsplat_cp.py:320-347``): it returns the tuple ``(tenOutput[:, :-1], tenOutput[:, -1:])``,
both UN-normalised views of one ``[N, C+1, H, W]`` buffer (the reference leaves the division to
its caller, ``Ours.py:811-814``).  The metric product (``in * metric`` / ``in * exp(metric)``)
is fused into the kernel instead of being materialised by torch first.

Deviation (documented in DESIGN.md): ``strType='summation'`` raises ``UnboundLocalError`` in the
reference (``tenNormalize`` is never bound, ``softsplat_cp.py:337-346``); here it returns
``(tenOutput, None)``.

Forward only (the reference path under ``torch.no_grad``, ``VideoSR_base_model.py:171``):
``backward`` raises.
"""
from __future__ import annotations

import ctypes

import torch
import torch.nn as nn

from . import _lib

def _workspace(device, nbytes):
    """Scratch of one call, from torch's caching allocator on the current stream (stream-ordered reuse is the
    allocator's job; nothing is cached here, so short-lived streams leak nothing)."""
    return torch.empty(max(nbytes, 16), dtype=torch.uint8, device=device)


def _splat(tenInput, tenFlow, tenMetric, mode, atomic=False):
    lib = _lib.load()
    _lib.require_cuda_f32("tenInput", tenInput, 4)
    _lib.require_cuda_f32("tenFlow", tenFlow, 4)
    n, c, h, w = tenInput.shape
    # softsplat_cp.py:228-230
    assert tenFlow.shape[1] == 2
    assert tenFlow.shape[2] == h and tenFlow.shape[3] == w and tenFlow.shape[0] == n
    tenInput = tenInput.contiguous()
    tenFlow = tenFlow.contiguous()
    metric_ptr = None
    if mode >= 2:
        _lib.require_cuda_f32("tenMetric", tenMetric, 4)
        assert tenMetric.shape == (n, 1, h, w)
        tenMetric = tenMetric.contiguous()
        metric_ptr = tenMetric.data_ptr()
    c_out = c if mode == 0 else c + 1
    out = torch.empty((n, c_out, h, w), dtype=torch.float32, device=tenInput.device)
    with torch.cuda.device(tenInput.device):
        stream = _lib.current_stream_ptr(tenInput.device)
        if atomic:
            rc = lib.motif_splat_fwd_atomic(tenInput.data_ptr(), tenFlow.data_ptr(), metric_ptr, out.data_ptr(), n, c, h, w, mode, stream)
        else:
            nbytes = lib.motif_splat_workspace_bytes(n, h, w)
            ws = _workspace(tenInput.device, nbytes)
            rc = lib.motif_splat_fwd(tenInput.data_ptr(), tenFlow.data_ptr(), metric_ptr, out.data_ptr(), n, c, h, w, mode,
                                     ws.data_ptr(), ctypes.c_size_t(ws.numel()), stream)
    _lib.check(rc, "motif_splat_fwd")
    return out


class _FunctionSoftsplat(torch.autograd.Function):
    @staticmethod
    def forward(ctx, tenInput, tenFlow, tenMetric, mode):
        return _splat(tenInput, tenFlow, tenMetric, mode)

    @staticmethod
    def backward(ctx, gradOutput):
        raise NotImplementedError("motif_b200 implements the inference (forward) path only")


def FunctionSoftsplat(tenInput, tenFlow, tenMetric, strType):
    assert tenMetric is None or tenMetric.shape[1] == 1
    assert strType in ["summation", "average", "linear", "softmax"]
    mode = _lib.SPLAT_MODES[strType]
    tenOutput = _FunctionSoftsplat.apply(tenInput, tenFlow, tenMetric, mode)
    if strType == "summation":
        return tenOutput, None
    tenNormalize = tenOutput[:, -1:, :, :]
    tenOutput = tenOutput[:, :-1, :, :]
    return tenOutput, tenNormalize


class Softsplat(nn.Module):
    def __init__(self, strType="softmax"):
        super().__init__()
        self.strType = strType

    def forward(self, img, flow, z):
        return FunctionSoftsplat(img, flow, z, self.strType)
