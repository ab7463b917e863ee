"""PWC-Net cost volume, drop-in for the reference's ``OpticalFlow/correlation.py``.

``FunctionCorrelation(tensorFirst, tensorSecond)`` -> ``[B, 81, H, W]`` with
``out[b, 9*(dy+4)+(dx+4), y, x] = mean_c first[b,c,y,x] * second[b,c,y+dy,x+dx]`` (zeros outside).
Like the reference it ASSERTS contiguity (``correlation.py:302-303``) and is CUDA-only
(``:343-344``); unlike the reference it launches on the current stream of every call (the
reference captures the stream once at import time, ``:7-8``).
"""
from __future__ import annotations

import torch

from . import _lib


class _FunctionCorrelation(torch.autograd.Function):
    @staticmethod
    def forward(ctx, first, second):
        lib = _lib.load()
        _lib.require_cuda_f32("tensorFirst", first, 4)
        _lib.require_cuda_f32("tensorSecond", second, 4)
        assert first.is_contiguous() == True  # noqa: E712  (correlation.py:302)
        assert second.is_contiguous() == True  # noqa: E712
        assert first.shape == second.shape
        b, c, h, w = first.shape
        out = torch.empty((b, 81, h, w), dtype=torch.float32, device=first.device)
        with torch.cuda.device(first.device):
            rc = lib.motif_corr_fwd(first.data_ptr(), second.data_ptr(), out.data_ptr(), b, c, h, w, _lib.current_stream_ptr(first.device))
        _lib.check(rc, "motif_corr_fwd")
        return out

    @staticmethod
    def backward(ctx, gradOutput):
        raise NotImplementedError("motif_b200 implements the inference (forward) path only")


def FunctionCorrelation(tensorFirst, tensorSecond):
    return _FunctionCorrelation.apply(tensorFirst, tensorSecond)


class ModuleCorrelation(torch.nn.Module):
    def forward(self, tensorFirst, tensorSecond):
        return _FunctionCorrelation.apply(tensorFirst, tensorSecond)
