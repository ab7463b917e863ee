"""Source-available stand-in for RAFT's ``alt_cuda_corr`` extension, which the reference imports (``models/core/corr.py:5``)
and calls from ``AlternateCorrBlock`` (``corr.py:69-87``) but ships only as a CPython-3.7 binary.

``forward(fmap1 [B,H,W,C], fmap2 [B,H2,W2,C], coords [B,1,H,W,2], r)`` returns ``[corr]`` with ``corr [B,1,(2r+1)^2,H,W]``: the
bilinear (zero-padded) lookup of the level's correlation volume in the window order of the in-repo ``CorrBlock``
(``corr.py:8-56``), not yet divided by ``sqrt(C)`` -- exactly what ``AlternateCorrBlock`` expects.  ``install()`` makes the
reference's ``import alt_cuda_corr`` resolve to this module (and patches an already imported ``models.core.corr``).
CUDA fp32 only; forward only.
"""
from __future__ import annotations

import sys

import torch

from . import _lib


def forward(fmap1: torch.Tensor, fmap2: torch.Tensor, coords: torch.Tensor, r: int):
    lib = _lib.load()
    _lib.require_cuda_f32("fmap1", fmap1, 4)
    _lib.require_cuda_f32("fmap2", fmap2, 4)
    _lib.require_cuda_f32("coords", coords, 5)
    B, H, W, C = fmap1.shape
    B2, H2, W2, C2 = fmap2.shape
    if B2 != B or C2 != C or coords.shape != (B, 1, H, W, 2):
        raise ValueError(f"alt_cuda_corr.forward: fmap1 {tuple(fmap1.shape)}, fmap2 {tuple(fmap2.shape)}, coords {tuple(coords.shape)} do not fit")
    fmap1, fmap2, coords = fmap1.contiguous(), fmap2.contiguous(), coords.contiguous()
    n = (2 * int(r) + 1) ** 2
    out = torch.empty(B, 1, n, H, W, dtype=torch.float32, device=fmap1.device)
    with torch.cuda.device(fmap1.device):
        rc = lib.motif_raft_corr_lookup(fmap1.data_ptr(), fmap2.data_ptr(), coords.data_ptr(), out.data_ptr(), B, H, W, H2, W2, C, int(r),
                                        _lib.current_stream_ptr(fmap1.device))
    _lib.check(rc, "motif_raft_corr_lookup")
    return [out]


def backward(*_args, **_kwargs):
    raise NotImplementedError("motif_b200 implements the inference (forward) path only")


def install():
    """Let ``import alt_cuda_corr`` (``models/core/corr.py:5``) resolve to this module; also rebinds the name inside a
    ``...core.corr`` module that was imported earlier with another (or a stub) ``alt_cuda_corr``."""
    me = sys.modules[__name__]
    sys.modules["alt_cuda_corr"] = me
    for name, mod in list(sys.modules.items()):
        if mod is not None and name.endswith("core.corr") and hasattr(mod, "AlternateCorrBlock"):
            mod.alt_cuda_corr = me
    return me
