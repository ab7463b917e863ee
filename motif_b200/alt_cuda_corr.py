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


def forward_pyramid(fmap1: torch.Tensor, fmap2_levels, coords: torch.Tensor, r: int, normalize: bool = True):
    """Every level of ``AlternateCorrBlock.__call__`` (``corr.py:69-87``) in one launch: ``fmap1 [B,H,W,C]``, ``fmap2_levels[l]
    [B,H2_l,W2_l,C]`` (``fmap2`` pooled ``l`` times), ``coords [B,H,W,2]`` the LEVEL-0 coordinates -> the stacked
    ``[B, L*(2r+1)^2, H, W]`` tensor, divided by ``sqrt(C)`` when ``normalize`` (``corr.py:87``).  ``r <= 3``, ``C`` 128 or 256, at
    most four levels (RAFT's configuration); other shapes go level by level through ``forward``."""
    import ctypes

    lib = _lib.load()
    _lib.require_cuda_f32("fmap1", fmap1, 4)
    _lib.require_cuda_f32("coords", coords, 4)
    B, H, W, C = fmap1.shape
    L = len(fmap2_levels)
    if coords.shape != (B, H, W, 2) or not 1 <= L <= 4 or C not in (128, 256) or not 0 <= int(r) <= 3:
        raise ValueError(f"alt_cuda_corr.forward_pyramid: fmap1 {tuple(fmap1.shape)}, coords {tuple(coords.shape)}, {L} levels, r={r} not supported")
    fmap1, coords = fmap1.contiguous(), coords.contiguous()
    levels = []
    for f2 in fmap2_levels:
        _lib.require_cuda_f32("fmap2", f2, 4)
        if f2.shape[0] != B or f2.shape[3] != C:
            raise ValueError(f"alt_cuda_corr.forward_pyramid: level {tuple(f2.shape)} does not fit fmap1 {tuple(fmap1.shape)}")
        levels.append(f2.contiguous())
    n = (2 * int(r) + 1) ** 2
    out = torch.empty(B, L * n, H, W, dtype=torch.float32, device=fmap1.device)
    ptrs = (ctypes.c_void_p * L)(*[t.data_ptr() for t in levels])
    h2 = (ctypes.c_int * L)(*[t.shape[1] for t in levels])
    w2 = (ctypes.c_int * L)(*[t.shape[2] for t in levels])
    with torch.cuda.device(fmap1.device):
        rc = lib.motif_raft_corr_lookup_pyramid(fmap1.data_ptr(), ptrs, h2, w2, L, coords.data_ptr(), out.data_ptr(), B, H, W, C, int(r),
                                                1 if normalize else 0, _lib.current_stream_ptr(fmap1.device))
    _lib.check(rc, "motif_raft_corr_lookup_pyramid")
    return out


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
