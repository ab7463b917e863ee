"""Drop-in for the reference's DCNv2 extension call (``models/modules/DCNv2/dcn_v2.py:13-47``).

``dcn_v2_conv(input, offset, mask, weight, bias, stride, padding, dilation, deformable_groups)`` with the reference's
argument order; 3x3 kernels with stride 1, padding 1, dilation 1 (every call site of the model, ``Ours.py:53-172``) run in
``libmotif_b200.so``.  ``install()`` rebinds the name inside an imported ``...DCNv2.dcn_v2`` module, whose ``DCN`` /
``DCN_sep`` classes then run unmodified -- the reference's own ``_ext`` needs ``THC/THC.h`` and does not build on torch 2.x.
CUDA fp32 only; forward only.
"""
from __future__ import annotations

import sys

import torch

from . import _lib


def _pair(v):
    return (int(v), int(v)) if not isinstance(v, (tuple, list)) else (int(v[0]), int(v[1]))


def dcn_v2_conv(input, offset, mask, weight, bias, stride, padding, dilation, deformable_groups):
    lib = _lib.load()
    for nm, t in (("input", input), ("offset", offset), ("mask", mask), ("weight", weight)):
        _lib.require_cuda_f32(nm, t, 4)
    if torch.is_grad_enabled() and any(t.requires_grad for t in (input, offset, mask, weight)):
        raise NotImplementedError("motif_b200 implements the inference (forward) path only; call under torch.no_grad()")
    if _pair(stride) != (1, 1) or _pair(padding) != (1, 1) or _pair(dilation) != (1, 1) or tuple(weight.shape[2:]) != (3, 3):
        raise NotImplementedError("motif_b200.dcn_v2_conv: 3x3 kernels with stride 1, padding 1, dilation 1 (the model's configuration)")
    B, Cin, H, W = input.shape
    Cout = weight.shape[0]
    dg = int(deformable_groups)
    if weight.shape[1] != Cin or offset.shape != (B, dg * 18, H, W) or mask.shape != (B, dg * 9, H, W):
        raise ValueError(f"dcn_v2_conv: input {tuple(input.shape)}, offset {tuple(offset.shape)}, mask {tuple(mask.shape)}, weight {tuple(weight.shape)} do not fit")
    input, offset, mask, weight = input.contiguous(), offset.contiguous(), mask.contiguous(), weight.contiguous()
    if bias is not None:
        _lib.require_cuda_f32("bias", bias, 1)
        bias = bias.contiguous()
    out = torch.empty(B, Cout, H, W, dtype=torch.float32, device=input.device)
    with torch.cuda.device(input.device):
        rc = lib.motif_dcn_v2_fwd(input.data_ptr(), offset.data_ptr(), mask.data_ptr(), weight.data_ptr(), bias.data_ptr() if bias is not None else None,
                                  out.data_ptr(), B, Cin, Cout, H, W, dg, _lib.current_stream_ptr(input.device))
    _lib.check(rc, "motif_dcn_v2_fwd")
    return out


def install():
    """Rebind ``dcn_v2_conv`` in every imported ``...DCNv2.dcn_v2`` module (``DCN.forward`` / ``DCN_sep.forward`` look the
    name up at call time, ``dcn_v2.py:94, 131``)."""
    n = 0
    for name, mod in list(sys.modules.items()):
        if mod is not None and name.endswith("DCNv2.dcn_v2") and hasattr(mod, "DCNv2"):
            mod.dcn_v2_conv = dcn_v2_conv
            n += 1
    return n
