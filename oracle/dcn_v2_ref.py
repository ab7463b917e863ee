"""CPU restatement of the reference's modulated deformable convolution forward (TEST INFRASTRUCTURE ONLY).

``models/modules/DCNv2``: ``dcn_v2_conv(input, offset, mask, weight, bias, stride, padding, dilation, deformable_groups)``
(``dcn_v2.py:13-47``) -> ``dcn_v2_forward`` -> ``modulated_deformable_im2col_gpu_kernel`` + SGEMM
(``src/cuda/dcn_v2_im2col_cuda.cu:25-55`` bilinear, ``:125-195`` im2col; ``src/cuda/dcn_v2_cuda.cu`` the product with the
``[C_out, C_in * kh * kw]`` weight matrix plus bias).  The reference's extension needs ``THC/THC.h`` and cannot be built on
torch 2.x, and the reference has no test for it, so parity is pinned against ``torchvision.ops.deform_conv2d`` (torchvision
0.26, a third-party implementation of the same algorithm, the stand-in the oracle already uses for the encoder, SURVEY 8c
shim 3): ``tests/test_oracle_pins.py``.
"""
from __future__ import annotations

import torch


def _bilinear(plane: torch.Tensor, h: torch.Tensor, w: torch.Tensor) -> torch.Tensor:
    """``dmcn_im2col_bilinear`` (``dcn_v2_im2col_cuda.cu:25-55``): plane ``[C, H, W]``, h / w ``[Ho, Wo]`` -> ``[C, Ho, Wo]``."""
    C, H, W = plane.shape
    h_low, w_low = torch.floor(h), torch.floor(w)
    lh, lw = h - h_low, w - w_low
    hh, hw = 1 - lh, 1 - lw
    h_low, w_low = h_low.long(), w_low.long()
    h_high, w_high = h_low + 1, w_low + 1

    def tap(hi, wi, ok):
        v = plane[:, hi.clamp(0, H - 1), wi.clamp(0, W - 1)]
        return torch.where(ok.unsqueeze(0), v, torch.zeros_like(v))

    v1 = tap(h_low, w_low, (h_low >= 0) & (w_low >= 0))
    v2 = tap(h_low, w_high, (h_low >= 0) & (w_high <= W - 1))
    v3 = tap(h_high, w_low, (h_high <= H - 1) & (w_low >= 0))
    v4 = tap(h_high, w_high, (h_high <= H - 1) & (w_high <= W - 1))
    return (hh * hw) * v1 + (hh * lw) * v2 + (lh * hw) * v3 + (lh * lw) * v4


def dcn_v2_conv(inp, offset, mask, weight, bias, stride=1, padding=1, dilation=1, deformable_groups=1):
    """input ``[B,Cin,H,W]``, offset ``[B, dg*2*kh*kw, Ho, Wo]`` (per group and tap: dh, dw), mask ``[B, dg*kh*kw, Ho, Wo]``,
    weight ``[Cout,Cin,kh,kw]`` -> ``[B,Cout,Ho,Wo]``."""
    B, Cin, H, W = inp.shape
    Cout, _, kh, kw = weight.shape
    Ho = (H + 2 * padding - (dilation * (kh - 1) + 1)) // stride + 1
    Wo = (W + 2 * padding - (dilation * (kw - 1) + 1)) // stride + 1
    cpg = Cin // deformable_groups
    hs = (torch.arange(Ho) * stride - padding).view(Ho, 1).to(inp.dtype)
    ws = (torch.arange(Wo) * stride - padding).view(1, Wo).to(inp.dtype)
    cols = torch.zeros(B, Cin, kh * kw, Ho, Wo, dtype=inp.dtype)
    for b in range(B):
        for g in range(deformable_groups):
            plane = inp[b, g * cpg:(g + 1) * cpg]
            for i in range(kh):
                for j in range(kw):
                    t = i * kw + j
                    off_h = offset[b, g * 2 * kh * kw + 2 * t]
                    off_w = offset[b, g * 2 * kh * kw + 2 * t + 1]
                    m = mask[b, g * kh * kw + t]
                    h_im = hs + i * dilation + off_h
                    w_im = ws + j * dilation + off_w
                    inside = (h_im > -1) & (w_im > -1) & (h_im < H) & (w_im < W)  # dcn_v2_im2col_cuda.cu:177
                    val = _bilinear(plane, h_im, w_im)
                    cols[b, g * cpg:(g + 1) * cpg, t] = torch.where(inside.unsqueeze(0), val, torch.zeros_like(val)) * m
    out = torch.einsum("ok,bkp->bop", weight.reshape(Cout, Cin * kh * kw), cols.reshape(B, Cin * kh * kw, Ho * Wo))
    return out.reshape(B, Cout, Ho, Wo) + bias.view(1, Cout, 1, 1)
