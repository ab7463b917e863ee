"""CPU restatement of the PWC-Net cost-volume operator (TEST INFRASTRUCTURE ONLY).

Follows ``OpticalFlow/correlation.py``:

* ``kernel_Correlation_rearrange`` ``:17-42``   NCHW -> zero-padded (pad 4) NHWC copy
* ``kernel_Correlation_updateOutput`` ``:44-112``  81 displacements (+-4), kernel 1,
  stride 1: ``out[b, tc, y, x] = (sum_c f1[b,c,y,x] * f2[b,c,y+tc/9-4,x+tc%9-4]) / C``
  with zeros outside the image
* ``_FunctionCorrelation.forward`` ``:294-348``, ``FunctionCorrelation`` ``:415-416``

``function_correlation`` is the fast shifted-product form used as the oracle at
realistic sizes.  ``function_correlation_lane_order`` reproduces the reference's
summation order (lane l accumulates c = l, l+32, ... then lane 0 adds the 32
partials in order, divide by C last) with separate multiply and add roundings;
the device build of the reference contracts the multiply-add into an FMA, so
even this order is only equal up to fp32 rounding -- tests compare with a
tolerance.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

__all__ = ["function_correlation", "function_correlation_lane_order"]


def function_correlation(first: torch.Tensor, second: torch.Tensor) -> torch.Tensor:
    assert first.shape == second.shape and first.dim() == 4
    b, c, h, w = first.shape
    first = first.float()
    pad2 = F.pad(second.float(), (4, 4, 4, 4))  # rearrange kernel: zero border of 4
    out = first.new_zeros(b, 81, h, w)
    for tc in range(81):
        dx = tc % 9  # s2o + 4
        dy = tc // 9  # s2p + 4
        out[:, tc] = (first * pad2[:, :, dy:dy + h, dx:dx + w]).sum(1) / float(c)
    return out


def function_correlation_lane_order(first: torch.Tensor, second: torch.Tensor) -> torch.Tensor:
    """Same values, reference summation order (``correlation.py:75-109``)."""
    b, c, h, w = first.shape
    first = first.float()
    pad2 = F.pad(second.float(), (4, 4, 4, 4))
    out = first.new_zeros(b, 81, h, w)
    for tc in range(81):
        dx, dy = tc % 9, tc // 9
        prod = first * pad2[:, :, dy:dy + h, dx:dx + w]  # [b,c,h,w]
        lanes = []
        for lane in range(32):
            acc = first.new_zeros(b, h, w)
            for ch in range(lane, c, 32):
                acc = acc + prod[:, ch]
            lanes.append(acc)
        total = first.new_zeros(b, h, w)
        for lane in range(32):
            total = total + lanes[lane]
        out[:, tc] = total / float(c)
    return out
