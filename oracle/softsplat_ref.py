"""CPU restatement of the three forward-splatting operators (TEST INFRASTRUCTURE ONLY).

Follows, line by line, the reference kernels and wrappers:

* sum splat   ``models/softsplat_cp.py:12-52`` (kernel), ``:221-259`` (launcher),
  ``:320-347`` (``FunctionSoftsplat`` mode wrapper)
* max splat   ``models/softsplat_max_cp.py:12-58`` (kernel + ``atomicMaxFloat``),
  ``:254`` (output initialised to ONES), ``:339-343``
* count splat ``models/softsplat_count_cp.py:14-52`` (kernel adds the *unweighted*
  input), ``:163-165`` (wrapper replaces the input by ones)

All arithmetic is fp32 exactly as in the kernels: ``fx = float(x) + flow``,
``x0 = (int)floor(fx)``, the four corner weights are products of two fp32
differences, a corner contributes iff it lies inside the image.  The only
freedom the reference leaves is the order of the float ``atomicAdd``s; this
restatement adds in source raster order, corner order NW, NE, SW, SE.

Pinned by ``tests/test_oracle_pins.py`` against the reference kernel strings
compiled for the host (``oracle/build_ref.py``) and the committed golden vectors.
"""
from __future__ import annotations

import torch

__all__ = [
    "splat_footprint",
    "splat_sum",
    "splat_max",
    "splat_count",
    "function_softsplat",
    "function_softsplat_max",
    "function_softsplat_count",
]


def splat_footprint(flow: torch.Tensor):
    """Corner indices, in-bounds masks and fp32 weights of the bilinear footprint.

    ``softsplat_cp.py:23-38`` (identical in the max and count files).
    Returns four tuples ``(dest_linear_index[N,H,W] int64, inside[N,H,W] bool,
    weight[N,H,W] fp32)`` in the order NW, NE, SW, SE.
    """
    assert flow.dtype == torch.float32 and flow.dim() == 4 and flow.shape[1] == 2
    n, _, h, w = flow.shape
    xs = torch.arange(w, dtype=torch.float32).view(1, 1, w)
    ys = torch.arange(h, dtype=torch.float32).view(1, h, 1)
    fx = xs + flow[:, 0]  # fltOutputX
    fy = ys + flow[:, 1]  # fltOutputY
    x0 = torch.floor(fx).to(torch.int64)
    y0 = torch.floor(fy).to(torch.int64)
    x1 = x0 + 1
    y1 = y0 + 1
    x0f, x1f = x0.to(torch.float32), x1.to(torch.float32)
    y0f, y1f = y0.to(torch.float32), y1.to(torch.float32)
    w_nw = (x1f - fx) * (y1f - fy)
    w_ne = (fx - x0f) * (y1f - fy)
    w_sw = (x1f - fx) * (fy - y0f)
    w_se = (fx - x0f) * (fy - y0f)

    def corner(cx, cy, wt):
        inside = (cx >= 0) & (cx < w) & (cy >= 0) & (cy < h)
        lin = cy.clamp(0, h - 1) * w + cx.clamp(0, w - 1)
        return lin, inside, wt

    return (
        corner(x0, y0, w_nw),
        corner(x1, y0, w_ne),
        corner(x0, y1, w_sw),
        corner(x1, y1, w_se),
    )


def _scatter(inp: torch.Tensor, flow: torch.Tensor, mode: str) -> torch.Tensor:
    n, c, h, w = inp.shape
    assert flow.shape == (n, 2, h, w)
    inp = inp.contiguous().float()
    flow = flow.contiguous().float()
    if mode == "max":
        out = inp.new_ones(n, c, h * w)  # softsplat_max_cp.py:254
    else:
        out = inp.new_zeros(n, c, h * w)
    src = inp.view(n, c, h * w)
    corners = splat_footprint(flow)
    # interleave the corners so that contributions are added in the order a single host
    # thread executes the reference kernel: source raster order, NW, NE, SW, SE per source
    lin = torch.stack([cn[0].view(n, h * w) for cn in corners], 2).view(n, h * w * 4)
    inside = torch.stack([cn[1].view(n, h * w) for cn in corners], 2).view(n, h * w * 4)
    wt = torch.stack([cn[2].view(n, h * w) for cn in corners], 2).view(n, 1, h * w * 4)
    srcpix = torch.arange(h * w).repeat_interleave(4)
    for b in range(n):
        sel = inside[b].nonzero(as_tuple=True)[0]
        if sel.numel() == 0:
            continue
        idx = lin[b, sel]
        vals = src[b][:, srcpix[sel]]
        if mode == "sum":
            contrib = vals * wt[b][:, sel]
        elif mode == "count":
            # kernel adds the raw input value, no weight (softsplat_count_cp.py:39-50)
            contrib = vals
        else:
            contrib = vals * wt[b][:, sel]
            # atomicMaxFloat (softsplat_max_cp.py:13-18): a negative or NaN candidate can
            # never replace a stored value >= 1.0 (uint atomicMin against a positive
            # pattern keeps the positive pattern); -0.0 has int pattern INT_MIN.
            contrib = torch.where(contrib >= 0, contrib, torch.full_like(contrib, -1.0))
            out[b].scatter_reduce_(1, idx.view(1, -1).expand(c, -1), contrib, reduce="amax", include_self=True)
            continue
        # sequential accumulation per channel row, in index order
        for ch in range(c):
            out[b, ch].index_add_(0, idx, contrib[ch].contiguous())
    return out.view(n, c, h, w)


def splat_sum(inp, flow):
    """``_FunctionSoftsplat.forward`` of ``softsplat_cp.py:221-259``."""
    return _scatter(inp, flow, "sum")


def splat_max(inp, flow):
    """``_FunctionSoftsplat.forward`` of ``softsplat_max_cp.py:240-278``."""
    return _scatter(inp, flow, "max")


def splat_count(inp, flow):
    """``_FunctionSoftsplat.forward`` of ``softsplat_count_cp.py:117-155`` (raw input added)."""
    return _scatter(inp, flow, "count")


def function_softsplat(tenInput, tenFlow, tenMetric, strType):
    """``FunctionSoftsplat`` of ``softsplat_cp.py:320-347``.

    Returns ``(tenOutput[:, :-1], tenOutput[:, -1:])`` UN-normalised (the
    division is commented out in the reference, ``:340-344``).  ``'summation'``
    raises ``UnboundLocalError`` in the reference (``tenNormalize`` never bound);
    the restatement mirrors the documented product behaviour and returns
    ``(full output, None)`` for it.
    """
    assert tenMetric is None or tenMetric.shape[1] == 1
    assert strType in ["summation", "average", "linear", "softmax"]
    if strType == "average":
        tenInput = torch.cat([tenInput, tenInput.new_ones(tenInput.shape[0], 1, tenInput.shape[2], tenInput.shape[3])], 1)
    elif strType == "linear":
        tenInput = torch.cat([tenInput * tenMetric, tenMetric], 1)
    elif strType == "softmax":
        tenInput = torch.cat([tenInput * tenMetric.exp(), tenMetric.exp()], 1)
    tenOutput = splat_sum(tenInput, tenFlow)
    if strType == "summation":
        return tenOutput, None
    return tenOutput[:, :-1, :, :], tenOutput[:, -1:, :, :]


def function_softsplat_max(tenInput, tenFlow):
    """``FunctionSoftsplat`` of ``softsplat_max_cp.py:339-343``."""
    return splat_max(tenInput, tenFlow)


def function_softsplat_count(tenInput, tenFlow):
    """``FunctionSoftsplat`` of ``softsplat_count_cp.py:163-165`` (input replaced by ones)."""
    ones = tenInput.new_ones(tenInput.shape[0], 1, tenInput.shape[2], tenInput.shape[3])
    return splat_count(ones, tenFlow)
