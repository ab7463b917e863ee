"""HR-resolution RAFT schedule of the surround (second half).

``LunaTokis.forward`` runs its pretrained RAFT on FOUR HR frame pairs -- 00, 01, 10, 11 -- and then multiplies the flows of
the pairs 00 and 11 by zero (``Ours.py:544-555``).  ``RAFT.forward`` (``models/core/raft.py:86-144``) in turn pushes both
images of every pair through its feature encoder, so the 4B-pair call encodes 8B HR images of which only 2B are distinct, runs
the context encoder on 4B (2B distinct) and iterates the update block on 4B pairs of which 2B are discarded.

``flow_two_pairs`` runs the same sub-modules of the instance's own ``flow_predictor`` -- nothing of RAFT is re-implemented --
on the two pairs that survive: the feature encoder once per distinct frame (2B images instead of 8B), the context encoder on 2B
images, the update iterations on the pairs 01 and 10.  ``four_pair_flows`` returns the reference's ``[4B, 2, HH, WW]`` layout
with exact zeros for 00 / 11 (the reference's ``flow *= 0.`` leaves zeros whose sign bit follows the discarded flow; every later
use -- back-warp coordinates, L1 means, the gaussian variance, the ``flow_process`` convolution -- gives equal values for +0 and -0).

``LookupBlock`` is ``AlternateCorrBlock`` (``models/core/corr.py:59-87``) with the layout changes hoisted out of the iteration
loop: the reference re-permutes ``fmap1`` and the four pooled ``fmap2`` levels to channels-last on EVERY call (every RAFT
iteration); here that happens once per pair and each iteration is four ``motif_raft_corr_lookup`` launches.

Both are glue around reference sub-modules (like ``luna_tokis.surround``); ``surround`` uses them when the model's
``flow_predictor`` has RAFT's module tree and falls back to the reference's four-pair call otherwise.
"""
from __future__ import annotations

import sys

import torch
import torch.nn.functional as F

from . import alt_cuda_corr


class LookupBlock:
    """``AlternateCorrBlock(fmap1, fmap2, radius=r)`` (``corr.py:59-87``): ``__call__(coords [B,2,H,W]) -> [B, 4*(2r+1)^2, H, W]``."""

    def __init__(self, fmap1, fmap2, num_levels: int = 4, radius: int = 4):
        self.num_levels, self.radius, self.dim = num_levels, radius, fmap1.shape[1]
        self.f1 = fmap1.permute(0, 2, 3, 1).contiguous()
        self.f2 = []
        for i in range(num_levels):  # pyramid[i][1], i = 0 .. num_levels - 1 (corr.py:64-68, 78; the reference pools once more than it uses)
            if i > 0:
                fmap2 = F.avg_pool2d(fmap2, 2, stride=2)
            self.f2.append(fmap2.permute(0, 2, 3, 1).contiguous())

    def __call__(self, coords):
        coords = coords.permute(0, 2, 3, 1)
        B, H, W, _ = coords.shape
        if self.f1.is_cuda and self.dim in (128, 256) and self.radius <= 3 and self.num_levels <= 4:
            # one launch for all levels: coords / 2^l, the stacking and the division by sqrt(dim) happen inside the kernel
            return alt_cuda_corr.forward_pyramid(self.f1, self.f2, coords.float(), self.radius, normalize=True)
        out = []
        for i in range(self.num_levels):
            coords_i = (coords / 2 ** i).reshape(B, 1, H, W, 2).contiguous()
            corr, = alt_cuda_corr.forward(self.f1, self.f2[i], coords_i, self.radius)
            out.append(corr.squeeze(1))
        corr = torch.stack(out, dim=1).reshape(B, -1, H, W)
        return corr / torch.sqrt(torch.tensor(self.dim).float())


def is_raft(flow_predictor) -> bool:
    """Does the module have the tree ``RAFT.forward`` uses (``raft.py:24-58``)?"""
    return all(hasattr(flow_predictor, a) for a in ("fnet", "cnet", "update_block", "initialize_flow", "upsample_flow", "args", "hidden_dim", "context_dim"))


def flow_two_pairs(raft, fr0, fr1, iters: int = 12, lookup: str = "auto"):
    """Final ``flow_up`` of ``raft(cat[fr0, fr1], cat[fr1, fr0], iters)`` (``raft.py:86-144``, ``upsample=True``, no
    ``flow_init``), ``[2B, 2, HH, WW]``: the pairs 01 then 10.  ``fr0`` / ``fr1`` are the images the reference hands to RAFT
    (already scaled to 0..255, ``Ours.py:544``).  ``lookup``: ``'motif'`` = ``LookupBlock``, ``'reference'`` = the class the
    reference's own forward would build (``args.alternate_corr``), ``'auto'`` = ``'motif'`` on CUDA when the model asks for the
    alternate block."""
    mod = sys.modules[type(raft).__module__]  # models.core.raft: autocast, the two correlation blocks, upflow8
    autocast = getattr(mod, "autocast")
    B = fr0.shape[0]
    image = torch.cat([fr0, fr1], dim=0)
    image = (2 * (image / 255.0) - 1.0).contiguous()
    hdim, cdim = raft.hidden_dim, raft.context_dim
    mixed = bool(getattr(raft.args, "mixed_precision", False))
    with autocast(enabled=mixed):
        f = raft.fnet(image)  # once per distinct frame; instance norm is per sample, so the batch composition does not matter
    f = f.float()
    fmap1, fmap2 = f, torch.cat([f[B:], f[:B]], dim=0)
    alternate = bool(getattr(raft.args, "alternate_corr", False))
    radius = raft.args.corr_radius
    if lookup == "motif" or (lookup == "auto" and alternate and f.is_cuda):
        corr_fn = LookupBlock(fmap1, fmap2, radius=radius)
    elif alternate:
        corr_fn = mod.AlternateCorrBlock(fmap1, fmap2, radius=radius)
    else:
        corr_fn = mod.CorrBlock(fmap1, fmap2, radius=radius)
    with autocast(enabled=mixed):
        cnet = raft.cnet(image)
        net, inp = torch.split(cnet, [hdim, cdim], dim=1)
        net = torch.tanh(net)
        inp = torch.relu(inp)
    coords0, coords1 = raft.initialize_flow(image)
    flow_up = None
    for it in range(iters):
        coords1 = coords1.detach()
        corr = corr_fn(coords1)
        flow = coords1 - coords0
        with autocast(enabled=mixed):
            net, up_mask, delta_flow = raft.update_block(net, inp, corr, flow)
        coords1 = coords1 + delta_flow
        if it < iters - 1:
            continue  # the reference upsamples every iteration's flow and keeps only the last (Ours.py:545: [-1])
        flow_up = mod.upflow8(coords1 - coords0) if up_mask is None else raft.upsample_flow(coords1 - coords0, up_mask)
    return flow_up


def four_pair_flows(flow_predictor, fr0, fr1, iters: int = 12):
    """What ``Ours.py:544-545`` computes, ``[4B, 2, HH, WW]`` in the pair order 00, 01, 10, 11 -- with the two pairs whose flow the
    caller discards (``:552-553``) left at zero instead of being estimated.  ``fr0`` / ``fr1``: HR frames in 0..1."""
    if not is_raft(flow_predictor) or iters < 1:
        return flow_predictor(torch.cat([fr0, fr0, fr1, fr1], dim=0) * 255.0, torch.cat([fr0, fr1, fr0, fr1], dim=0) * 255.0, iters=iters)[-1]
    live = flow_two_pairs(flow_predictor, fr0 * 255.0, fr1 * 255.0, iters)
    B = fr0.shape[0]
    zeros = torch.zeros_like(live[:B])
    return torch.cat([zeros, live, zeros], dim=0)


def work_ratio(B: int = 1):
    """Images / pairs the reference's call processes per image / pair of this schedule (for the documentation and the bench)."""
    return {"fnet_images": (8 * B, 2 * B), "cnet_images": (4 * B, 2 * B), "update_pairs": (4 * B, 2 * B), "upsample_calls_per_pair": ("iters", 1)}

