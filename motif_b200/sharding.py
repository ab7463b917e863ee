"""Multi-GPU plumbing of the decoder: one process per GPU, ``torch.distributed``.

The path has exactly one exchange step: the clip-invariant LR latents
``[feat(2B*64) | flow_feat(2B*64) | residual(B*64)] x H x W`` are broadcast once per clip from the
rank that ran the encoder, after which every rank decodes its own timestamps with no further
communication (``Ours.py:783-856`` is independent per timestamp).  The reference itself has no
multi-GPU inference path (``DataParallel`` over ``gpu_ids: [0]``, ``VideoSR_base_model.py:36``).
"""
from __future__ import annotations

from typing import List, Tuple

import torch
import torch.distributed as dist


def partition_timestamps(n_timestamps: int, world_size: int) -> List[Tuple[int, int]]:
    """Contiguous, balanced ``[begin, end)`` timestamp ranges, one per rank (may be empty)."""
    base, extra = divmod(n_timestamps, world_size)
    out, start = [], 0
    for r in range(world_size):
        size = base + (1 if r < extra else 0)
        out.append((start, start + size))
        start += size
    return out


def broadcast_latents(feat: torch.Tensor, flow_feat: torch.Tensor, residual: torch.Tensor, src: int = 0, group=None):
    """One flat broadcast of the three latent tensors (NCCL over NVLink on GPUs, gloo in CPU tests).
    Non-source ranks pass correctly shaped buffers; returns the three tensors (views of one buffer)."""
    sizes = [feat.numel(), flow_feat.numel(), residual.numel()]
    flat = torch.empty(sum(sizes), dtype=feat.dtype, device=feat.device)
    if dist.get_rank(group) == src:
        torch.cat([feat.reshape(-1), flow_feat.reshape(-1), residual.reshape(-1)], out=flat)
    dist.broadcast(flat, src=src, group=group)
    a, b, c = torch.split(flat, sizes)
    return a.view_as(feat), b.view_as(flow_feat), c.view_as(residual)


def gather_frames(local: torch.Tensor, ranges: List[Tuple[int, int]], group=None) -> torch.Tensor:
    """All-gather the per-rank ``[n_local, B, 3, HH, WW]`` frames into ``[N, B, 3, HH, WW]`` (optional:
    12 bytes per output pixel-timestamp; the decode itself needs no collective)."""
    world = dist.get_world_size(group)
    n_max = max(e - b for b, e in ranges)
    pad = torch.zeros((n_max,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    return torch.cat([bufs[r][: e - b] for r, (b, e) in enumerate(ranges)], 0)
