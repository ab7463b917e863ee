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


def partition_rows(n_rows: int, world_size: int, align: int = 16) -> List[Tuple[int, int]]:
    """Contiguous destination row bands ``[begin, end)``, one per rank, begins aligned to ``align`` rows (the L2 band of the
    gather kernel's CTA order in band mode: two 8-row blocks), as balanced as the alignment allows (a band may be empty when there are more ranks than blocks)."""
    blocks = (n_rows + align - 1) // align
    base, extra = divmod(blocks, world_size)
    out, start = [], 0
    for r in range(world_size):
        size = base + (1 if r < extra else 0)
        b, e = min(start * align, n_rows), min((start + size) * align, n_rows)
        out.append((b, e))
        start += size
    return out


def slice_plan(sizes, world_size: int, rank: int):
    """Sliced copy-in of a clip every rank can read (pinned host memory of the node): the flat latent buffer
    ``[feat | flow_feat | residual]`` is padded to ``per * world_size`` elements and rank ``r`` fills ``[r * per, (r + 1) * per)``
    over its own PCIe link before an in-place all-gather.  Returns ``(per, [(tensor index, src begin, src end, dst begin)])``."""
    total = int(sum(sizes))
    per = (total + world_size - 1) // world_size
    lo, hi, off, parts = rank * per, (rank + 1) * per, 0, []
    for i, n in enumerate(sizes):
        a, b = max(lo, off), min(hi, off + int(n))
        if b > a:
            parts.append((i, a - off, b - off, a))
        off += int(n)
    return per, parts


def all_gather_slices(flat: torch.Tensor, per: int, group=None):
    """In-place all-gather of the ranks' slices of ``flat`` (``per * world_size`` elements; NCCL over NVLink, gloo in CPU tests)."""
    r = dist.get_rank(group)
    dist.all_gather_into_tensor(flat, flat[r * per:(r + 1) * per].clone() if flat.device.type == "cpu" else flat[r * per:(r + 1) * per], group=group)
    return flat


def check_halo(flow_y_max: torch.Tensor, halo: int, group=None) -> float:
    """Largest ``|flow_y|`` (HR pixels) any rank saw among the sources of its own band -- every source row belongs to exactly
    one band, so this is the maximum over the frame.  A band decode with this ``halo`` was exact iff the value is below
    ``halo - 1`` (a source farther away than that from a band cannot reach it).  One 1-element all-reduce."""
    m = flow_y_max.max().reshape(1)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(m, op=dist.ReduceOp.MAX, group=group)
    return float(m.item())


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


class LatentExchange:
    """The path's one exchange step, off the critical path: the source rank's latents are broadcast into one of ``depth``
    preallocated flat buffers on a side stream, so that the broadcast of clip k + 1 overlaps the decode of clip k.

        ex = LatentExchange(shapes, device, src=0)
        ex.start(lat_or_None)          # enqueue copy + broadcast of the next clip (source rank passes its three tensors)
        feat, flow_feat, residual = ex.take()   # current stream waits for the oldest started clip; returns views
        ...decode...
        ex.release()                   # the decode reading those views has been enqueued on the current stream
    """

    def __init__(self, shapes, device, src: int = 0, group=None, depth: int = 2, dtype=torch.float32):
        self.shapes = [tuple(s) for s in shapes]
        self.sizes = [int(torch.Size(s).numel()) for s in self.shapes]
        self.device, self.src, self.group, self.depth = torch.device(device), src, group, depth
        self.flat = [torch.empty(sum(self.sizes), dtype=dtype, device=self.device) for _ in range(depth)]
        self.stream = torch.cuda.Stream(self.device) if self.device.type == "cuda" else None
        self.ready = [None] * depth   # event: buffer filled
        self.free = [None] * depth    # event: the decode that read the buffer has been enqueued and finished
        self.started = 0
        self.taken = 0

    def _views(self, i):
        parts = torch.split(self.flat[i], self.sizes)
        return tuple(p.view(s) for p, s in zip(parts, self.shapes))

    def start(self, latents=None):
        if self.started - self.taken >= self.depth:
            raise RuntimeError("LatentExchange: all buffers are in flight (take / release first)")
        i = self.started % self.depth
        self.started += 1
        is_src = dist.get_rank(self.group) == self.src
        if self.stream is None:  # CPU (gloo tests): same semantics, no streams
            if is_src:
                torch.cat([t.reshape(-1) for t in latents], out=self.flat[i])
            dist.broadcast(self.flat[i], src=self.src, group=self.group)
            return
        cur = torch.cuda.current_stream(self.device)
        self.stream.wait_stream(cur)  # the source tensors were produced on the caller's stream
        with torch.cuda.stream(self.stream):
            if self.free[i] is not None:
                self.stream.wait_event(self.free[i])
            if is_src:
                for dst, t in zip(torch.split(self.flat[i], self.sizes), latents):
                    dst.copy_(t.reshape(-1), non_blocking=True)
            dist.broadcast(self.flat[i], src=self.src, group=self.group)
            ev = torch.cuda.Event()
            ev.record(self.stream)
            self.ready[i] = ev

    def take(self):
        if self.taken >= self.started:
            raise RuntimeError("LatentExchange: nothing started")
        i = self.taken % self.depth
        if self.stream is not None:
            torch.cuda.current_stream(self.device).wait_event(self.ready[i])
        return self._views(i)

    def release(self):
        i = self.taken % self.depth
        self.taken += 1
        if self.stream is not None:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(self.device))
            self.free[i] = ev


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
