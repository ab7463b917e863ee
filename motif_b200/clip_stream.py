"""Host-buffer front end of the decoder: a double-buffered clip pipeline.

``SpaceTimeDecoder.decode`` works on resident device tensors.  A caller that holds the LR latents in
host memory (the encoder ran elsewhere, or clips are streamed from disk) pays a host->device copy
of 320*H*W*4 bytes and a device->host copy of 12 bytes per output pixel-timestamp per clip -- at
Adobe240 size 1.4 ms each way over PCIe against ~6 ms of decode.  ``ClipStream`` overlaps them
with the decode of the neighbouring clips: three CUDA streams (copy-in, compute, copy-out), two
sets of device buffers, events between them.  Every clip's copies are issued by ``submit`` itself
(nothing is cached across clips); ``synchronize`` waits for everything submitted so far.

With ``torch.distributed`` initialised and ``world_size > 1`` the source rank copies the latents
in and broadcasts them (the path's one exchange) ON THE COPY-IN STREAM, i.e. behind the decode of the
previous clip; every rank then decodes its share -- a range of timestamps (``n_range``), a band of
destination rows with a source halo (``row_range`` / ``halo``), or both -- and copies only
that share out.

``sliced_copy_in=True`` (every rank holds the clip in pinned host memory, e.g. one shared-memory segment of the node): no rank
pulls the whole clip through its own PCIe link -- at 8 GPUs that copy (1.3 ms for the 73.7 MB of an Adobe clip) is longer than the
1.2 ms band decode -- but every rank copies 1 / world_size of the flat latent buffer over ITS OWN link and the parts are
all-gathered over NVLink (in place, on the copy-in stream, behind the previous clip's decode).

``band_copy_in=True`` (same assumption, destination row bands): no collective at all -- a band decode reads only the LR rows of its
band and halo (``SpaceTimeDecoder.lr_rows_of_band``: 18 % of the clip at 8 ranks), so every rank pulls exactly those rows of the three
NCHW tensors out of the pinned host memory with one strided copy each (``motif_memcpy2d_async``) over its own PCIe link.
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import torch

from .decoder import SpaceTimeDecoder
from .sharding import all_gather_slices, slice_plan


class ClipStream:
    def __init__(self, decoder: SpaceTimeDecoder, depth: int = 2, distributed: bool = False, src: int = 0, group=None, return_flow: bool = False,
                 sliced_copy_in: bool = False, band_copy_in: bool = False):
        if depth < 1:
            raise ValueError("depth must be >= 1")
        self.dec = decoder
        self.dev = decoder.device
        self.depth = depth
        self.distributed = distributed
        self.src = src
        self.group = group
        self.sliced = bool(sliced_copy_in and distributed)
        self.band_rows_only = bool(band_copy_in)
        self.return_flow = return_flow  # also produce the forward's second output (flow / 20 / (HH/H), Ours.py:858) on the device
        self.s_in = torch.cuda.Stream(self.dev)
        self.s_out = torch.cuda.Stream(self.dev)
        self._slots = [None] * depth
        self._k = 0
        self.flow_y_max = torch.zeros(64, dtype=torch.float32, device=self.dev)  # of the last band decode (sharding.check_halo)

    def _slot(self, k, shapes, out_shape):
        i = k % self.depth
        sl = self._slots[i]
        if sl is None or sl["shapes"] != shapes or sl["out"].shape != out_shape:
            if sl is not None:
                # the old buffers may still be the target / source of copies in flight on the side streams: tell the
                # caching allocator (they were allocated on the compute stream) before dropping them
                sl["flat"].record_stream(self.s_in)
                sl["out"].record_stream(self.s_out)
            sizes = [int(torch.Size(s).numel()) for s in shapes]
            total, world = sum(sizes), self._world()
            per = (total + world - 1) // world  # elements per rank of a sliced copy-in (the buffer is padded to per * world)
            flat = torch.empty(per * world, dtype=torch.float32, device=self.dev)
            sl = {"shapes": shapes, "flat": flat, "per": per, "lat": [p.view(s) for p, s in zip(torch.split(flat[:total], sizes), shapes)],
                  "out": torch.empty(out_shape, dtype=torch.float32, device=self.dev),
                  "ev_in": torch.cuda.Event(), "ev_free": None, "ev_done": torch.cuda.Event(), "ev_out": None}
            self._slots[i] = sl
        return sl

    def _world(self):
        import torch.distributed as dist

        return dist.get_world_size(self.group) if self.distributed else 1

    def submit(self, feat_h: Optional[torch.Tensor], flow_feat_h: Optional[torch.Tensor], residual_h: Optional[torch.Tensor],
               target_t, hr_size: Tuple[int, int], out_h: Optional[torch.Tensor], n_range: Optional[Tuple[int, int]] = None,
               shapes: Optional[Sequence[Tuple[int, ...]]] = None, row_range: Optional[Tuple[int, int]] = None, halo: int = 0):
        """Queue one clip.  ``feat_h`` / ``flow_feat_h`` / ``residual_h``: pinned host tensors (on ranks other
        than ``src`` of a distributed stream pass ``None`` and the three ``shapes``).  ``out_h``: pinned host tensor
        ``[n_end - n_begin, B, 3, r1 - r0, WW]`` receiving this rank's share of the frames (or ``None`` to leave them on the
        device).  Returns the device frame buffer ``[N, B, 3, HH, WW]`` of this slot (valid until the slot is reused)."""
        import torch.distributed as dist

        band_pull = self.band_rows_only and row_range is not None
        is_src = (not self.distributed) or self.sliced or band_pull or dist.get_rank(self.group) == self.src
        if is_src:
            shapes = tuple(tuple(t.shape) for t in (feat_h, flow_feat_h, residual_h))
        elif shapes is None:
            raise ValueError("non-source ranks must pass the latent shapes")
        shapes = tuple(tuple(s) for s in shapes)
        tt = torch.as_tensor(target_t, dtype=torch.float32).reshape(shapes[2][0], -1)
        B, N = tt.shape
        HH, WW = int(hr_size[0]), int(hr_size[1])
        n0, n1 = (0, N) if n_range is None else n_range
        r0, r1 = (0, HH) if row_range is None else row_range
        sl = self._slot(self._k, shapes, (N, B, 3, HH, WW))
        self._k += 1
        compute = torch.cuda.current_stream(self.dev)
        with torch.cuda.stream(self.s_in):
            if sl["ev_free"] is not None:
                self.s_in.wait_event(sl["ev_free"])            # the decode that last read these buffers has finished
            if band_pull:
                # only the LR rows this rank's band (+ halo) can select, straight from the pinned NCHW tensors: no collective
                from . import _lib

                lib = _lib.load()
                H, W = shapes[2][2], shapes[2][3]
                lr0, lr1 = self.dec.lr_rows_of_band(H, HH, (r0, r1), halo)
                if lr1 > lr0:
                    for dst, src_t in zip(sl["lat"], (feat_h, flow_feat_h, residual_h)):
                        planes, pitch = src_t.shape[0] * src_t.shape[1], H * W * 4
                        rc = lib.motif_memcpy2d_async(dst.data_ptr() + lr0 * W * 4, pitch, src_t.data_ptr() + lr0 * W * 4, pitch, (lr1 - lr0) * W * 4, planes, 1,
                                                      self.s_in.cuda_stream)
                        _lib.check(rc, "motif_memcpy2d_async")
            elif self.sliced:
                # this rank's 1 / world of the flat buffer over its own PCIe link, then an in-place all-gather over NVLink
                hosts = (feat_h, flow_feat_h, residual_h)
                per, parts = slice_plan([t.numel() for t in hosts], self._world(), dist.get_rank(self.group))
                for i, a, b, dst in parts:
                    sl["flat"][dst:dst + (b - a)].copy_(hosts[i].reshape(-1)[a:b], non_blocking=True)
                all_gather_slices(sl["flat"], per, self.group)
            else:
                if is_src:
                    for dst, src_t in zip(sl["lat"], (feat_h, flow_feat_h, residual_h)):
                        dst.copy_(src_t, non_blocking=True)
                if self.distributed:
                    dist.broadcast(sl["flat"], src=self.src, group=self.group)   # one flat buffer, no staging copy
            sl["ev_in"].record(self.s_in)
        compute.wait_event(sl["ev_in"])
        if sl["ev_out"] is not None:
            compute.wait_event(sl["ev_out"])                   # the copy-out that last read this frame buffer has finished
        lat = sl["lat"]
        band = {} if row_range is None else {"row_range": (r0, r1), "halo": halo, "flow_y_max": self.flow_y_max}
        if n1 > n0 and r1 > r0:
            self.dec.decode(lat[0], lat[1], lat[2], tt, (HH, WW), n_range=(n0, n1), return_flow=self.return_flow, out=sl["out"], **band)
        sl["ev_done"].record(compute)
        sl["ev_free"] = sl["ev_done"]
        if out_h is not None and n1 > n0 and r1 > r0:
            with torch.cuda.stream(self.s_out):
                self.s_out.wait_event(sl["ev_done"])
                out_h[: n1 - n0, :, :, : r1 - r0].copy_(sl["out"][n0:n1, :, :, r0:r1], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(self.s_out)
                sl["ev_out"] = ev
        return sl["out"]

    def synchronize(self):
        torch.cuda.current_stream(self.dev).synchronize()
        self.s_in.synchronize()
        self.s_out.synchronize()
