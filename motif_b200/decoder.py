"""Host mirror of the space-time local implicit decoder (reference ``models/modules/Ours.py:659-858``).

``SpaceTimeDecoder`` owns the three SIREN MLPs' weights in the checkpoint (``state_dict``) layout
of ``LunaTokis`` -- ``imnet.net.{0,1,2}.linear.{weight,bias}``, ``imnet.net.3.{weight,bias}``, the
same for ``flow_imnet``, ``synth_net.net.{0..3}.linear.*``, ``synth_net.net.4.*`` and ``alpha``
 -- and decodes HR frames from the resident LR latents with ONE call into
``libmotif_b200.so`` (``motif_decode``): nearest-latent gather, ``imnet`` once per clip, then per
timestamp ``flow_imnet`` -> three forward splats of both reference frames -> blend + reliability
features -> ``synth_net`` -> clamp.

The coordinate sequences are built on the host exactly as ``make_coord`` (``Ours.py:874-889``)
builds them (fp32, two roundings per element) and only these 1-D tables are uploaded; the
reference uploads the full ``[HH*WW, 2]`` meshgrid on every call (``Ours.py:667-668``).
"""
from __future__ import annotations

import ctypes
from typing import Dict, Optional, Sequence, Tuple

import torch

from . import _lib

HOT_PREFIXES = ("imnet.", "flow_imnet.", "synth_net.")
_SIREN_LAYERS = {"imnet": 4, "flow_imnet": 4, "synth_net": 5}
_SIREN_SHAPES = {
    "flow_imnet": [(64, 67), (64, 64), (256, 64), (3, 256)],
    "imnet": [(64, 66), (64, 64), (256, 64), (64, 256)],
    "synth_net": [(64, 198), (64, 64), (64, 64), (256, 64), (3, 256)],
}
PRECISIONS = {"tf32x3": 0, "fp32": 1, "f16x3": 2}
DEFAULT_PRECISION = "f16x3"


def coord_sequence(n: int) -> torch.Tensor:
    """1-D pixel-centre sequence of ``make_coord`` for an axis of length ``n`` (``Ours.py:874-889``):
    python-double ``v0 + r`` and ``2 * r`` cast to fp32, one fp32 multiply, one fp32 add."""
    r = (1 - (-1)) / (2 * n)
    return -1 + r + (2 * r) * torch.arange(n).float()


def hr_size_from_scale(H: int, W: int, scale) -> Tuple[int, int]:
    """``Ours.py:525-529``: list form ``[[HH], [WW]]`` or ``round(H * scale)``."""
    if isinstance(scale, list):
        return int(scale[0][0]), int(scale[1][0])
    return round(H * scale), round(W * scale)


def _key(name: str, layer: int, last: bool, what: str) -> str:
    return f"{name}.net.{layer}.{what}" if last else f"{name}.net.{layer}.linear.{what}"


class SpaceTimeDecoder:
    def __init__(self, params: Dict[str, torch.Tensor], device="cuda", precision: str = DEFAULT_PRECISION, local_ensemble: bool = False):
        """``local_ensemble``: the reference's ``LunaTokis.local_ensemble`` flag (``Ours.py:453``; ``False`` as shipped).
        ``True`` evaluates every query at four shifted latents and blends them by the diagonally swapped area weights
        (``Ours.py:660-663, 754-764``); implemented by ``precision='f16x3'`` (four tensor-core passes) and ``'fp32'``."""
        _lib.load()  # fail loudly when the CUDA library is missing
        if precision not in PRECISIONS:
            raise ValueError(f"precision must be one of {list(PRECISIONS)}")
        if local_ensemble and precision == "tf32x3":
            raise NotImplementedError("local_ensemble=True is implemented by precision='f16x3' and 'fp32' (the shipped checkpoint runs with it off, Ours.py:453)")
        self.local_ensemble = bool(local_ensemble)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise NotImplementedError("SpaceTimeDecoder is CUDA-only")
        self.precision = precision
        self.params = {}
        for name, n_layers in _SIREN_LAYERS.items():
            for layer in range(n_layers):
                last = layer == n_layers - 1
                for what in ("weight", "bias"):
                    k = _key(name, layer, last, what)
                    if k not in params:
                        raise KeyError(f"missing hot-path parameter {k!r}")
                    t = params[k].detach().to(device=self.device, dtype=torch.float32).contiguous()
                    want = _SIREN_SHAPES[name][layer] if what == "weight" else (_SIREN_SHAPES[name][layer][0],)
                    if tuple(t.shape) != tuple(want):
                        raise ValueError(f"{k}: expected shape {want}, got {tuple(t.shape)}")
                    self.params[k] = t
        if "alpha" not in params:
            raise KeyError("missing hot-path parameter 'alpha'")
        self.alpha = float(params["alpha"].detach().float().reshape(-1)[0].item())
        self._seq_cache = {}
        self._workspace = None
        self._ws_weights_key = None

    @classmethod
    def from_state_dict(cls, state_dict, device="cuda", precision=DEFAULT_PRECISION, local_ensemble: bool = False):
        """Accepts a full ``LunaTokis`` ``state_dict`` (e.g. ``best.pth``, optional ``module.`` prefix)."""
        clean = {}
        for k, v in state_dict.items():
            k = k[7:] if k.startswith("module.") else k
            if k == "alpha" or k.startswith(HOT_PREFIXES):
                clean[k] = v
        return cls(clean, device=device, precision=precision, local_ensemble=local_ensemble)

    # ------------------------------------------------------------------------------------------
    def _sequences(self, H, W, HH, WW):
        key = (H, W, HH, WW)
        if key not in self._seq_cache:
            self._seq_cache[key] = tuple(coord_sequence(n).to(self.device) for n in (HH, WW, H, W))
        return self._seq_cache[key]

    def _geom(self, B, N, H, W, HH, WW):
        s_hh, s_ww, s_h, s_w = self._sequences(H, W, HH, WW)
        g = _lib.GeomT()
        g.B, g.N, g.H, g.W, g.HH, g.WW = B, N, H, W, HH, WW
        g.seq_hh, g.seq_ww, g.seq_h, g.seq_w = s_hh.data_ptr(), s_ww.data_ptr(), s_h.data_ptr(), s_w.data_ptr()
        g.flow_scale = HH / H  # c_float rounds the python double to fp32, as torch does for a scalar operand
        return g

    def _siren(self, name):
        s = _lib.SirenT()
        n_layers = _SIREN_LAYERS[name]
        s.n_layers = n_layers
        for layer in range(n_layers):
            last = layer == n_layers - 1
            s.weight[layer] = self.params[_key(name, layer, last, "weight")].data_ptr()
            s.bias[layer] = self.params[_key(name, layer, last, "bias")].data_ptr()
        return s

    def _get_workspace(self, nbytes):
        if self._workspace is None or self._workspace.numel() < nbytes:
            self._workspace = None
            self._workspace = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        return self._workspace

    # ------------------------------------------------------------------------------------------
    def query_geometry(self, H, W, HH, WW):
        """Nearest-latent index map, shifted coordinates and relative coordinates of every HR query
        (``Ours.py:667-689, 704, 720-722``).  Returns ``iy, ix`` int32 ``[HH*WW]``, ``coord``, ``rel`` ``[HH*WW, 2]``."""
        lib = _lib.load()
        g = self._geom(1, 1, H, W, HH, WW)
        qs = HH * WW
        iy = torch.empty(qs, dtype=torch.int32, device=self.device)
        ix = torch.empty(qs, dtype=torch.int32, device=self.device)
        coord = torch.empty(qs, 2, dtype=torch.float32, device=self.device)
        rel = torch.empty(qs, 2, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            rc = lib.motif_query_geometry(ctypes.byref(g), iy.data_ptr(), ix.data_ptr(), coord.data_ptr(), rel.data_ptr(),
                                          _lib.current_stream_ptr(self.device))
        _lib.check(rc, "motif_query_geometry")
        return iy, ix, coord, rel

    def pack_latents(self, x: torch.Tensor, lr_rows: Optional[Tuple[int, int]] = None) -> torch.Tensor:
        """NCHW ``[R, C, H, W]`` -> pixel-major ``[R, H*W, C]`` (device transpose kernel).  ``lr_rows``: only these LR rows are
        transposed (a destination row band never reads the others; the rest of the result is uninitialised)."""
        lib = _lib.load()
        _lib.require_cuda_f32("latents", x, 4)
        x = x.contiguous()
        r, c, h, w = x.shape
        out = torch.empty(r, h * w, c, dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            if lr_rows is None:
                rc = lib.motif_pack_latents(x.data_ptr(), out.data_ptr(), r, c, h * w, _lib.current_stream_ptr(x.device))
            else:
                rc = lib.motif_pack_latents_range(x.data_ptr(), out.data_ptr(), r, c, h * w, lr_rows[0] * w, lr_rows[1] * w, _lib.current_stream_ptr(x.device))
        _lib.check(rc, "motif_pack_latents")
        return out

    @staticmethod
    def lr_rows_of_band(H: int, HH: int, row_range: Tuple[int, int], halo: int) -> Tuple[int, int]:
        """LR rows whose latents the HR source rows ``[r0 - halo, r1 + halo)`` can select (nearest latent, one row of slack):
        the same bound the library uses for its per-LR-pixel tables."""
        s0, s1 = max(row_range[0] - halo, 0), min(row_range[1] + halo, HH)
        return max((s0 * H) // HH - 1, 0), min(-((-s1 * H) // HH) + 1, H)

    def decode(
        self,
        feat: torch.Tensor,  # [2B, 64, H, W]  F_0^L, F_1^L (leading index r*B + b)
        flow_feat: torch.Tensor,  # [2B, 64, H, W]  T_0^L, T_1^L
        residual: torch.Tensor,  # [B, 64, H, W]   F_01^L
        target_t,  # [B, N] tensor or nested sequence
        hr_size: Tuple[int, int],
        n_range: Optional[Tuple[int, int]] = None,
        return_flow: bool = True,
        debug_synth_in: bool = False,
        precision: Optional[str] = None,
        debug_pre0: bool = False,
        out: Optional[torch.Tensor] = None,
        row_range: Optional[Tuple[int, int]] = None,
        halo: int = 0,
        flow_y_max: Optional[torch.Tensor] = None,
    ):
        """``Ours.py:659-858``.  Returns ``(rgb [N,B,3,HH,WW] in [0,1], flow_out [2BN,2,HH,WW] or None)``
        (plus the ``[B*N,198,HH,WW]`` synth_net input when ``debug_synth_in`` -- precisions ``fp32`` / ``tf32x3`` --
        or the ``[B*N,64,HH,WW]`` layer-0 pre-activation of synth_net when ``debug_pre0`` -- precision ``f16x3``,
        which never forms the 198-channel input).  ``out``: optional preallocated frame buffer (``ClipStream``).

        ``row_range=(r0, r1)`` (multiples of 16, or ``r1 == HH``) decodes only the destination rows ``[r0, r1)`` of every
        frame -- the other rows of ``rgb`` are left untouched -- from the sources of rows ``[r0 - halo, r1 + halo)``
        (the second sharding axis).  Exact iff no source outside them lands in the band, i.e. iff
        ``max |flow_y| < halo - 1`` HR pixels; ``flow_y_max`` (a 64-element fp32 CUDA tensor) receives values whose maximum is
        the largest ``|flow_y|`` over the band's own source rows (``sharding.check_halo`` reduces it over the ranks)."""
        lib = _lib.load()
        for nm, t in (("feat", feat), ("flow_feat", flow_feat), ("residual", residual)):
            _lib.require_cuda_f32(nm, t, 4)
        B = residual.shape[0]
        H, W = residual.shape[-2:]
        if feat.shape != (2 * B, 64, H, W) or flow_feat.shape != (2 * B, 64, H, W) or residual.shape[1] != 64:
            raise ValueError(f"latent shapes do not match: feat {tuple(feat.shape)}, flow_feat {tuple(flow_feat.shape)}, residual {tuple(residual.shape)}")
        HH, WW = int(hr_size[0]), int(hr_size[1])
        tt = torch.as_tensor(target_t, dtype=torch.float32).detach().cpu().reshape(B, -1).contiguous() if not isinstance(target_t, torch.Tensor) \
            else target_t.detach().to(device="cpu", dtype=torch.float32).reshape(B, -1).contiguous()
        N = tt.shape[1]
        n0, n1 = (0, N) if n_range is None else n_range
        dev = self.device
        with torch.cuda.device(dev):
            # f16x3 reads the reference's NCHW tensors as they are (its per-LR-pixel tables are the only readers of the latents);
            # the fp32 / tf32x3 kernels gather pixel-major rows and need the transposed copy
            nchw = (precision or self.precision) == "f16x3" and not getattr(self, "force_packed_latents", False)  # (test hook: the packed form of the C ABI)
            if nchw:
                featp, ffp, resp = feat.contiguous(), flow_feat.contiguous(), residual.contiguous()
            else:
                lr_rows = None if row_range is None else self.lr_rows_of_band(H, HH, row_range, halo)
                featp = self.pack_latents(feat, lr_rows)
                ffp = self.pack_latents(flow_feat, lr_rows)
                resp = self.pack_latents(residual, lr_rows)
            if out is not None:
                if out.shape != (N, B, 3, HH, WW) or out.dtype != torch.float32 or out.device != residual.device or not out.is_contiguous():
                    raise ValueError(f"out must be a contiguous fp32 [{N},{B},3,{HH},{WW}] tensor on {dev}")
                rgb = out
            else:
                rgb = torch.empty(N, B, 3, HH, WW, dtype=torch.float32, device=dev)
            flow_out = torch.empty(2 * B * N, 2, HH, WW, dtype=torch.float32, device=dev) if return_flow else None
            dbg = torch.zeros(B * N, 198, HH, WW, dtype=torch.float32, device=dev) if debug_synth_in else None
            pre0 = torch.zeros(B * N, 64, HH, WW, dtype=torch.float32, device=dev) if debug_pre0 else None
            nbytes = lib.motif_decode_workspace_bytes(B, N, H, W, HH, WW)
            ws = self._get_workspace(nbytes)
            a = _lib.DecodeT()
            a.geom = self._geom(B, N, H, W, HH, WW)
            a.feat, a.flow_feat, a.residual = featp.data_ptr(), ffp.data_ptr(), resp.data_ptr()
            tt_arr = (ctypes.c_float * (B * N))(*tt.reshape(-1).tolist())
            a.target_t = ctypes.cast(tt_arr, ctypes.POINTER(ctypes.c_float))
            a.imnet, a.flow_imnet, a.synth_net = self._siren("imnet"), self._siren("flow_imnet"), self._siren("synth_net")
            a.alpha = self.alpha
            a.rgb = rgb.data_ptr()
            a.flow_out = flow_out.data_ptr() if flow_out is not None else None
            a.workspace, a.workspace_bytes = ws.data_ptr(), ws.numel()
            a.dbg_synth_in = dbg.data_ptr() if dbg is not None else None
            a.dbg_pre0 = pre0.data_ptr() if pre0 is not None else None
            a.n_begin, a.n_end = int(n0), int(n1)
            a.precision = PRECISIONS[precision or self.precision]
            a.local_ensemble = int(self.local_ensemble)
            a.latents_nchw = int(nchw)
            if row_range is not None:
                a.row_begin, a.row_end, a.halo = int(row_range[0]), int(row_range[1]), int(halo)
                if flow_y_max is not None:
                    if flow_y_max.numel() != 64 or flow_y_max.dtype != torch.float32 or flow_y_max.device != residual.device:
                        raise ValueError("flow_y_max must be a 64-element fp32 tensor on the decode device")
                    a.flow_y_max = flow_y_max.data_ptr()
            # the weight images in the workspace are clip-invariant: repacked only when the workspace or the arithmetic changed
            # (the parameters may alias live model tensors -- .to() is a no-op for cuda fp32 -- so their version counters are
            # part of the key: an in-place update by the owner of the weights forces a repack)
            ws_key = (ws.data_ptr(), a.precision, self.alpha, tuple(t._version for t in self.params.values()))
            a.weights_ready = int(self._ws_weights_key == ws_key)
            self._ws_weights_key = None  # a failing call leaves the workspace in an unknown state
            rc = lib.motif_decode(ctypes.byref(a), _lib.current_stream_ptr(dev))
        _lib.check(rc, "motif_decode")
        self._ws_weights_key = ws_key
        if debug_synth_in:
            return rgb, flow_out, dbg
        if debug_pre0:
            return rgb, flow_out, pre0
        return rgb, flow_out
