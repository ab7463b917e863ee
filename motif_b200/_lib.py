"""ctypes binding of ``libmotif_b200.so`` (the C ABI declared in ``include/motif_b200.h``).

The library is built in-tree by ``motif_b200/build.py`` (``nvcc -gencode
arch=compute_100a,code=sm_100a``).  There is no CPU or PyTorch fallback: every
operator of this package raises ``MotifLibraryError`` when the library is
missing, and ``MotifError`` when a call returns a non-zero status.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_float, c_int, c_int32, c_longlong, c_size_t, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libmotif_b200.so")

# every symbol include/motif_b200.h declares (checked by tests/test_abi.py)
EXPORTS = [
    "motif_abi_version",
    "motif_memcpy2d_async",
    "motif_last_error",
    "motif_launch_count",
    "motif_reset_launch_count",
    "motif_prof_enable",
    "motif_prof_collect",
    "motif_splat_workspace_bytes",
    "motif_splat_fwd",
    "motif_splat_fwd_atomic",
    "motif_splat_max_fwd",
    "motif_splat_count_fwd",
    "motif_corr_fwd",
    "motif_flow_front",
    "motif_raft_corr_lookup",
    "motif_raft_corr_lookup_pyramid",
    "motif_dcn_v2_fwd",
    "motif_frame_metrics",
    "motif_query_geometry",
    "motif_pack_latents",
    "motif_pack_latents_range",
    "motif_decode_workspace_bytes",
    "motif_sizeof_decode_t",
    "motif_decode",
    "motif_tc_selftest",
    "motif_tc_set_trace",
    "motif_tc_mma_rate",
    "motif_tc_wait_debug_buffer",
]

SPLAT_MODES = {"summation": 0, "average": 1, "linear": 2, "softmax": 3}


class MotifLibraryError(RuntimeError):
    pass


class MotifError(RuntimeError):
    pass


class SirenT(Structure):
    _fields_ = [("n_layers", c_int), ("weight", c_void_p * 5), ("bias", c_void_p * 5)]


class GeomT(Structure):
    _fields_ = [
        ("B", c_int), ("N", c_int), ("H", c_int), ("W", c_int), ("HH", c_int), ("WW", c_int),
        ("seq_hh", c_void_p), ("seq_ww", c_void_p), ("seq_h", c_void_p), ("seq_w", c_void_p),
        ("flow_scale", c_float),
    ]


class DecodeT(Structure):
    _fields_ = [
        ("geom", GeomT),
        ("feat", c_void_p), ("flow_feat", c_void_p), ("residual", c_void_p),
        ("target_t", POINTER(c_float)),
        ("imnet", SirenT), ("flow_imnet", SirenT), ("synth_net", SirenT),
        ("alpha", c_float),
        ("rgb", c_void_p), ("flow_out", c_void_p),
        ("workspace", c_void_p), ("workspace_bytes", c_size_t),
        ("dbg_synth_in", c_void_p),
        ("dbg_pre0", c_void_p),
        ("n_begin", c_int), ("n_end", c_int),
        ("precision", c_int),
        ("local_ensemble", c_int),
        ("row_begin", c_int), ("row_end", c_int), ("halo", c_int),
        ("flow_y_max", c_void_p),
        ("weights_ready", c_int),
        ("latents_nchw", c_int),
    ]


_lib = None


def _declare(lib):
    lib.motif_abi_version.restype = c_int
    lib.motif_memcpy2d_async.restype = c_int
    lib.motif_memcpy2d_async.argtypes = [c_void_p, c_size_t, c_void_p, c_size_t, c_size_t, c_size_t, c_int, c_void_p]
    lib.motif_last_error.restype = c_char_p
    lib.motif_launch_count.restype = c_longlong
    lib.motif_reset_launch_count.restype = None
    lib.motif_prof_enable.restype = None
    lib.motif_prof_enable.argtypes = [c_int]
    lib.motif_prof_collect.restype = c_int
    lib.motif_prof_collect.argtypes = [POINTER(c_char_p), c_int, POINTER(ctypes.c_double), POINTER(c_longlong)]
    lib.motif_splat_workspace_bytes.restype = c_size_t
    lib.motif_splat_workspace_bytes.argtypes = [c_int, c_int, c_int]
    lib.motif_splat_fwd.restype = c_int
    lib.motif_splat_fwd.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_size_t, c_void_p]
    lib.motif_splat_fwd_atomic.restype = c_int
    lib.motif_splat_fwd_atomic.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]
    lib.motif_splat_max_fwd.restype = c_int
    lib.motif_splat_max_fwd.argtypes = [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]
    lib.motif_splat_count_fwd.restype = c_int
    lib.motif_splat_count_fwd.argtypes = [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]
    lib.motif_corr_fwd.restype = c_int
    lib.motif_corr_fwd.argtypes = [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]
    lib.motif_flow_front.restype = c_int
    lib.motif_flow_front.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]
    lib.motif_raft_corr_lookup.restype = c_int
    lib.motif_raft_corr_lookup.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]
    lib.motif_raft_corr_lookup_pyramid.restype = c_int
    lib.motif_raft_corr_lookup_pyramid.argtypes = [c_void_p, POINTER(c_void_p), POINTER(c_int), POINTER(c_int), c_int, c_void_p, c_void_p,
                                                   c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]
    lib.motif_dcn_v2_fwd.restype = c_int
    lib.motif_dcn_v2_fwd.argtypes = [c_void_p] * 6 + [c_int] * 6 + [c_void_p]
    lib.motif_frame_metrics.restype = c_int
    lib.motif_frame_metrics.argtypes = [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]
    lib.motif_query_geometry.restype = c_int
    lib.motif_query_geometry.argtypes = [POINTER(GeomT), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]
    lib.motif_pack_latents.restype = c_int
    lib.motif_pack_latents.argtypes = [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]
    lib.motif_pack_latents_range.restype = c_int
    lib.motif_pack_latents_range.argtypes = [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]
    lib.motif_decode_workspace_bytes.restype = c_size_t
    lib.motif_decode_workspace_bytes.argtypes = [c_int, c_int, c_int, c_int, c_int, c_int]
    lib.motif_sizeof_decode_t.restype = c_size_t
    lib.motif_tc_set_trace.restype = c_int
    lib.motif_tc_set_trace.argtypes = [c_void_p, c_int]
    lib.motif_tc_mma_rate.restype = c_int
    lib.motif_tc_mma_rate.argtypes = [c_void_p, c_int, c_int, c_int, c_int, c_void_p]
    lib.motif_tc_selftest.restype = c_int
    lib.motif_tc_selftest.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p]
    lib.motif_tc_wait_debug_buffer.restype = POINTER(ctypes.c_uint)
    lib.motif_decode.restype = c_int
    lib.motif_decode.argtypes = [POINTER(DecodeT), c_void_p]


def load():
    """Load the shared library (once).  Raises ``MotifLibraryError`` when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise MotifLibraryError(
                f"{LIB_PATH} is missing: build it with `python -m motif_b200.build` "
                "(there is no CPU / PyTorch fallback for the motif_b200 operators)"
            )
        lib = ctypes.CDLL(LIB_PATH)
        _declare(lib)
        _lib = lib
    return _lib


def check(rc: int, what: str):
    if rc != 0:
        msg = load().motif_last_error().decode("utf-8", "replace")
        raise MotifError(f"{what} failed (status {rc}): {msg}")


def current_stream_ptr(device=None):
    import torch

    return c_void_p(torch.cuda.current_stream(device).cuda_stream)


def require_cuda_f32(name, t, dims=None):
    import torch

    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name}: expected a torch.Tensor")
    if not t.is_cuda:
        # the reference raises here as well (assert is_cuda / NotImplementedError, softsplat_cp.py:232-252)
        raise NotImplementedError(f"{name}: motif_b200 operators are CUDA-only (got a {t.device} tensor)")
    if t.dtype != torch.float32:
        raise TypeError(f"{name}: expected float32, got {t.dtype}")
    if dims is not None and t.dim() != dims:
        raise ValueError(f"{name}: expected {dims} dimensions, got {tuple(t.shape)}")


def prof_enable(on: bool):
    load().motif_prof_enable(1 if on else 0)


def prof_collect(names):
    """{name: (total_ms, launches)} of the kernels recorded since prof_enable(True)."""
    lib = load()
    arr = (c_char_p * len(names))(*[n.encode() for n in names])
    ms = (ctypes.c_double * len(names))()
    cnt = (c_longlong * len(names))()
    lib.motif_prof_collect(arr, len(names), ms, cnt)
    return {n: (ms[i], cnt[i]) for i, n in enumerate(names)}
