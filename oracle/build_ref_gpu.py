"""Compile the reference's own forward kernels for the B200 (TEST / MEASUREMENT INFRASTRUCTURE ONLY).

Same sources as ``build_ref.py`` -- the five CUDA-C kernel strings of ``models/softsplat_cp.py:12-52``,
``softsplat_max_cp.py:12-58``, ``softsplat_count_cp.py:14-52`` and ``OpticalFlow/correlation.py:17-112``, read from
the reference checkout where it lies and expanded by the reference's own ``cupy_kernel`` pre-processor (which bakes
sizes and strides in per shape) -- but compiled UNMODIFIED for the device with ``nvcc -arch sm_100a`` instead of for
the host, for a fixed list of shapes, into ``oracle/_ref/ref_gpu_sm100a.so`` (git-ignored, travels to the GPU box).
Each kernel gets a unique symbol per shape (the generated text is renamed, nothing else is touched) and an
``extern "C"`` launcher that uses the reference's own grid / block / shared-memory configuration
(``softsplat_cp.py:244-249``, ``correlation.py:306-341``).

This is "the reference GPU path" of BASELINE.md section 4: what the reference would run on this B200 if cupy were
installed.  ``tests/test_ref_gpu.py`` compares the product kernels against it at full Adobe size;
``tools/bench_ref_gpu.py`` times it beside them.  No reference source is copied into the repository.
"""
from __future__ import annotations

import ctypes
import hashlib
import os
import subprocess

import torch

from oracle import build_ref

SO = os.path.join(build_ref.OUT_DIR, "ref_gpu_sm100a.so")

# (tag, N, C, H, W): sum splat with the normaliser channel (C = 130 + 1), max / count (C = 1)
# x3p5 / uhdq: HR sizes of BASELINE configs 3 (630x1120) and of a quarter-area crop of config 4 (LR 270x480 -> 1080x1920)
SPLAT_SHAPES = [("adobe", 1, 131, 720, 1280), ("vimeo", 1, 131, 256, 448), ("small", 2, 6, 37, 52), ("x3p5", 1, 131, 630, 1120), ("uhdq", 1, 131, 1080, 1920)]
UNIT_SHAPES = [("adobe", 1, 1, 720, 1280), ("vimeo", 1, 1, 256, 448), ("small", 2, 1, 37, 52), ("x3p5", 1, 1, 630, 1120), ("uhdq", 1, 1, 1080, 1920)]
# PWC-Net pyramid levels of a 720x1280 pair and a small case
CORR_SHAPES = [("l2", 1, 32, 192, 320), ("l3", 1, 64, 96, 160), ("l6", 1, 196, 12, 20), ("small", 2, 16, 12, 20)]

_SPLAT_LAUNCH = """
extern "C" int launch_{name}(int n, const float* a, const float* b, float* c, void* stream) {{
  {name}<<<(n + 511) / 512, 512, 0, (cudaStream_t)stream>>>(n, a, b, c);
  return (int)cudaGetLastError();
}}
"""
_CORR_LAUNCH = """
extern "C" int launch_corr_{tag}(const float* first, const float* second, float* rbot0, float* rbot1, float* top, void* stream) {{
  const int n = {h} * {w};
  corr_rearrange_{tag}<<<dim3((n + 15) / 16, {c}, {b}), 16, 0, (cudaStream_t)stream>>>(n, first, rbot0);
  corr_rearrange_{tag}<<<dim3((n + 15) / 16, {c}, {b}), 16, 0, (cudaStream_t)stream>>>(n, second, rbot1);
  corr_update_{tag}<<<dim3({w}, {h}, {b}), 32, {c} * 4, (cudaStream_t)stream>>>(81 * n, rbot0, rbot1, top);
  return (int)cudaGetLastError();
}}
"""


def _sources() -> str:
    parts = ["#include <cuda_runtime.h>\n#include <assert.h>\n"]
    for module, prefix, shapes in (("models.softsplat_cp", "splat_sum", SPLAT_SHAPES), ("models.softsplat_max_cp", "splat_max", UNIT_SHAPES),
                                   ("models.softsplat_count_cp", "splat_count", UNIT_SHAPES)):
        mod = build_ref._ref_module(module)
        for tag, n, c, h, w in shapes:
            x = torch.empty(n, c, h, w)
            f = torch.empty(n, 2, h, w)
            src = mod.cupy_kernel("kernel_Softsplat_updateOutput", {"input": x, "flow": f, "output": x})
            name = f"{prefix}_{tag}"
            # the max file defines a helper (atomicMaxFloat) once per expansion: make it unique as well
            src = src.replace("kernel_Softsplat_updateOutput", name).replace("atomicMaxFloat", f"atomicMaxFloat_{tag}")
            parts.append(src + _SPLAT_LAUNCH.format(name=name))
    mod = build_ref._ref_module("OpticalFlow.correlation")
    for tag, b, c, h, w in CORR_SHAPES:
        first = torch.empty(b, c, h, w)
        rbot = torch.empty(b, h + 8, w + 8, c)
        top = torch.empty(b, 81, h, w)
        src_r = mod.cupy_kernel("kernel_Correlation_rearrange", {"input": first, "output": rbot}).replace("kernel_Correlation_rearrange", f"corr_rearrange_{tag}")
        src_c = mod.cupy_kernel("kernel_Correlation_updateOutput", {"rbot0": rbot, "rbot1": rbot, "top": top}).replace(
            "kernel_Correlation_updateOutput", f"corr_update_{tag}")
        parts.append(src_r + "\n" + src_c + _CORR_LAUNCH.format(tag=tag, b=b, c=c, h=h, w=w))
    return "\n".join(parts)


def build() -> str | None:
    """nvcc cross-compiles without a GPU; a no-op when the reference checkout is absent (GPU box: prebuilt file)."""
    if not build_ref.reference_available():
        return SO if os.path.exists(SO) else None
    os.makedirs(build_ref.OUT_DIR, exist_ok=True)
    cu = os.path.join(build_ref.OUT_DIR, "ref_gpu_sm100a.cu")
    text = _sources()
    digest = hashlib.sha1(text.encode()).hexdigest()
    if os.path.exists(SO) and os.path.exists(cu + ".sha") and open(cu + ".sha").read() == digest:
        return SO
    with open(cu, "w") as f:
        f.write(text)
    subprocess.check_call([os.environ.get("NVCC", "nvcc"), "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-w", "-Xcompiler", "-fPIC", "-shared", cu, "-o", SO])
    os.remove(cu)  # generated from the reference's strings: not kept
    with open(cu + ".sha", "w") as f:
        f.write(digest)
    return SO


_lib = None


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(SO):
            return None
        _lib = ctypes.CDLL(SO)
    return _lib


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def splat(kind: str, tag: str, inp: torch.Tensor, flow: torch.Tensor) -> torch.Tensor:
    """Run the reference kernel ``kind`` in {'sum', 'max', 'count'} (shape ``tag``) on CUDA tensors, with the output
    initialisation of the reference wrapper (zeros; ones for max, ``softsplat_max_cp.py:254``)."""
    lib = load()
    fn = getattr(lib, f"launch_splat_{kind}_{tag}")
    fn.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    out = torch.ones_like(inp) if kind == "max" else torch.zeros_like(inp)
    rc = fn(out.numel(), inp.data_ptr(), flow.data_ptr(), out.data_ptr(), _stream())
    assert rc == 0, rc
    return out


def correlation(tag: str, first: torch.Tensor, second: torch.Tensor) -> torch.Tensor:
    lib = load()
    fn = getattr(lib, f"launch_corr_{tag}")
    fn.argtypes = [ctypes.c_void_p] * 6
    b, c, h, w = first.shape
    rbot0 = first.new_zeros(b, h + 8, w + 8, c)
    rbot1 = first.new_zeros(b, h + 8, w + 8, c)
    out = first.new_zeros(b, 81, h, w)
    rc = fn(first.data_ptr(), second.data_ptr(), rbot0.data_ptr(), rbot1.data_ptr(), out.data_ptr(), _stream())
    assert rc == 0, rc
    return out


if __name__ == "__main__":
    print(build())
