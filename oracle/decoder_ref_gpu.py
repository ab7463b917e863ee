"""The reference's GPU path for the decoder (TEST / MEASUREMENT INFRASTRUCTURE ONLY).

``oracle.decoder_ref.decode`` -- the restatement of ``models/modules/Ours.py:659-858`` in the reference's own eager torch
operators -- executed on CUDA tensors, with the three forward splats done by the reference's OWN kernels compiled
unmodified for sm_100a (``oracle/build_ref_gpu.py`` -> ``oracle/_ref/ref_gpu_sm100a.so``).  This is what the reference
would run on this B200 if cupy were installed (SURVEY 8d "the number to beat"): the full-size
parity oracle of ``tests/test_decoder_fullsize_gpu.py`` (the CPU oracle needs ~4 s per Adobe timestamp) and the
``gpu_reference`` leg of ``bench.py`` / ``tools/bench_gpu_reference.py``.  Never imported by ``motif_b200/``.

The reference kernels bake sizes in per shape, so only the HR sizes listed in ``build_ref_gpu.SPLAT_SHAPES`` (one sample
per launch) are available; samples are looped over.
"""
from __future__ import annotations

import torch

from . import build_ref_gpu, decoder_ref


def available() -> bool:
    return torch.cuda.is_available() and build_ref_gpu.load() is not None


def _tag(h: int, w: int) -> str:
    for tag, n, c, hh, ww in build_ref_gpu.SPLAT_SHAPES:
        if n == 1 and (hh, ww) == (h, w):
            return tag
    raise KeyError(f"no reference kernel was compiled for HR size {h}x{w} (oracle/build_ref_gpu.py: SPLAT_SHAPES)")


class RefKernelSplats:
    """``function_softsplat*`` of ``oracle/softsplat_ref.py`` with the reference kernels doing the scatter."""

    @staticmethod
    def function_softsplat(tenInput, tenFlow, tenMetric, strType):
        assert strType == "softmax" and tenInput.shape[1] == 130
        tag = _tag(*tenInput.shape[-2:])
        e = tenMetric.exp()
        outs = []
        for i in range(tenInput.shape[0]):  # softsplat_cp.py:332-346: [in * exp(metric) | exp(metric)] formed by torch, then the kernel
            ref_in = torch.cat([tenInput[i:i + 1] * e[i:i + 1], e[i:i + 1]], 1).contiguous()
            outs.append(build_ref_gpu.splat("sum", tag, ref_in, tenFlow[i:i + 1].contiguous()))
        out = torch.cat(outs, 0)
        return out[:, :-1], out[:, -1:]

    @staticmethod
    def function_softsplat_max(tenInput, tenFlow):
        tag = _tag(*tenInput.shape[-2:])
        return torch.cat([build_ref_gpu.splat("max", tag, tenInput[i:i + 1].contiguous(), tenFlow[i:i + 1].contiguous()) for i in range(tenInput.shape[0])], 0)

    @staticmethod
    def function_softsplat_count(tenInput, tenFlow):
        tag = _tag(*tenFlow.shape[-2:])
        ones = tenFlow.new_ones(1, 1, *tenFlow.shape[-2:])  # softsplat_count_cp.py:163-165
        return torch.cat([build_ref_gpu.splat("count", tag, ones, tenFlow[i:i + 1].contiguous()) for i in range(tenFlow.shape[0])], 0)


def decode(feat, flow_feat, residual, target_t, HH, WW, params, chunk: int = 0):
    """Reference eager GPU forward of the hot path.  ``chunk`` > 0 decodes that many timestamps per pass, re-running the
    clip-invariant part each time -- exactly what ``VideoSRBaseModel.test`` does with chunks of three
    (``VideoSR_base_model.py:188-193``) -- which also bounds memory; 0 = all timestamps in one pass.
    Returns ``(rgb [N,B,3,HH,WW], flow_out [2BN,2,HH,WW], flow_hr [2BN,2,HH,WW])`` on the device."""
    dev = feat.device
    assert dev.type == "cuda"
    p = {k: v.to(dev) for k, v in params.items()}
    B, N = target_t.shape
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False  # the reference runs plain fp32 (torch default)
    try:
        with torch.no_grad():
            if chunk <= 0 or chunk >= N:
                rgb, flow_out, inter = decoder_ref.decode(feat, flow_feat, residual, target_t, HH, WW, p, return_intermediates=True, splat_ops=RefKernelSplats)
                return rgb, flow_out, inter["flow_hr"]
            rgbs, fos, fhs = [], [], []
            for n0 in range(0, N, chunk):
                r, fo, inter = decoder_ref.decode(feat, flow_feat, residual, target_t[:, n0:n0 + chunk], HH, WW, p, return_intermediates=True, splat_ops=RefKernelSplats)
                n = r.shape[0]
                rgbs.append(r)
                fos.append(fo.reshape(2 * B, n, 2, HH, WW))
                fhs.append(inter["flow_hr"].reshape(2 * B, n, 2, HH, WW))
                del inter
            return torch.cat(rgbs, 0), torch.cat(fos, 1).reshape(2 * B * N, 2, HH, WW), torch.cat(fhs, 1).reshape(2 * B * N, 2, HH, WW)
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev


def count_unstable_mask(flow_hr, B, N, eps: float = 2.5e-4):
    return decoder_ref.count_unstable_mask(flow_hr, B, N, eps, splat_ops=RefKernelSplats)
