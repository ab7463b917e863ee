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
    Returns ``(rgb [N,B,3,HH,WW], flow_out [2BN,2,HH,WW], aux)`` on the device; ``aux['flow_hr'] [2BN,2,HH,WW]`` are the HR
    flows and ``aux['wz'] [N,B,1,HH,WW]`` the blended normaliser ``W_0 + W_1`` (``Ours.py:812``) the discontinuity masks of
    the parity tests are built from."""
    dev = feat.device
    assert dev.type == "cuda"
    p = {k: v.to(dev) for k, v in params.items()}
    B, N = target_t.shape
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False  # the reference runs plain fp32 (torch default)
    try:
        with torch.no_grad():
            rgbs, fos, fhs, wzs = [], [], [], []
            step = N if chunk <= 0 else chunk
            for n0 in range(0, N, step):
                r, fo, inter = decoder_ref.decode(feat, flow_feat, residual, target_t[:, n0:n0 + step], HH, WW, p, return_intermediates=True, splat_ops=RefKernelSplats)
                n = r.shape[0]
                rgbs.append(r)
                fos.append(fo.reshape(2 * B, n, 2, HH, WW))
                fhs.append(inter["flow_hr"].reshape(2 * B, n, 2, HH, WW))
                wzs.append(inter["splat_norm"].reshape(2, B, n, 1, HH, WW).sum(0).permute(1, 0, 2, 3, 4))
                del inter
            aux = {"flow_hr": torch.cat(fhs, 1).reshape(2 * B * N, 2, HH, WW), "wz": torch.cat(wzs, 0)}
            return torch.cat(rgbs, 0), torch.cat(fos, 1).reshape(2 * B * N, 2, HH, WW), aux
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev


def equality_unstable_mask(wz: torch.Tensor, alpha: float = 0.0, z_err: float = 0.0, ulps: int = 8) -> torch.Tensor:
    """Destinations at which the reference's exact-equality tests on the blended normaliser are undecided by rounding.

    ``Ours.py:813`` replaces ``Wz == 0`` by 1 and ``:829`` replaces ``Wz == 1.0`` by 0 before forming the ``Wz / count``
    input of ``synth_net`` (``:834``).  ``Wz`` is a sum of float ``atomicAdd``s (``softsplat_cp.py:40-51``) of
    ``exp(relu(z_raw) * alpha) * w``: wherever the flow is locally a translation and ``z_raw <= 0`` the terms add up to
    1 +- a few ulp in an order the hardware picks, and a term whose ``z_raw`` is within the evaluation error ``z_err`` of the
    relu kink contributes ``exp(alpha * z_err) - 1`` more or less -- so two correct evaluations land on different sides of
    the ``== 1.0`` test, ``Wz / count`` jumps between ``1 / count`` and 0 and RGB moves by O(1e-2).  The excluded set is
    ``|Wz - 1| <= ulps * 2^-23 + 4 * |alpha| * z_err`` (and ``|Wz|`` tiny).  ``wz`` ``[N,B,1,HH,WW]``; returns bool."""
    eps = ulps * 2.0 ** -23 + 4.0 * abs(alpha) * z_err
    return ((wz - 1.0).abs() <= eps) | (wz.abs() <= ulps * 2.0 ** -23 * 1e-3)


def count_unstable_mask(flow_hr, B, N, eps: float = 2.5e-4):
    return decoder_ref.count_unstable_mask(flow_hr, B, N, eps, splat_ops=RefKernelSplats)
