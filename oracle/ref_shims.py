"""Shims that let the UNMODIFIED reference checkout import and run on a CPU-only box.

TEST INFRASTRUCTURE ONLY (used by ``oracle/make_golden.py`` and by the tests that
are skipped when ``/root/reference`` is absent, e.g. on the GPU box).  Nothing is
copied from the reference; its modules are imported from where they lie.

The six shims of SURVEY.md section 8c:

1. ``cupy`` stub (``memoize``, ``int32``, ``ndarray`` -- einops probes the latter);
2. ``alt_cuda_corr`` stub + ``args.alternate_corr=False`` so RAFT uses the in-repo
   ``CorrBlock`` (``models/core/corr.py:8-56``);
3. ``_ext`` stub + ``dcn_v2_conv`` -> ``torchvision.ops.deform_conv2d``
   (``models/modules/DCNv2/dcn_v2.py:50``);
4. ``torch.load`` of the author-machine RAFT checkpoint (``Ours.py:424``) answered with
   a seeded RAFT-small state dict;
5. ``torch.cuda.FloatTensor`` -> CPU tensor, ``Tensor.cuda`` -> identity
   (``Ours.py:443, 621, 677``; ``convlstm.py:62-63``);
6. the three ``_FunctionSoftsplat`` launchers answered by ``oracle/softsplat_ref.py``
   (or, when built, by the reference's own kernel strings compiled for the host,
   ``oracle/build_ref.py``).
"""
from __future__ import annotations

import contextlib
import os
import sys
import types

import torch

REFERENCE_ROOT = os.environ.get("MOTIF_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "models", "modules"))


def _install_module_stubs():
    if "cupy" not in sys.modules:
        cupy = types.ModuleType("cupy")

        def memoize(for_each_device=False):
            def deco(fn):
                cache = {}

                def wrapped(*a):
                    if a not in cache:
                        cache[a] = fn(*a)
                    return cache[a]

                return wrapped

            return deco

        class _NdArray:  # einops' backend probe does isinstance(x, cupy.ndarray)
            pass

        cupy.memoize = memoize
        cupy.int32 = int
        cupy.ndarray = _NdArray
        cupy.RawModule = None
        sys.modules["cupy"] = cupy
    for name in ("alt_cuda_corr", "_ext"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)


@contextlib.contextmanager
def cpu_cuda_aliases():
    """Shim 5: make ``.cuda()`` and ``torch.cuda.FloatTensor`` harmless on a CPU box."""
    orig_cuda = torch.Tensor.cuda
    orig_ft = getattr(torch.cuda, "FloatTensor", None)
    torch.Tensor.cuda = lambda self, *a, **k: self

    def _float_tensor(data, device=None):
        return torch.tensor(data, dtype=torch.float32)

    torch.cuda.FloatTensor = _float_tensor
    try:
        yield
    finally:
        torch.Tensor.cuda = orig_cuda
        if orig_ft is not None:
            torch.cuda.FloatTensor = orig_ft


def import_reference():
    """Import the reference's ``models.modules.Ours`` from ``REFERENCE_ROOT`` under shims 1-3."""
    if not reference_available():
        raise RuntimeError(f"reference checkout not found at {REFERENCE_ROOT}")
    _install_module_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import warnings

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import torchvision.ops as tvo
        import models.modules.DCNv2.dcn_v2 as dcn_mod

        def dcn_v2_conv(inp, offset, mask, weight, bias, stride, padding, dilation, groups):
            return tvo.deform_conv2d(inp, offset, weight, bias, stride, padding, dilation, mask)

        dcn_mod.dcn_v2_conv = dcn_v2_conv
        import models.modules.Ours as ours
    return ours


def build_reference_model(seed: int = 0, splat_backend: str = "restatement"):
    """Instantiate ``LunaTokis(setting=5)`` on the CPU with seeded weights (shims 4-6)."""
    ours = import_reference()
    from oracle import softsplat_ref as S

    torch.manual_seed(seed)
    real_load = torch.load

    def fake_load(path, *a, **k):
        if isinstance(path, str) and path.endswith("raft_smooth_0728_iter12.pth"):
            import argparse

            args = argparse.Namespace(small=True, mixed_precision=False, alternate_corr=False)
            raft = ours.RAFT(args)
            return {"model": {"flow_predictor." + k_: v for k_, v in raft.state_dict().items()}}
        return real_load(path, *a, **k)

    torch.load = fake_load
    try:
        with cpu_cuda_aliases():
            model = ours.LunaTokis(setting=5)
    finally:
        torch.load = real_load
    model.flow_predictor.args.alternate_corr = False

    # shim 6: answer the three cupy launchers on the CPU
    import models.softsplat_cp as sp
    import models.softsplat_max_cp as spm
    import models.softsplat_count_cp as spc

    if splat_backend == "restatement":
        fns = (S.splat_sum, S.splat_max, S.splat_count)
    else:
        from oracle import build_ref

        fns = (build_ref.ref_splat_sum, build_ref.ref_splat_max, build_ref.ref_splat_count)

    def make_apply(fn):
        class _F:
            @staticmethod
            def apply(inp, flow):
                return fn(inp.contiguous(), flow.contiguous())

        return _F

    sp._FunctionSoftsplat = make_apply(fns[0])
    spm._FunctionSoftsplat = make_apply(fns[1])
    spc._FunctionSoftsplat = make_apply(fns[2])
    model.eval()
    return model


def run_reference_forward(model, x, target_t, scale, iters=4):
    """``LunaTokis.forward(x, None, target_t, scale, use_GT=False, iter=4)`` on the CPU,
    also capturing the hot-path inputs (encoder output, ``flow_process`` output)."""
    captured = {}

    def enc_hook(_m, _i, out):
        captured["encoder_out"] = out.detach().clone()

    def fp_hook(_m, _i, out):
        captured["flow_feat"] = out.detach().clone()
        captured["flow_process_in"] = _i[0].detach().clone()  # [2B,14,H,W]: flow / 20, psi maps, durations (Ours.py:614-637)

    def raft_hook(_m, _i, out):
        captured["flow_hr"] = out[-1].detach().clone()  # [4B,2,HH,WW] (Ours.py:545-546)

    h1 = model.encoder.register_forward_hook(enc_hook)
    h2 = model.flow_process.register_forward_hook(fp_hook)
    h3 = model.flow_predictor.register_forward_hook(raft_hook)
    try:
        with torch.no_grad(), cpu_cuda_aliases():
            out, flow, flow_gt = model(x, None, target_t, scale, use_GT=False, iter=iters)
    finally:
        h1.remove()
        h2.remove()
        h3.remove()
    enc = captured["encoder_out"]  # [B,3,64,H,W]
    B = enc.shape[0]
    H, W = enc.shape[-2:]
    residual = enc[:, enc.shape[1] // 2].reshape(B, -1, H, W)
    feat = torch.cat((enc[:, enc.shape[1] // 2 - 1], enc[:, enc.shape[1] // 2 + 1]), 0)
    return {"out": out, "flow_out": flow, "feat": feat, "flow_feat": captured["flow_feat"], "residual": residual,
            "flow_process_in": captured["flow_process_in"], "flow_hr": captured["flow_hr"]}
