"""Bind the sm_100a decoder into a reference ``LunaTokis`` instance (``models/modules/Ours.py``).

    from models.modules.Ours import LunaTokis          # the user's reference checkout
    import motif_b200.luna_tokis as mb
    model = LunaTokis(setting=5); model.load_state_dict(torch.load("best.pth"))   # unchanged
    mb.install(model)                                   # forward() now runs the B200 hot path

``install`` keeps the module tree (and therefore the ``state_dict`` layout ``best.pth`` needs) and
replaces ``forward`` by ``forward_b200``, which

* runs the surround exactly as the reference does, through the instance's own sub-modules
  (RAFT on the four frame pairs, ``ZSM_encoder``, ``flow_process``; ``Ours.py:512-638``) -- glue re-stated here
  because the reference's ``forward`` is monolithic; on a CUDA device the psi reliability maps and the assembly of the
  ``flow_process`` input (``Ours.py:562-578, 613-637``) run as one kernel (``motif_b200.flow_front``);
* hands the three LR latents to ``SpaceTimeDecoder`` (``Ours.py:659-858`` on sm_100a) and returns the
  reference's triple ``(clamp(out) [N,B,3,HH,WW], flow / 20 / (HH/H), flow_GT)`` (``Ours.py:858``).

Also swaps the three splat modules (``self.fwarp*``) for the motif_b200 operators, so any other
caller of them inside the model goes through the same library.

Inference only (``use_GT=False``, ``eval()``; the reference's ``test()`` path,
``VideoSR_base_model.py:169-195``); training-time teacher forcing raises.
"""
from __future__ import annotations

import types

import torch
from torch.nn.functional import interpolate

from . import alt_cuda_corr, dcn_v2, raft_schedule
from .decoder import SpaceTimeDecoder, hr_size_from_scale
from .flow_front import flow_front
from .softsplat_count_cp import Softsplat_Count
from .softsplat_cp import Softsplat
from .softsplat_max_cp import Softsplat_Max


def surround(self, x, target_t, scale, iter=12, front=flow_front):
    """``Ours.py:512-638``: everything before the hot path.  Returns
    ``(feat [2B,64,H,W], flow_feat [2B,64,H,W], residual [B,64,H,W], target_t [B,N], (HH, WW))``.

    ``front``: the operator that turns the LR frames and flows into the ``flow_process`` input (``Ours.py:562-578,
    613-637``); the product's is the fused CUDA kernel ``motif_b200.flow_front`` (CUDA tensors only -- there is no eager
    branch here; the CPU glue test passes the oracle's restatement, ``oracle/flow_front_ref.py``)."""
    if self.trans or not self.input_Z:
        raise NotImplementedError("motif_b200 runs the shipped front end (trans=False, input_Z=True; Ours.py:613-637)")
    x = x.permute(0, 2, 1, 3, 4)
    x = x[:, :, x.shape[2] // 2 - 1:x.shape[2] // 2 + 1]
    with torch.no_grad():
        target_t = torch.stack(target_t, 1).squeeze(-1)
        B, N = target_t.shape
        B, _, _, H, W = x.shape
        HH, WW = hr_size_from_scale(H, W, scale)

        # HR input motion from the pretrained RAFT on the pairs 00, 01, 10, 11 (Ours.py:540-555)
        x_norm = interpolate(x.reshape(B, -1, H, W), size=(HH, WW), mode="bilinear", align_corners=False).reshape(B, -1, 2, HH, WW)
        fr0, fr1 = x_norm[:, :, 0], x_norm[:, :, 1]
        # (raft_schedule: only the pairs 01 and 10 are estimated -- the reference zeroes 00 and 11 right below -- and the
        #  feature encoder runs once per distinct frame)
        flow = raft_schedule.four_pair_flows(self.flow_predictor, fr0, fr1, iter)
        fr0, fr1 = x[:, :, 0], x[:, :, 1]
        flow = interpolate(flow, size=(H, W), mode="bilinear", align_corners=False) * (H / HH)
        flow = flow.reshape(4, B, 2, H, W)
        flow[0] *= 0.0
        flow[3] *= 0.0
        flow = flow.reshape(4 * B, 2, H, W)
        front_in = front(fr0.float(), fr1.float(), flow.float(), self.g_filter)  # Ours.py:562-578 + 613-637
    # encoder features (Ours.py:601-611)
    feat = self.encoder(torch.stack([fr0, fr1], 1), None)
    residual = feat[:, feat.shape[1] // 2].reshape(B, -1, H, W)
    feat = torch.cat((feat[:, feat.shape[1] // 2 - 1], feat[:, feat.shape[1] // 2 + 1]), 0)
    return feat, self.flow_process(front_in), residual, target_t, (HH, WW)


def forward_b200(self, x, input_target_frames, target_t, scale=None, rank=0, train_idx=0, use_GT=True, iter=12, flows=None):
    """Same signature and return value as ``LunaTokis.forward`` (``Ours.py:512, 858``)."""
    if self.training or use_GT:
        raise NotImplementedError("motif_b200 implements the inference path (eval(), use_GT=False), as VideoSRBaseModel.test() calls it")
    for flag, want in (("res_liff", False), ("siren", True), ("trans", False), ("warp_to_many", False)):
        if getattr(self, flag, want) != want:
            raise NotImplementedError(f"motif_b200 decodes the shipped configuration (setting 5); {flag}={getattr(self, flag)} is not supported")
    with torch.no_grad():
        feat, flow_feat, residual, tt, (HH, WW) = surround(self, x, target_t, scale, iter)
        # One decoder (weights + workspace) per device.  DataParallel replicas are shallow copies of the module, so they
        # share this dict and each finds (or builds, from ITS OWN parameters) the decoder of the device it runs on.
        cache = self.__dict__.setdefault("_motif_decoders", {})
        sd_version = sum(p._version for p in self.parameters())
        ens = bool(getattr(self, "local_ensemble", False))  # Ours.py:453
        key = (feat.device.type, feat.device.index)
        hit = cache.get(key)
        if hit is None or hit[1] != sd_version or hit[0].local_ensemble != ens:
            dec = SpaceTimeDecoder.from_state_dict(self.state_dict(), device=feat.device, local_ensemble=ens,
                                                   precision=getattr(self, "_motif_precision", "f16x3"))
            cache[key] = (dec, sd_version)
        dec = cache[key][0]
        rgb, flow_out = dec.decode(feat.float(), flow_feat.float(), residual.float(), tt, (HH, WW))
    return rgb, flow_out, 0.0  # Ours.py:580, 858: flow_GT = 0 on the inference path, returned as (0 / 20.0) / (HH / H)


def install(model, precision: str = "f16x3", raft_lookup: bool = True, dcn: bool = True):
    """Patch a reference ``LunaTokis`` instance in place and return it.  ``raft_lookup``: answer the reference's
    ``import alt_cuda_corr`` (``models/core/corr.py:5, 82`` -- a binary it does not ship) with ``motif_b200.alt_cuda_corr``,
    so that the shipped ``alternate_corr=True`` RAFT (``Ours.py:417-430``) runs without materialising the all-pairs volume.
    ``dcn``: rebind ``dcn_v2_conv`` of an imported reference ``DCNv2.dcn_v2`` module (the encoder's ``DCN_sep`` layers,
    ``Ours.py:53-172``) to the tcgen05 implicit GEMM of ``motif_b200.dcn_v2``."""
    if raft_lookup:
        alt_cuda_corr.install()
    if dcn:
        dcn_v2.install()
    object.__setattr__(model, "_motif_precision", precision)
    model.fwarp = Softsplat()
    model.fwarp_max = Softsplat_Max()
    model.fwarp_count = Softsplat_Count()
    # forward is replaced at CLASS level (a one-off subclass of the instance's own class): a bound method stored on the
    # instance would be copied into DataParallel replicas still bound to the device-0 module (VideoSR_base_model.py:36).
    cls = type(model)
    if not getattr(cls, "_motif_patched", False):
        model.__class__ = type(cls.__name__, (cls,), {"forward": forward_b200, "_motif_patched": True, "__module__": cls.__module__})
    return model


def test_b200(self, output=False):
    """Drop-in for ``VideoSRBaseModel.test`` (``models/VideoSR_base_model.py:169-197``) for the ``"Ours"`` networks.

    The reference decodes the clip's timestamps in chunks of three and re-runs the WHOLE forward -- RAFT on four HR frame
    pairs, the encoder, ``flow_process``, ``imnet`` -- for every chunk (``:188-193``), although none of that depends on the
    timestamps.  Here the surround runs once and all timestamps are decoded by one call (the decoder groups them by eight
    internally).  Attributes set as the reference sets them: ``fake_H [N,B,3,HH,WW]`` (the chunks were concatenated along
    dim 0, ``:193``), ``flow`` = the flow of the LAST chunk (``:194``; index ``(r*B + b) * n_chunk + n``), ``flow_GT``."""
    net = self.netG.module if hasattr(self.netG, "module") else self.netG
    if not ("Ours" in self.net_base and self.net_base != "Ours_44") or self.times is None or not hasattr(net, "_motif_precision"):
        return type(self)._motif_reference_test(self, output)
    self.netG.eval()
    with torch.no_grad():
        fake_H, flow, flow_GT = self.netG(self.var_L, self.real_H, self.times, self.scale, use_GT=False, iter=4)
        n_all = len(self.times)
        last0 = 3 * ((n_all - 1) // 3)  # first timestamp of the reference's last chunk
        two_b = flow.shape[0] // n_all
        self.fake_H = fake_H
        self.flow = flow.reshape(two_b, n_all, *flow.shape[1:])[:, last0:].reshape(two_b * (n_all - last0), *flow.shape[1:])
        self.flow_GT = flow_GT
    self.netG.train()
    if output == True:  # noqa: E712  (as the reference writes it)
        return self.fake_H


def install_test(model_wrapper, precision: str = "f16x3"):
    """``install`` on the wrapped ``LunaTokis`` plus the single-pass ``test`` above on a ``VideoSRBaseModel`` instance."""
    net = model_wrapper.netG.module if hasattr(model_wrapper.netG, "module") else model_wrapper.netG
    install(net, precision)
    cls = type(model_wrapper)
    if not hasattr(cls, "_motif_reference_test"):
        cls._motif_reference_test = cls.test
    model_wrapper.test = types.MethodType(test_b200, model_wrapper)
    return model_wrapper
