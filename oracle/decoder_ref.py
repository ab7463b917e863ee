"""CPU restatement of the space-time local implicit decoder (TEST INFRASTRUCTURE ONLY).

Restates ``models/modules/Ours.py:659-858`` (``LunaTokis.forward`` from the
coordinate generation to the clamp), ``make_coord`` ``Ours.py:874-889`` and
``models/modules/SIREN.py:44-45, 76-79`` with the same torch fp32 operators in
the same order, for the shipped configuration (``test.yml:50`` ``setting: 5``:
``warp_to_many=False``, ``decoder_Z=predict_Z=True``, ``siren=True``,
``res_liff=False``, ``groups=1``, ``use_GT=False``).  ``local_ensemble`` is a flag
(reference default ``False``, ``Ours.py:453``).

Stage names follow SURVEY.md section 3.2 (f..n).  Every intermediate the CUDA
path is compared against is returned in a dict.

Pinned by ``tests/test_oracle_pins.py``: the captured hot-path inputs/outputs of
the unmodified reference forward (``oracle/make_golden.py``, committed under
``tests/golden/decoder_*.npz``) must be reproduced bit-for-bit on the CPU.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn.functional as F

from . import softsplat_ref as S

__all__ = ["make_coord", "siren", "query_geometry", "decode", "SIREN_SPECS", "random_params", "count_unstable_mask", "REALISTIC"]

# (in_features, hidden widths, out_features) -- Ours.py:470-471, 487-491
SIREN_SPECS = {
    "flow_imnet": (67, [64, 64, 256], 3),
    "imnet": (66, [64, 64, 256], 64),
    "synth_net": (198, [64, 64, 64, 256], 3),
}


def make_coord(shape, flatten=True):
    """Pixel-centre coordinates in [-1, 1] (``Ours.py:874-889``), built in fp32 on the CPU."""
    seqs = []
    for n in shape:
        v0, v1 = -1, 1
        r = (v1 - v0) / (2 * n)
        seqs.append(v0 + r + (2 * r) * torch.arange(n).float())
    ret = torch.stack(torch.meshgrid(*seqs, indexing="ij"), dim=-1)
    if flatten:
        ret = ret.view(-1, ret.shape[-1])
    return ret


def coord_sequence(n: int) -> torch.Tensor:
    """The 1-D sequence ``make_coord`` builds for an axis of length ``n``."""
    r = (1 - (-1)) / (2 * n)
    return -1 + r + (2 * r) * torch.arange(n).float()


def siren(x: torch.Tensor, params: Dict[str, torch.Tensor], name: str) -> torch.Tensor:
    """``Siren.forward`` (``SIREN.py:49-79``): sine layers ``sin(30*(xW^T+b))`` then a plain linear."""
    n_sine = len(SIREN_SPECS[name][1])
    for i in range(n_sine):
        w = params[f"{name}.net.{i}.linear.weight"]
        b = params[f"{name}.net.{i}.linear.bias"]
        x = torch.sin(30 * F.linear(x, w, b))  # SIREN.py:45
    return F.linear(x, params[f"{name}.net.{n_sine}.weight"], params[f"{name}.net.{n_sine}.bias"])


def query_geometry(H: int, W: int, HH: int, WW: int, vx: int = 0, vy: int = 0):
    """Steps f, g(index), h(rel) of SURVEY 3.2: ``Ours.py:667-689, 704, 720-722``.

    Returns ``hr_coord [qs,2]`` (y,x), ``coord_ [qs,2]`` (shifted/clamped),
    ``iy, ix [qs] int64`` (ATen nearest index), ``q_coord [qs,2]``, ``rel [qs,2]``.
    The index is taken from ATen ``grid_sample`` itself by sampling an index image.
    """
    hr_coord = make_coord((HH, WW)).unsqueeze(0)
    rx = 2 / H / 2
    ry = 2 / W / 2
    eps_shift = 1e-6  # Ours.py:669 (reassigned after the ensemble switch)
    coord_ = hr_coord.clone()
    coord_[:, :, 0] += vx * rx + eps_shift
    coord_[:, :, 1] += vy * ry + eps_shift
    coord_.clamp_(-1 + 1e-6, 1 - 1e-6)
    feat_coord = make_coord((H, W), flatten=False).permute(2, 0, 1).unsqueeze(0).expand(1, 2, H, W)
    idx_img = torch.stack(
        [
            torch.arange(H, dtype=torch.float32).view(H, 1).expand(H, W),
            torch.arange(W, dtype=torch.float32).view(1, W).expand(H, W),
        ]
    ).unsqueeze(0)
    grid = coord_.flip(-1).unsqueeze(1)
    got = F.grid_sample(torch.cat([idx_img, feat_coord], 1), grid, mode="nearest", align_corners=False)[:, :, 0, :]
    iy = got[0, 0].to(torch.int64)
    ix = got[0, 1].to(torch.int64)
    q_coord = got[:, 2:4].permute(0, 2, 1)  # [1,qs,2]
    rel = hr_coord - q_coord
    rel[:, :, 0] *= H
    rel[:, :, 1] *= W
    return {
        "hr_coord": hr_coord[0],
        "coord_": coord_[0],
        "iy": iy,
        "ix": ix,
        "q_coord": q_coord[0],
        "rel": rel[0],
    }


def decode(
    feat: torch.Tensor,  # [2B,64,H,W]   F_0^L, F_1^L   (Ours.py:609)
    flow_feat: torch.Tensor,  # [2B,64,H,W]   T_0^L, T_1^L   (Ours.py:638)
    residual: torch.Tensor,  # [B,64,H,W]    F_01^L         (Ours.py:607)
    target_t: torch.Tensor,  # [B,N]
    HH: int,
    WW: int,
    params: Dict[str, torch.Tensor],
    local_ensemble: bool = False,
    return_intermediates: bool = False,
    splat_ops=None,
):
    """``Ours.py:659-858`` on the CPU.  Returns ``(rgb [N,B,3,HH,WW], flow_out [2BN,2,HH,WW])``.

    The same torch operators run on CUDA tensors when the inputs live there (``oracle/decoder_ref_gpu.py``: the
    reference's eager GPU path, with ``splat_ops`` = the reference's own CUDA kernels); ``splat_ops`` defaults to the
    CPU restatement of the three splats (``oracle/softsplat_ref.py``)."""
    S_ = splat_ops if splat_ops is not None else S
    feat = feat.float()
    flow_feat = flow_feat.float()
    residual = residual.float()
    dev = feat.device
    target_t = target_t.float().to(dev)
    B, N = target_t.shape
    H, W = feat.shape[-2:]
    alpha = params["alpha"].float().to(dev)

    if local_ensemble:
        vx_lst, vy_lst = [-1, 1], [-1, 1]
    else:
        vx_lst, vy_lst = [0], [0]

    hr_coord = make_coord((HH, WW)).unsqueeze(0).to(dev)  # built on the CPU, then moved (Ours.py:667-668, 677)
    eps_shift = 1e-6
    rx = 2 / H / 2
    ry = 2 / W / 2
    feat_coord = make_coord((H, W), flatten=False).permute(2, 0, 1).unsqueeze(0).expand(1, 2, H, W).to(dev)

    preds, areas = [], []
    inter = {}
    for vx in vx_lst:
        for vy in vy_lst:
            coord_ = hr_coord.clone()
            coord_[:, :, 0] += vx * rx + eps_shift
            coord_[:, :, 1] += vy * ry + eps_shift
            coord_.clamp_(-1 + 1e-6, 1 - 1e-6)

            c1, c3, c4, c5 = 2 * B * feat.shape[1], 2 * B * flow_feat.shape[1], 2, residual.shape[1] * B
            to_be_warp = torch.cat(
                (feat.reshape(1, c1, H, W), flow_feat.reshape(1, c3, H, W), feat_coord.reshape(1, c4, H, W), residual.reshape(1, c5, H, W)), 1
            )
            warped = F.grid_sample(to_be_warp, coord_.flip(-1).unsqueeze(1), mode="nearest", align_corners=False)[:, :, 0, :]
            q_feat, warped = warped[:, :c1], warped[:, c1:]
            q_flow_feat, warped = warped[:, :c3], warped[:, c3:]
            q_coord, warped = warped[:, :c4], warped[:, c4:]
            q_residual = warped[:, :c5]
            q_feat = q_feat.reshape(2 * B, -1, HH * WW).permute(0, 2, 1)
            q_flow_feat = q_flow_feat.reshape(2 * B, -1, HH * WW).permute(0, 2, 1)
            q_coord = q_coord.reshape(1, -1, HH * WW).permute(0, 2, 1)
            q_residual = q_residual.reshape(B, -1, HH * WW).permute(0, 2, 1)

            rel_coord = hr_coord - q_coord
            rel_coord[:, :, 0] *= H
            rel_coord[:, :, 1] *= W
            qs = rel_coord.shape[1]
            q_feat_low = q_feat.clone()

            flow_in = torch.cat(
                [
                    q_flow_feat.repeat(1, N, 1).reshape(2 * B * N, qs, -1),
                    target_t.reshape(B * N, 1, 1).repeat(2, qs, 1),
                    rel_coord.repeat(2 * B * N, 1, 1),
                ],
                dim=-1,
            )
            im_in = torch.cat([q_feat.reshape(2 * B, qs, -1), rel_coord.repeat(2 * B, 1, 1)], dim=-1)
            q_flow_out = siren(flow_in, params, "flow_imnet")
            q_feat_out = siren(im_in, params, "imnet")
            preds.append([q_feat_out, q_feat_low, q_residual, q_flow_out])
            area = torch.abs(rel_coord[:, :, 0] * rel_coord[:, :, 1])
            areas.append(area + 1e-9)
            if return_intermediates and not inter:
                inter["rel"] = rel_coord[0].clone()
                inter["coord_"] = coord_[0].clone()

    tot_area = torch.stack(areas).sum(dim=0)
    if local_ensemble:
        areas[0], areas[3] = areas[3], areas[0]
        areas[1], areas[2] = areas[2], areas[1]
    ret = [0, 0, 0, 0]
    for pred, area in zip(preds, areas):
        for i, p in enumerate(pred):
            ret[i] = ret[i] + p * (area / tot_area).unsqueeze(-1)
    q_feat, q_feat_low, q_residual, q_flow = ret

    featm = q_feat.reshape(2 * B, HH, WW, -1).permute(0, 3, 1, 2)
    feat_low = q_feat_low.reshape(2 * B, HH, WW, -1).permute(0, 3, 1, 2)
    q_residual = q_residual.reshape(B, HH, WW, -1).permute(0, 3, 1, 2)
    flow = q_flow.reshape(2 * B * N, HH, WW, -1).permute(0, 3, 1, 2).reshape(2 * B * N, -1, HH, WW)
    splat_in = torch.cat(
        [
            featm.reshape(2 * B, -1, HH, WW).repeat(1, N, 1, 1).reshape(2 * B * N, -1, HH, WW),
            flow[:, :-1],
            feat_low.reshape(2 * B, -1, HH, WW).repeat(1, N, 1, 1).reshape(2 * B * N, -1, HH, WW),
        ],
        1,
    )
    raw_flow = flow
    flow, z = flow[:, :-1] * 20.0 * (HH / H), (torch.relu(flow[:, -1].unsqueeze(1)) * alpha)

    output, warped_z = S_.function_softsplat(splat_in, flow, z, "softmax")
    z_max = S_.function_softsplat_max(z.exp(), flow)
    count = S_.function_softsplat_count(z, flow)
    output = output.clone()
    warped_z = warped_z.clone()
    if return_intermediates:
        inter.update(
            imnet_out=featm, raw_flow=raw_flow, flow_hr=flow, z=z, splat_sum=output.clone(), splat_norm=warped_z.clone(),
            splat_max=z_max.clone(), splat_count=count.clone(),
        )

    output = output.reshape(2, B * N, -1, HH, WW).sum(0)
    warped_z = warped_z.reshape(2, B * N, -1, HH, WW).sum(0)
    warped_z[warped_z == 0] = 1.0
    output /= warped_z
    z_max = z_max.reshape(2, B * N, -1, HH, WW).max(0)[0]
    count = count.reshape(2, B * N, -1, HH, WW).sum(0)

    count_ = count.clone()
    count_[count_ == 0.0] = 1.0
    warped_z_ = warped_z.clone()
    warped_z_[warped_z_ == 1.0] = 0.0
    extra = torch.cat((z_max, count / 16.0, (warped_z_ / count_)), 1)

    output_all = torch.cat(
        (
            output.reshape(B * N, -1, HH, WW),
            extra.reshape(B * N, -1, HH, WW),
            q_residual.repeat(1, N, 1, 1).reshape(B * N, -1, HH, WW),
            target_t.reshape(B * N, 1, 1, 1).repeat(1, 1, HH, WW),
        ),
        1,
    ).reshape(B * N, -1, HH, WW)
    if return_intermediates:
        inter["synth_in"] = output_all.clone()
    rgb = siren(output_all.reshape(B * N, -1, HH * WW).permute(0, 2, 1), params, "synth_net")
    rgb = rgb.permute(0, 2, 1).reshape(B, N, -1, HH, WW).permute(1, 0, 2, 3, 4)
    rgb = torch.clamp(rgb, 0, 1)
    flow_out = flow / 20.0 / (HH / H)
    if return_intermediates:
        return rgb, flow_out, inter
    return rgb, flow_out


def random_params(seed: int = 0, weight_gain: float = 1.0, alpha: float = -20.0, first_gain: Optional[float] = None,
                  rgb_bias: Optional[float] = None, rgb_gain: float = 1.0, z_bias: Optional[float] = None):
    """Seeded synthetic hot-path weights in the ``best.pth`` key layout.

    ``weight_gain=1`` reproduces the reference initialisation (``SIREN.py:35-42, 63-67``
    plus ``nn.Linear`` default bias init); larger gains give the O(1)-scaled variant of
    SURVEY 8c so that sine arguments and ``exp(z)`` are not degenerate; ``rgb_bias`` /
    ``rgb_gain`` move the synthetic RGB off the clamp at 0.
    """
    import math

    g = torch.Generator().manual_seed(seed)
    p = {}
    for name, (fin, hidden, fout) in SIREN_SPECS.items():
        widths = [fin] + hidden
        for i in range(len(hidden)):
            k_in, k_out = widths[i], widths[i + 1]
            if i == 0:
                bound = (1.0 / k_in) * (first_gain if first_gain is not None else weight_gain)
            else:
                bound = math.sqrt(6.0 / k_in) / 30.0 * weight_gain
            p[f"{name}.net.{i}.linear.weight"] = (torch.rand(k_out, k_in, generator=g) * 2 - 1) * bound
            p[f"{name}.net.{i}.linear.bias"] = (torch.rand(k_out, generator=g) * 2 - 1) / math.sqrt(k_in)
        k_in = hidden[-1]
        bound = math.sqrt(6.0 / k_in) / 30.0 * weight_gain
        i = len(hidden)
        p[f"{name}.net.{i}.weight"] = (torch.rand(fout, k_in, generator=g) * 2 - 1) * bound
        p[f"{name}.net.{i}.bias"] = (torch.rand(fout, generator=g) * 2 - 1) / math.sqrt(k_in)
    if rgb_bias is not None:  # centre the synthetic RGB inside the clamp range
        p["synth_net.net.4.bias"] = torch.tensor([rgb_bias - 0.1, rgb_bias, rgb_bias + 0.1])
    p["synth_net.net.4.weight"] = p["synth_net.net.4.weight"] * rgb_gain
    if z_bias is not None:  # make relu(z_raw) switch on and off across the image
        p["flow_imnet.net.3.bias"][2] = z_bias
    p["alpha"] = torch.ones(1) * alpha
    return p


# Non-degenerate but well-conditioned synthetic weights used by the parity tests: reference SIREN scale for
# the hidden layers, first layers x4, RGB centred in the clamp range, z_raw switching sign, alpha as
# initialised by the reference (Ours.py:509).
REALISTIC = dict(weight_gain=1.0, first_gain=4.0, alpha=-20.0, rgb_bias=0.5, rgb_gain=3.0, z_bias=0.03)


def count_unstable_mask(flow_hr: torch.Tensor, B: int, N: int, eps: float = 2.5e-4, splat_ops=None) -> torch.Tensor:
    """Destination pixels at which the REFERENCE FUNCTION ITSELF is discontinuous in the flow.

    The count splat (``softsplat_count_cp.py:25-50``) adds 1 to the four corners of
    ``floor(position)``: when a source lands within ``eps`` pixels of an integer coordinate, an
    arbitrarily small change of the flow (fp32 re-association inside ``flow_imnet``, TF32x3 vs fp32,
    CUDA ``sinf`` vs the host's) moves a whole unit of ``count`` between neighbouring destinations, and
    ``count/16`` and ``wz/count`` are ``synth_net`` inputs (``Ours.py:834``).  Two correct evaluations of
    the reference differ by O(1e-2) in RGB at such pixels, so the 1e-3 gate is applied to the others.
    ``flow_hr`` is ``[2*B*N, 2, HH, WW]`` (reference-major); returns bool ``[N, B, 1, HH, WW]``.
    """
    S_ = splat_ops if splat_ops is not None else S
    base = S_.function_softsplat_count(flow_hr[:, :1], flow_hr)
    bad = torch.zeros_like(base, dtype=torch.bool)
    for sx in (-eps, eps):
        for sy in (-eps, eps):
            f = flow_hr.clone()
            f[:, 0] += sx
            f[:, 1] += sy
            bad |= S_.function_softsplat_count(f[:, :1], f) != base
    # The joint shifts above can cancel (one source leaves a destination while another enters it).  Per source: a source
    # that lands within eps of an integer grid line can move its unit of count between the destinations of the 3 x 3
    # neighbourhood of its rounded landing position, whatever the other sources do.
    n, _, HH, WW = flow_hr.shape
    dev = flow_hr.device
    px = torch.arange(WW, dtype=torch.float32, device=dev).view(1, 1, WW) + flow_hr[:, 0]
    py = torch.arange(HH, dtype=torch.float32, device=dev).view(1, HH, 1) + flow_hr[:, 1]
    rx, ry = torch.round(px), torch.round(py)
    near = ((px - rx).abs() < eps) | ((py - ry).abs() < eps)
    near &= (rx >= -1) & (rx <= WW) & (ry >= -1) & (ry <= HH)
    idx = (ry.clamp(0, HH - 1) * WW + rx.clamp(0, WW - 1)).long()
    hit = torch.zeros(n, HH * WW, dtype=torch.float32, device=dev)
    hit.scatter_add_(1, idx.view(n, -1), near.view(n, -1).float())
    bad |= F.max_pool2d(hit.view(n, 1, HH, WW), 3, stride=1, padding=1) > 0
    return bad.reshape(2, B, N, 1, HH, WW).any(0).permute(1, 0, 2, 3, 4)
