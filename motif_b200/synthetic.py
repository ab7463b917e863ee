"""Seeded synthetic inputs for benchmarks and smoke runs (no datasets or checkpoints are reachable).

``synthetic_params`` draws the three hot-path SIREN MLPs in the ``best.pth`` key layout with the
reference's initialisation rule (``SIREN.py:35-42, 63-67``): hidden layers at the reference scale,
first layers x4, RGB centred in the clamp range and ``z_raw`` switching sign, so that sine arguments,
flows of a few HR pixels and a non-degenerate ``exp(z)`` are exercised while the fp32 problem stays
well-conditioned.  ``synthetic_latents`` draws
smooth LR latents of the encoder's output shapes (``Ours.py:601-638``).
"""
from __future__ import annotations

import math

import torch

SPECS = {
    "flow_imnet": (67, [64, 64, 256], 3),
    "imnet": (66, [64, 64, 256], 64),
    "synth_net": (198, [64, 64, 64, 256], 3),
}

# (name, LR H, LR W, HR HH, HR WW, timestamps) -- BASELINE.json configs
WORKLOADS = {
    "vimeo_x4": (64, 112, 256, 448, [0.5]),
    "adobe240_x4_t8": (180, 320, 720, 1280, [k / 8 for k in range(1, 8)]),
    "adobe240_x3p5_t12": (180, 320, 630, 1120, [k / 12 for k in range(1, 12)]),
    "uhd4k_x4_t8": (540, 960, 2160, 3840, [k / 8 for k in range(1, 8)]),
}


def synthetic_params(seed=0, weight_gain=1.0, first_gain=4.0, alpha=-20.0, rgb_bias=0.5, rgb_gain=3.0, z_bias=0.03):
    g = torch.Generator().manual_seed(seed)
    p = {}
    for name, (fin, hidden, fout) in SPECS.items():
        widths = [fin] + hidden + [fout]
        for i in range(len(widths) - 1):
            k_in, k_out = widths[i], widths[i + 1]
            last = i == len(widths) - 2
            bound = (first_gain / k_in) if i == 0 else weight_gain * math.sqrt(6.0 / k_in) / 30.0
            key = f"{name}.net.{i}." + ("" if last else "linear.")
            p[key + "weight"] = (torch.rand(k_out, k_in, generator=g) * 2 - 1) * bound
            p[key + "bias"] = (torch.rand(k_out, generator=g) * 2 - 1) / math.sqrt(k_in)
    p["synth_net.net.4.weight"] *= rgb_gain
    p["synth_net.net.4.bias"] = torch.tensor([rgb_bias - 0.1, rgb_bias, rgb_bias + 0.1])
    p["flow_imnet.net.3.bias"][2] = z_bias
    p["alpha"] = torch.ones(1) * alpha
    return p


def synthetic_latents(B, H, W, seed=0, scale=0.4):
    """(feat [2B,64,H,W], flow_feat [2B,64,H,W], residual [B,64,H,W]) smooth fp32 CPU tensors."""
    g = torch.Generator().manual_seed(seed)
    low = torch.randn(5 * B, 64, max(H // 4, 2), max(W // 4, 2), generator=g)
    lat = torch.nn.functional.interpolate(low, size=(H, W), mode="bilinear", align_corners=False) * scale
    return lat[: 2 * B].contiguous(), lat[2 * B: 4 * B].contiguous(), lat[4 * B:].contiguous()
