"""Gather-kernel time of the softmax splat against the channel count (fixed per-tile cost vs per-channel cost)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from motif_b200 import _lib
from motif_b200.softsplat_cp import FunctionSoftsplat

HH, WW = 720, 1280
dev = torch.device("cuda:0")
torch.manual_seed(0)
low = torch.randn(1, 2, HH // 64, WW // 64, device=dev) * 6
fl = torch.nn.functional.interpolate(low, size=(HH, WW), mode="bilinear", align_corners=False).contiguous()
z = -torch.rand(1, 1, HH, WW, device=dev)
for C in (2, 10, 34, 66, 130, 258):
    x = torch.randn(1, C, HH, WW, device=dev)
    for _ in range(3):
        FunctionSoftsplat(x, fl, z, "softmax")
    torch.cuda.synchronize()
    _lib.prof_enable(True)
    for _ in range(10):
        FunctionSoftsplat(x, fl, z, "softmax")
    p = _lib.prof_collect(["splat_bin_kernel", "splat_gather_kernel", "splat_scatter_kernel"])
    _lib.prof_enable(False)
    g = p["splat_gather_kernel"][0] / 10
    print("C=%3d gather %.4f ms  (%.0f GB/s algorithmic)  bin %.4f" % (C, g, 4 * (2 * C + 4) * HH * WW / g / 1e6, p["splat_bin_kernel"][0] / 10))
