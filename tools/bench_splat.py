"""Stand-alone softmax splat operator at BASELINE size: per-kernel CUDA-event times and algorithmic GB/s."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from motif_b200 import _lib
from motif_b200.softsplat_cp import FunctionSoftsplat, _splat

HH, WW, C = 720, 1280, 130
dev = torch.device("cuda:0")
torch.manual_seed(0)
x = torch.randn(1, C, HH, WW, device=dev)
z = -torch.rand(1, 1, HH, WW, device=dev)
for name, cell, mag in (("smooth flow (1/64-res noise x 6 px)", 64, 6.0), ("distorting flow (1/16-res noise x 6 px, 6% of destinations > 8 contributions)", 16, 6.0)):
    low = torch.randn(1, 2, HH // cell, WW // cell, device=dev) * mag
    fl = torch.nn.functional.interpolate(low, size=(HH, WW), mode="bilinear", align_corners=False).contiguous()
    for _ in range(3):
        FunctionSoftsplat(x, fl, z, "softmax")
    torch.cuda.synchronize()
    _lib.prof_enable(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    reps = 10
    for _ in range(reps):
        FunctionSoftsplat(x, fl, z, "softmax")
    e1.record()
    torch.cuda.synchronize()
    p = _lib.prof_collect(["splat_bin_kernel", "splat_gather_kernel", "splat_scatter_kernel"])
    _lib.prof_enable(False)
    alg = 1056 * HH * WW
    op_ms = e0.elapsed_time(e1) / reps
    print(name, "operator %.3f ms (%.0f GB/s)" % (op_ms, alg / op_ms / 1e6), {k: round(v[0] / max(v[1], 1), 4) for k, v in p.items()},
          "gather-only %.0f GB/s" % (alg / (p["splat_gather_kernel"][0] / reps) / 1e6))
    a = _splat(x, fl, z, 3, atomic=True)
    o, n = FunctionSoftsplat(x, fl, z, "softmax")
    print("   max|gather - atomic| =", float((torch.cat([o, n], 1) - a).abs().max()))
    torch.cuda.synchronize()
    e0.record()
    for _ in range(3):
        _splat(x, fl, z, 3, atomic=True)
    e1.record()
    torch.cuda.synchronize()
    print("   atomic scatter variant: %.3f ms (%.0f GB/s)" % (e0.elapsed_time(e1) / 3, alg / (e0.elapsed_time(e1) / 3) / 1e6))
