"""DCNv2 forward at the Adobe LR size (64 -> 64 channels, 8 deformable groups, 180x320): the fused kernel beside
torchvision's deform_conv2d on the same B200 (im2col + GEMM like the reference's extension, which does not build here)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torchvision  # noqa: E402

from motif_b200 import _lib  # noqa: E402
from motif_b200.dcn_v2 import dcn_v2_conv  # noqa: E402

torch.manual_seed(0)
B, C, H, W, dg = 1, 64, 180, 320, 8
x = torch.randn(B, C, H, W, device="cuda")
off = torch.randn(B, dg * 18, H, W, device="cuda") * 2
m = torch.rand(B, dg * 9, H, W, device="cuda")
w = torch.randn(C, C, 3, 3, device="cuda") / 24
b = torch.randn(C, device="cuda")


def timed(fn, reps=30):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


with torch.no_grad():
    a, r = dcn_v2_conv(x, off, m, w, b, 1, 1, 1, dg), torchvision.ops.deform_conv2d(x, off, w, b, 1, 1, 1, m)
    print("max|new - torchvision| = %.2e" % float((a - r).abs().max()))
    t_new = timed(lambda: dcn_v2_conv(x, off, m, w, b, 1, 1, 1, dg))
    t_tv = timed(lambda: torchvision.ops.deform_conv2d(x, off, w, b, 1, 1, 1, m))
flops = 2 * C * C * 9 * H * W * B
print("DCNv2 64->64, 8 groups, 180x320: this repo %.3f ms (%.1f TFLOP/s fp32), torchvision %.3f ms (%.2fx)" % (t_new, flops / t_new / 1e9, t_tv, t_tv / t_new))
