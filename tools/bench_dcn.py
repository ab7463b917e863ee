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


flops = 2 * C * C * 9 * H * W * B
# the offsets of the model are convolution outputs (spatially smooth); independent noise per pixel is the worst case for
# the sampling (every lane of a warp lands on another cache line)
off_smooth = torch.nn.functional.interpolate(torch.randn(B, dg * 18, H // 8 + 1, W // 8 + 1, device="cuda") * 4, size=(H, W), mode="bilinear", align_corners=False)
with torch.no_grad():
    for name, o in (("independent offsets per pixel (sigma 2 px)", off), ("smooth offsets (1/8-resolution noise x 4 px)", off_smooth.contiguous())):
        a, r = dcn_v2_conv(x, o, m, w, b, 1, 1, 1, dg), torchvision.ops.deform_conv2d(x, o, w, b, 1, 1, 1, m)
        err = float((a - r).abs().max())
        t_new = timed(lambda: dcn_v2_conv(x, o, m, w, b, 1, 1, 1, dg))
        t_tv = timed(lambda: torchvision.ops.deform_conv2d(x, o, w, b, 1, 1, 1, m))
        print("DCNv2 64->64, 8 groups, 180x320, %s: this repo %.3f ms (%.1f TFLOP/s fp32-equivalent), torchvision %.3f ms (%.2fx); max|new - torchvision| = %.2e (values up to %.1f)"
              % (name, t_new, flops / t_new / 1e9, t_tv, t_tv / t_new, err, float(r.abs().max())))
