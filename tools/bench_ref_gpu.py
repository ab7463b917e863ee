"""The reference's own CUDA kernels (oracle/_ref/ref_gpu_sm100a.so, built by oracle/build_ref_gpu.py) timed on the
B200 beside the product kernels, same inputs: the forward splat at Adobe size and the PWC-Net cost volume."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from motif_b200 import _lib  # noqa: E402
from motif_b200.correlation import FunctionCorrelation  # noqa: E402
from motif_b200.softsplat_cp import FunctionSoftsplat  # noqa: E402
from oracle import build_ref_gpu as ref  # noqa: E402

assert ref.load() is not None, "oracle/_ref/ref_gpu_sm100a.so missing"


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


torch.manual_seed(0)
n, c, h, w = 1, 130, 720, 1280
x = torch.randn(n, c, h, w, device="cuda")
z = -torch.rand(n, 1, h, w, device="cuda")
low = torch.randn(n, 2, h // 64, w // 64, device="cuda") * 6
fl = torch.nn.functional.interpolate(low, size=(h, w), mode="bilinear", align_corners=False).contiguous()
alg = 1056 * h * w


def ref_softmax():
    e = z.exp()
    return ref.splat("sum", "adobe", torch.cat([x * e, e], 1), fl)


ref_in = torch.cat([x * z.exp(), z.exp()], 1).contiguous()
t_ref_kernel = timed(lambda: ref.splat("sum", "adobe", ref_in, fl))
t_ref_op = timed(ref_softmax)
t_new = timed(lambda: FunctionSoftsplat(x, fl, z, "softmax"))
print(f"softmax splat [1,130,720,1280]: reference kernel alone (incl. its zero-fill) {t_ref_kernel:.3f} ms ({alg / t_ref_kernel / 1e6:.0f} GB/s), "
      f"reference operator (exp, mul, cat + kernel) {t_ref_op:.3f} ms ({alg / t_ref_op / 1e6:.0f} GB/s), this repo {t_new:.3f} ms ({alg / t_new / 1e6:.0f} GB/s), "
      f"speed-up {t_ref_op / t_new:.1f}x")
for tag, b, cc, hh, ww in (("l2", 1, 32, 192, 320), ("l3", 1, 64, 96, 160), ("l6", 1, 196, 12, 20)):
    a = torch.randn(b, cc, hh, ww, device="cuda")
    bb = torch.randn(b, cc, hh, ww, device="cuda")
    t_r = timed(lambda: ref.correlation(tag, a, bb))
    t_n = timed(lambda: FunctionCorrelation(a, bb))
    # kernel alone (events recorded by the library around the launch): at these sizes the operator time above is mostly
    # Python / ctypes call overhead on both sides
    _lib.prof_enable(True)
    for _ in range(20):
        FunctionCorrelation(a, bb)
    pk = _lib.prof_collect(["corr_kernel"])["corr_kernel"]
    _lib.prof_enable(False)
    k_ms = pk[0] / max(pk[1], 1)
    flops, nbytes = 2 * 81 * cc * b * hh * ww, 4 * (2 * b * cc * hh * ww + 81 * b * hh * ww)
    print(f"correlation [{b},{cc},{hh},{ww}]: reference (2 rearranges + kernel) {t_r:.3f} ms, this repo {t_n:.3f} ms, speed-up {t_r / t_n:.1f}x; "
          f"kernel alone {k_ms * 1e3:.1f} us = {flops / k_ms / 1e9:.2f} TFLOP/s fp32, {nbytes / k_ms / 1e6:.0f} GB/s algorithmic")
