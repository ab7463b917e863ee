"""RAFT-small correlation lookup at the HR size of the Adobe workload (720x1280 / 8 = 90x160, C = 128, r = 3, 4 levels):
motif_b200.alt_cuda_corr (one launch per level, nothing materialised) beside the reference's in-repo CorrBlock
(models/core/corr.py:8-56 as eager torch ops on the same GPU: all-pairs volume built once, four grid_samples per iteration)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from motif_b200 import _lib, alt_cuda_corr  # noqa: E402

B, C, H, W, r, L = 1, 128, 90, 160, 3, 4
dev = torch.device("cuda:0")
torch.manual_seed(0)
fmap1, fmap2 = torch.randn(B, C, H, W, device=dev), torch.randn(B, C, H, W, device=dev)
ys, xs = torch.meshgrid(torch.arange(H, device=dev).float(), torch.arange(W, device=dev).float(), indexing="ij")
coords = torch.stack([xs, ys], 0)[None] + torch.randn(B, 2, H, W, device=dev) * 2


def sampler(img, c):
    h, w = img.shape[-2:]
    x, y = c.split([1, 1], dim=-1)
    return F.grid_sample(img, torch.cat([2 * x / (w - 1) - 1, 2 * y / (h - 1) - 1], -1), align_corners=True)


def build():
    corr = torch.matmul(fmap1.view(B, C, H * W).transpose(1, 2), fmap2.view(B, C, H * W)).view(B * H * W, 1, H, W) / C ** 0.5
    pyr = [corr]
    for _ in range(L - 1):
        pyr.append(F.avg_pool2d(pyr[-1], 2, stride=2))
    return pyr


def lookup_ref(pyr):
    c = coords.permute(0, 2, 3, 1)
    d = torch.linspace(-r, r, 2 * r + 1, device=dev)
    delta = torch.stack(torch.meshgrid(d, d, indexing="ij"), -1).view(1, 2 * r + 1, 2 * r + 1, 2)
    out = [sampler(pyr[i], c.reshape(B * H * W, 1, 1, 2) / 2 ** i + delta).view(B, H, W, -1) for i in range(L)]
    return torch.cat(out, -1).permute(0, 3, 1, 2).contiguous()


f1n = fmap1.permute(0, 2, 3, 1).contiguous()
f2n = [fmap2]
for _ in range(L - 1):
    f2n.append(F.avg_pool2d(f2n[-1], 2, stride=2))
f2n = [f.permute(0, 2, 3, 1).contiguous() for f in f2n]


def lookup_new():
    c = coords.permute(0, 2, 3, 1)
    out = [alt_cuda_corr.forward(f1n, f2n[i], (c / 2 ** i).reshape(B, 1, H, W, 2).contiguous(), r)[0].squeeze(1) for i in range(L)]
    return torch.stack(out, 1).reshape(B, -1, H, W) / C ** 0.5


def lookup_pyramid():
    return alt_cuda_corr.forward_pyramid(f1n, f2n, coords.permute(0, 2, 3, 1).contiguous(), r, normalize=True)


def timed(fn, reps=20):
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


pyr = build()
print("max|new - CorrBlock (GPU eager)| = %.2e (values up to %.1f)" % (float((lookup_new() - lookup_ref(pyr)).abs().max()), float(lookup_ref(pyr).abs().max())))
t_build, t_ref, t_new = timed(build, 5), timed(lambda: lookup_ref(pyr)), timed(lookup_new)
_lib.prof_enable(True)
for _ in range(10):
    lookup_new()
p = _lib.prof_collect(["raft_corr_lookup_kernel"])["raft_corr_lookup_kernel"]
_lib.prof_enable(False)
t_pyr = timed(lookup_pyramid)
print("one-launch pyramid lookup (motif_raft_corr_lookup_pyramid): %.3f ms per RAFT iteration, bit-equal to the level-by-level path: %s"
      % (t_pyr, bool(torch.equal(lookup_pyramid(), lookup_new()))))
flops = sum(2 * (2 * r + 2) ** 2 * C * B * H * W for _ in range(L))
print("lookup, one RAFT iteration (4 levels, 90x160, C=128, r=3): this repo %.3f ms (kernels %.3f ms, %.2f TFLOP/s fp32), CorrBlock lookup %.3f ms + "
      "%.2f ms once per pair to build its %.0f MB volume pyramid; 12 iterations: %.2f ms vs %.2f ms"
      % (t_new, p[0] / 10, flops / (p[0] / 10) / 1e9, t_ref, t_build, sum(x.numel() for x in pyr) * 4 / 1e6, 12 * t_new, t_build + 12 * t_ref))
