"""Reliability-map front end at Adobe240 LR size on the B200: the fused kernel beside the reference's own op sequence
(Ours.py:562-578, 613-637 as eager torch CUDA ops -- the reference's code path for this step) and the CPU oracle."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from motif_b200 import _lib  # noqa: E402
from motif_b200.flow_front import flow_front  # noqa: E402
from oracle import flow_front_ref  # noqa: E402  (tools/ is measurement infrastructure, not the product path)

B, H, W = 1, 180, 320
dev = torch.device("cuda:0")
gen = torch.Generator().manual_seed(0)
fr0, fr1 = torch.rand(B, 3, H, W, generator=gen), torch.rand(B, 3, H, W, generator=gen)
low = torch.randn(4 * B, 2, H // 16, W // 16, generator=gen) * 3
flow = torch.nn.functional.interpolate(low, size=(H, W), mode="bilinear", align_corners=False).contiguous()
flow[:B] = 0
flow[3 * B:] = 0
gf = torch.tensor([[1, 2, 1], [2, 4, 2], [1, 2, 1]], dtype=torch.float32) / 16.0
d = [t.to(dev) for t in (fr0, fr1, flow, gf)]


def eager(fr0, fr1, flow, gf):
    """the oracle's op sequence is the reference's own (same torch calls), run here on the GPU tensors"""
    dur = torch.tensor([[0, 0], [0, 8], [8, 0], [8, 8]], dtype=torch.float32, device=flow.device).unsqueeze(1)
    F = torch.nn.functional
    bw = lambda img, fl: F.grid_sample(img, torch.stack((((torch.arange(W, device=dev).view(1, 1, W) + fl[:, 0]) / W) * 2 - 1,  # noqa: E731
                                                        ((torch.arange(H, device=dev).view(1, H, 1) + fl[:, 1]) / H) * 2 - 1), -1),
                                       mode="bilinear", align_corners=True, padding_mode="border")
    warped = bw(torch.cat([fr0, fr1, fr0, fr1], 0), flow)
    psi_photo = F.l1_loss(torch.cat([fr0, fr0, fr1, fr1], 0), warped, reduction="none").mean(1)
    f4 = flow.reshape(4, B, 2, H, W)
    warped = bw(-torch.cat([f4[0], f4[2], f4[1], f4[3]], 0), flow)
    psi_flow = F.l1_loss(flow, warped, reduction="none").mean(1)
    sq, mn = torch.split(F.conv3d(F.pad(torch.cat([flow ** 2, flow], 1), (1, 1, 1, 1), mode="reflect").unsqueeze(1), gf.reshape(1, 1, 1, 3, 3)).squeeze(1), 2, dim=1)
    psi_var = (sq - mn ** 2).clip(1e-9, None).sqrt().mean(1)
    psies = torch.stack([psi_photo, psi_flow / 10.0, psi_var], 1)
    return torch.cat(((flow / 20.0).reshape(2, 2, B, -1, H, W).permute(0, 2, 1, 3, 4, 5).reshape(2 * B, 2, -1, H, W),
                      psies.reshape(2, 2, B, -1, H, W).permute(0, 2, 1, 3, 4, 5).reshape(2 * B, 2, -1, H, W),
                      dur.reshape(2, 4, 1, 1).unsqueeze(1).repeat(1, B, 1, H, W).reshape(2 * B, 2, 2, H, W) / 8.0), 2).reshape(2 * B, -1, H, W)


def timed(fn, reps=50):
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


out = flow_front(*d)
ref = eager(*d)
print("max|kernel - eager torch (GPU)| = %.2e" % float((out - ref).abs().max()))
print("max|kernel - CPU oracle|        = %.2e" % float((out.cpu() - flow_front_ref.flow_front(fr0, fr1, flow, gf)).abs().max()))
_lib.prof_enable(True)
t_k = timed(lambda: flow_front(*d))
p = _lib.prof_collect(["flow_front_kernel"])
_lib.prof_enable(False)
t_e = timed(lambda: eager(*d))
t0 = time.perf_counter()
for _ in range(5):
    flow_front_ref.flow_front(fr0, fr1, flow, gf)
t_c = (time.perf_counter() - t0) / 5 * 1e3
k_ms = p["flow_front_kernel"][0] / max(p["flow_front_kernel"][1], 1)
# algorithmic bytes: frames (2 x 3) + flows (4 x 2) read once, 2 x 14 channels written
alg = 4 * H * W * B * (6 + 8 + 28)
print("front end [B=%d, %dx%d]: fused operator %.4f ms (kernel %.4f ms = %.0f GB/s algorithmic), eager torch on the same GPU %.3f ms (%.0fx), CPU oracle %.2f ms (%d threads)"
      % (B, H, W, t_k, k_ms, alg / k_ms / 1e6, t_e, t_e / t_k, t_c, torch.get_num_threads()))
