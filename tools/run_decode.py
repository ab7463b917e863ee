"""One decode of a named workload (for ncu / timing sessions on the GPU box)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from motif_b200 import synthetic  # noqa: E402
from motif_b200.decoder import SpaceTimeDecoder  # noqa: E402
from motif_b200.softsplat_cp import FunctionSoftsplat  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="adobe240_x4_t8")
ap.add_argument("--precision", default="tf32x3")
ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--timestamps", type=int, default=0)
ap.add_argument("--splat", action="store_true", help="also run the stand-alone softmax splat operator (C=130)")
a = ap.parse_args()
H, W, HH, WW, times = synthetic.WORKLOADS[a.workload]
if a.timestamps:
    times = times[: a.timestamps]
dev = torch.device("cuda:0")
dec = SpaceTimeDecoder(synthetic.synthetic_params(0), device=dev, precision=a.precision)
feat, ff, res = [t.to(dev) for t in synthetic.synthetic_latents(1, H, W, seed=0)]
tt = torch.tensor([times])
for _ in range(a.reps):
    rgb, _ = dec.decode(feat, ff, res, tt, (HH, WW), return_flow=False)
torch.cuda.synchronize()
print("decode ok", tuple(rgb.shape), float(rgb.mean()))
if a.splat:
    torch.manual_seed(0)
    x = torch.randn(1, 130, HH, WW, device=dev)
    low = torch.randn(1, 2, HH // 64, WW // 64, device=dev) * 6  # the smooth field bench.py's roofline_splat uses
    fl = torch.nn.functional.interpolate(low, size=(HH, WW), mode="bilinear", align_corners=False).contiguous()
    z = -torch.rand(1, 1, HH, WW, device=dev)
    for _ in range(a.reps):
        o, n = FunctionSoftsplat(x, fl, z, "softmax")
    torch.cuda.synchronize()
    print("splat ok", float(o.abs().mean()))
