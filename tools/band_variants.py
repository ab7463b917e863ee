"""Time the band-mode gather variants (MOTIF_GATHER_VARIANT) on the whole Adobe frame decoded as ONE band (same work as the full decode)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from motif_b200 import _lib, synthetic
from motif_b200.decoder import SpaceTimeDecoder
H, W, HH, WW, times = synthetic.WORKLOADS["adobe240_x4_t8"]
dec = SpaceTimeDecoder(synthetic.synthetic_params(0), device="cuda")
lat = [t.cuda() for t in synthetic.synthetic_latents(1, H, W, seed=0)]
tt = torch.tensor([times])
band = {} if os.environ.get("FULL") else {"row_range": (0, HH), "halo": 0}
for _ in range(3):
    dec.decode(*lat, tt, (HH, WW), return_flow=False, **band)
torch.cuda.synchronize()
_lib.prof_enable(True)
for _ in range(10):
    dec.decode(*lat, tt, (HH, WW), return_flow=False, **band)
p = _lib.prof_collect(["gather_l0_kernel"])["gather_l0_kernel"]
print("variant", os.environ.get("MOTIF_GATHER_VARIANT", "-"), "FULL" if os.environ.get("FULL") else "band", "gather ms", round(p[0] / p[1], 3))
