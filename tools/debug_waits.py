"""Run one decode with the expired-wait reporter installed and print which mbarrier waits timed out (debug aid)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from motif_b200 import _lib, synthetic  # noqa: E402
from motif_b200.decoder import SpaceTimeDecoder  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "vimeo_x4"
lib = _lib.load()
torch.zeros(1, device="cuda")
buf = lib.motif_tc_wait_debug_buffer()
H, W, HH, WW, times = synthetic.WORKLOADS[wl]
dec = SpaceTimeDecoder(synthetic.synthetic_params(0), device="cuda", precision="f16x3")
lat = [t.cuda() for t in synthetic.synthetic_latents(1, H, W, seed=0)]
try:
    rgb, _ = dec.decode(*lat, torch.tensor([times]), (HH, WW), return_flow=False)
    torch.cuda.synchronize()
    print("decode ok", float(rgb.mean()))
except Exception as e:  # noqa: BLE001
    print("decode failed:", str(e).splitlines()[0])
n = buf[0]
print("expired waits:", n)
seen = {}
for i in range(min(n, 255)):
    key = (buf[4 + 4 * i], buf[5 + 4 * i], buf[6 + 4 * i], buf[7 + 4 * i] >> 5)
    seen[key] = seen.get(key, 0) + 1
for (addr, par, blk, warp), c in sorted(seen.items(), key=lambda kv: (kv[0][2], kv[0][3])):
    print(f"  block {blk:3d} warp {warp:2d}: barrier smem 0x{addr:x} parity {par} ({c} threads)")
