"""Pipeline trace of the tensor-core decoder kernels (CTA 0): prints per-event clock deltas."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from motif_b200 import _lib, synthetic  # noqa: E402
from motif_b200.decoder import SpaceTimeDecoder  # noqa: E402

H, W, HH, WW, times = synthetic.WORKLOADS["adobe240_x4_t8"]
dev = torch.device("cuda:0")
dec = SpaceTimeDecoder(synthetic.synthetic_params(0), device=dev)
feat, ff, res = [t.to(dev) for t in synthetic.synthetic_latents(1, H, W, seed=0)]
tt = torch.tensor([times[:1]])
dec.decode(feat, ff, res, tt, (HH, WW), return_flow=False)
torch.cuda.synchronize()
lib = _lib.load()
cap = 200000
buf = torch.zeros(cap, 2, dtype=torch.int64, device=dev)
lib.motif_tc_set_trace(buf.data_ptr(), cap)
dec.decode(feat, ff, res, tt, (HH, WW), return_flow=False)
torch.cuda.synchronize()
lib.motif_tc_set_trace(None, 0)
ev = buf.cpu()
ev = ev[ev[:, 1] > 0]
ev = ev[torch.argsort(ev[:, 1])]
ids = ev[:, 0].tolist()
ts = ev[:, 1].tolist()
for tag, name in ((0, "imnet"), (100000, "flow_splat"), (200000, "synth")):
    sel = [i for i in range(len(ids)) if tag <= ids[i] < tag + 100000]
    if not sel:
        continue
    print(f"=== {name}: events {len(sel)} span {ts[sel[-1]] - ts[sel[0]]} cycles")
    starts = [i for i in sel if ids[i] - tag == 1000]
    if len(starts) > 6:
        a, b = starts[4], starts[6]
        for i in sel:
            if a <= i < b:
                print(f"   +{ts[i] - ts[a]:7d}  id {ids[i] - tag}")
