"""Pipeline trace of one f16x3 decoder kernel (CTA 0); needs a library built with MOTIF_TRACE=1.

    MOTIF_TRACE=1 python -m motif_b200.build --force; MOTIF_TRACE_KERNEL=1 python tools/trace_f16.py
Event ids: 100*tile + k for the (quad 0, half 0, lane 0) epilogue thread of a tile (1 iteration start, 10/11 D0/D1
ready, 20 A published, 2..6 kernel-specific), 1000 + 100*tile + step = issuer waits done, 2000 + ... = block issued.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from motif_b200 import _lib, synthetic  # noqa: E402
from motif_b200.decoder import SpaceTimeDecoder  # noqa: E402

H, W, HH, WW, times = synthetic.WORKLOADS["adobe240_x4_t8"]
dev = torch.device("cuda:0")
dec = SpaceTimeDecoder(synthetic.synthetic_params(0), device=dev, precision="f16x3")
feat, ff, res = [t.to(dev) for t in synthetic.synthetic_latents(1, H, W, seed=0)]
tt = torch.tensor([times[3:4]])
dec.decode(feat, ff, res, tt, (HH, WW), return_flow=False)
torch.cuda.synchronize()
lib = _lib.load()
cap = 400000
buf = torch.zeros(cap, 2, dtype=torch.int64, device=dev)
lib.motif_tc_set_trace(buf.data_ptr(), cap)
dec.decode(feat, ff, res, tt, (HH, WW), return_flow=False)
torch.cuda.synchronize()
lib.motif_tc_set_trace(None, 0)
ev = buf.cpu()
ev = ev[ev[:, 1] > 0]
ev = ev[torch.argsort(ev[:, 1])]
ids, ts = ev[:, 0].tolist(), ev[:, 1].tolist()
print(f"events {len(ids)} span {ts[-1] - ts[0]} cycles")
starts = [i for i in range(len(ids)) if ids[i] == 1]
print(f"iterations of tile 0: {len(starts)}; mean period {(ts[starts[-1]] - ts[starts[0]]) / max(len(starts) - 1, 1):.0f} cycles")
first = int(os.environ.get("TRACE_FIRST", "10"))
n_it = int(os.environ.get("TRACE_ITERS", "2"))
a, b = starts[first], starts[first + n_it]
skip = (lambda e: 1000 <= e < 3000) if os.environ.get("TRACE_NO_ISSUER") else (lambda e: False)
for i in range(a, b):
    if not skip(ids[i]):
        print(f"   +{ts[i] - ts[a]:7d}  id {ids[i]}")
