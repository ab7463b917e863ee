"""RAFT schedule of the surround (motif_b200.raft_schedule) beside the reference's four-pair call (Ours.py:544-545), with the
reference's own RAFT-small module.  The reference checkout cannot travel to the GPU box, so this runs in the build container on
the host cores: what it shows is the WORK ratio (8B -> 2B encoded images, 4B -> 2B iterated pairs), not a GPU time."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from motif_b200 import raft_schedule  # noqa: E402
from oracle import ref_shims  # noqa: E402

if not ref_shims.reference_available():
    sys.exit("reference checkout absent")
model = ref_shims.build_reference_model(seed=0)
raft = model.flow_predictor
hh, ww, iters = 256, 448, 4  # Vimeo HR size; iter=4 is what VideoSRBaseModel.test passes (VideoSR_base_model.py:189)
low = torch.rand(2, 3, hh // 8, ww // 8, generator=torch.Generator().manual_seed(0))
fr0, fr1 = torch.nn.functional.interpolate(low, size=(hh, ww), mode="bilinear", align_corners=False).split(1)


def timed(fn, reps=3):
    fn()
    t = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        t.append(time.perf_counter() - t0)
    return sorted(t)[len(t) // 2]


with torch.no_grad():
    four = lambda: raft(torch.cat([fr0, fr0, fr1, fr1]) * 255.0, torch.cat([fr0, fr1, fr0, fr1]) * 255.0, iters=iters)[-1]  # noqa: E731
    sched = lambda: raft_schedule.four_pair_flows(raft, fr0, fr1, iters)  # noqa: E731
    a, b = four(), sched()
    print("max |schedule - four-pair call| on the live pairs: %.2e px (flows up to %.1f px)" % (float((a[1:3] - b[1:3]).abs().max()), float(a.abs().max())))
    t4, t2 = timed(four), timed(sched)
print("RAFT-small, HR %dx%d, %d iterations, %d host threads: reference four-pair call %.3f s, schedule %.3f s (%.2fx)" % (hh, ww, iters, torch.get_num_threads(), t4, t2, t4 / t2))
