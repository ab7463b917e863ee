import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from motif_b200 import _lib
lib = _lib.load()
out = torch.zeros(2, dtype=torch.int64, device="cuda")
for a_tmem in (1, 0):
    for n, n_acc in ((64, 1), (64, 2), (128, 1), (128, 2), (256, 1)):
        reps = 192
        for _ in range(2):
            _lib.check(lib.motif_tc_mma_rate(out.data_ptr(), n, reps, a_tmem, n_acc, None), "rate")
            torch.cuda.synchronize()
        tot, iss = out.tolist()
        print(f"A_in_tmem={a_tmem} N={n:3d} accumulators={n_acc}: total {tot:6d} cyc ({tot/reps:6.1f}/mma), issue {iss:6d} cyc ({iss/reps:5.1f}/mma)")
