"""FunctionCorrelation at the PWC-Net pyramid sizes of a 720x1280 pair (for ncu: true kernel durations)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from motif_b200.correlation import FunctionCorrelation
torch.manual_seed(0)
for b, c, h, w in ((1, 32, 192, 320), (1, 64, 96, 160), (1, 96, 48, 80), (1, 128, 24, 40), (1, 196, 12, 20)):
    a, bb = torch.randn(b, c, h, w, device="cuda"), torch.randn(b, c, h, w, device="cuda")
    for _ in range(3):
        o = FunctionCorrelation(a, bb)
torch.cuda.synchronize()
print("ok", float(o.abs().mean()))
