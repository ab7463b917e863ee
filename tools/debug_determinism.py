"""Run-to-run differences of the f16x3 decoder on one workload (debug aid): list order is atomic order, so last-bit differences
are expected; anything larger points at an order-dependent discontinuity or a race."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from motif_b200 import synthetic  # noqa: E402
from motif_b200.decoder import SpaceTimeDecoder  # noqa: E402
from oracle import decoder_ref  # noqa: E402

H, W, HH, WW, times = synthetic.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "adobe240_x3p5_t12"]
times = times[:8]
feat, ff, res = [t.cuda() for t in synthetic.synthetic_latents(1, H, W, seed=5)]
params = decoder_ref.random_params(seed=2, **decoder_ref.REALISTIC)
dec = SpaceTimeDecoder(params, device="cuda", precision="f16x3")
tt = torch.tensor([times])
base = None
for run in range(8):
    rgb, flow, pre0 = dec.decode(feat, ff, res, tt, (HH, WW), debug_pre0=True)
    if base is None:
        base = (rgb.clone(), flow.clone(), pre0.clone())
        continue
    d = (rgb - base[0]).abs()
    dp = (pre0 - base[2]).abs()
    print(f"run {run}: flow equal {torch.equal(flow, base[1])}; rgb max diff {d.max().item():.3e} (> 1e-3: {(d > 1e-3).sum().item()}); pre0 max diff {dp.max().item():.3e} (> 1e-4: {(dp > 1e-4).sum().item()})")
    if d.max().item() > 1e-3:
        flat = d.flatten().argmax().item()
        n_, b_, c_, y_, x_ = [int(v) for v in torch.unravel_index(torch.tensor(flat), d.shape)]
        print("   worst rgb at", (n_, y_, x_), "pre0 diff per channel:", [f"{v:.2e}" for v in dp[n_, :, y_, x_].tolist()][:16], "...")
        dv = (pre0[n_, :, y_, x_] - base[2][n_, :, y_, x_]).double().cpu()
        w0 = params["synth_net.net.0.linear.weight"].double()
        for name, col in (("dx'", 64), ("dy'", 65), ("zmax", 130), ("cnt/16", 131), ("wz/cnt", 132), ("t", 197)):
            v = w0[:, col]
            coef = (dv @ v) / (v @ v)
            resid = (dv - coef * v).norm() / dv.norm()
            print(f"   fit along column {name}: coefficient {coef.item():+.5f}, relative residual {resid.item():.3f}")
