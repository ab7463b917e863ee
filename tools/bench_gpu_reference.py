"""The reference's GPU path for the WHOLE hot path (Ours.py:659-858) timed on the B200 beside this repo's decoder --
SURVEY 8d: "the number to beat".  Reference arm = oracle/decoder_ref_gpu.py: the reference's eager torch operators on CUDA
tensors + its own splat kernels compiled unmodified for sm_100a; run (a) as VideoSRBaseModel.test runs it, in chunks of
three timestamps with the clip-invariant part recomputed per chunk (VideoSR_base_model.py:188-193), and (b) all timestamps
in one pass.  Same synthetic Adobe240 clip and weights as bench.py.  Tools-side only: bench.py never imports this.

    python tools/bench_gpu_reference.py [--workload adobe240_x4_t8] [--reps 3] [--out gpurun_out/gpu_reference.json]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from motif_b200 import synthetic  # noqa: E402
from motif_b200.decoder import SpaceTimeDecoder  # noqa: E402
from oracle import decoder_ref_gpu  # noqa: E402


def timed(fn, reps, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="adobe240_x4_t8")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--out", default="gpurun_out/gpu_reference.json")
    a = ap.parse_args()
    assert decoder_ref_gpu.available(), "needs a GPU and oracle/_ref/ref_gpu_sm100a.so"
    H, W, HH, WW, times = synthetic.WORKLOADS[a.workload]
    dev = torch.device("cuda")
    feat, ff, res = [t.to(dev) for t in synthetic.synthetic_latents(1, H, W, seed=0)]
    params = synthetic.synthetic_params(seed=0)
    tt = torch.tensor([times])
    N = len(times)
    units = N * HH * WW
    ms_chunk3 = timed(lambda: decoder_ref_gpu.decode(feat, ff, res, tt, HH, WW, params, chunk=3), a.reps)
    torch.cuda.empty_cache()
    ms_one = timed(lambda: decoder_ref_gpu.decode(feat, ff, res, tt, HH, WW, params, chunk=0), a.reps)
    peak_gb = torch.cuda.max_memory_allocated() / 1e9
    torch.cuda.empty_cache()
    dec = SpaceTimeDecoder(params, device=dev)
    ms_new = timed(lambda: dec.decode(feat, ff, res, tt, (HH, WW)), 20, warm=3)
    out = {"workload": a.workload, "hr": [HH, WW], "timestamps": N,
           "reference_gpu_chunks_of_3_ms": ms_chunk3, "reference_gpu_single_pass_ms": ms_one, "reference_gpu_peak_mem_gb": peak_gb,
           "motif_b200_ms": ms_new,
           "reference_gpu_px_t_per_s": units / (ms_chunk3 * 1e-3), "motif_b200_px_t_per_s": units / (ms_new * 1e-3),
           "speedup_vs_reference_test_loop": ms_chunk3 / ms_new, "speedup_vs_reference_single_pass": ms_one / ms_new,
           "note": "reference arm: eager torch fp32 (cuBLAS SGEMM, allow_tf32=False) + the reference's own splat kernels (nvcc sm_100a, unmodified)"}
    os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
    with open(a.out, "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
