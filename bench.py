#!/usr/bin/env python
"""Benchmark of the MoTIF per-pixel inference hot path on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port)

A "step" is one pass of the hot path (Ours.py:659-858: gather -> imnet/flow_imnet -> 3 splats x 2
references -> blend -> synth_net -> clamp) over one synthetic Adobe240-shaped clip
(180x320 -> 720x1280, 7 intermediate timestamps) from resident LR latents.  `value` is output
pixel-timestamps per second of the whole job, inputs resident in HBM; `e2e` is the same through the
public host-buffer API (`ClipStream.submit`) with pinned HOST latents copied in and the frames copied
out inside the timed region, double-buffered against the neighbouring clips' decode.  With N > 1 ranks (torchrun) rank 0 owns the latents, broadcasts them over
NCCL inside every step (the path's one exchange) and each rank decodes its own timestamps.
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "hr_output_pixels_per_sec"
UNIT = "px-t/s"

# algorithmic work per unit (BASELINE.md section 3 / SURVEY.md 8d)
FLOP_FLOW_IMNET_ROW = 2 * 25536  # per (reference, pixel, timestamp)
FLOP_IMNET_ROW = 2 * 41088       # per (reference, pixel)
FLOP_SYNTH_ROW = 2 * 38016       # per (pixel, timestamp)
SPLAT_BYTES_PER_SRC = 1056       # softmax splat, C=130: 4*(2C+4)


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"hbm_gbs": p["hbm_gbs"], "bf16_burst": p["bf16_tflops"], "bf16_sustained": p.get("bf16_tflops_sustained", p["bf16_tflops"]),
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "bf16_burst": 1590.0, "bf16_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


def load_traffic():
    """DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of each kernel from the committed
    `ncu --set full` capture of this workload (profiles/r2_traffic.json; written by tools/ncu_traffic.py)."""
    path = os.path.join(ROOT, "profiles", "r2_traffic.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f).get("traffic_bytes_per_launch", {})
    return {}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""

    FIELDS = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "200", "-i", str(self.gpu)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            parts = [x.strip() for x in line.split(",")]
            if len(parts) >= 8:
                self.rows.append(parts)

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx = float(r[2])
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def cpu_reference_run(workload, steps, warmup, sample_note):
    """The reference's CPU implementation of the path (oracle port of Ours.py:659-858 with the cupy
    splats transcribed to index_add_), all host threads, on a bounded sample of the workload."""
    from motif_b200 import synthetic
    from oracle import decoder_ref

    H, W, HH, WW, times = workload
    torch.set_num_threads(os.cpu_count() or 1)
    feat, ff, res = synthetic.synthetic_latents(1, H, W, seed=0)
    params = synthetic.synthetic_params(seed=0)
    tt = torch.tensor([times])
    durations = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        decoder_ref.decode(feat, ff, res, tt, HH, WW, params)
        dt = time.perf_counter() - t0
        if i >= warmup:
            durations.append(dt)
    total = sum(durations)
    units = len(times) * HH * WW * len(durations)
    return {"value": units / total, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port", "sample": sample_note,
            "ms_per_step": 1e3 * total / len(durations)}


# bounded CPU sample of the Adobe workload: same x4 ratio and 7 timestamps on a 45x80 -> 180x320 crop
CPU_SAMPLE = (45, 80, 180, 320, [k / 8 for k in range(1, 8)])
CPU_SAMPLE_NOTE = "Adobe240 workload cropped to LR 45x80 -> 180x320, all 7 timestamps (1/16 of the pixels), fp32, torch CPU"


_REAL_STDOUT = None


def emit(line: dict):
    """The ONE JSON line of the contract, on the process's real stdout."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    # Libraries may chat on stdout (NCCL prints its version line there on the first communicator): everything but the JSON line goes
    # to stderr, so that the driver always finds exactly one line.
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="motif", choices=["motif", "reference"])
    ap.add_argument("--workload", default="adobe240_x4_t8")
    ap.add_argument("--precision", default="f16x3", choices=["f16x3", "tf32x3", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--parallel", default="bands", choices=["bands", "clips"],
                    help="N > 1: 'bands' = ONE clip sharded by destination row bands (BASELINE config 3, strong scaling); "
                         "'clips' = one clip per GPU, no collective (BASELINE config 5: --workload uhd4k_x4_t8, weak scaling)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    n_ranks = world                      # processes of the job
    per_clip = args.parallel == "clips" and world > 1
    if per_clip:
        world = 1                        # from here on `world` = ranks that share ONE clip; every rank decodes its own clip alone

    from motif_b200 import synthetic

    H, W, HH, WW, times = synthetic.WORKLOADS[args.workload]
    B, N = 1, len(times)
    qs = HH * WW
    HALO = 16  # source halo (HR rows) of a destination row band; checked against the largest |flow_y| after the timed loops
    config = {"workload": f"{args.workload}: LR {H}x{W} -> HR {HH}x{WW}, {N} timestamps, B=1, synthetic latents + synthetic best.pth-layout weights",
              "l2": "per-step working set (per-source rows 472 MB + destination lists 118 MB + frames 77 MB) >> 126 MB L2; no explicit flush",
              "parallelism": f"one clip per GPU on {n_ranks} GPUs, no collective (data parallel by clip)" if per_clip else (f"destination row bands (+{HALO}-row source halo, verified) over {world} ranks, every rank all {N} timestamps; NCCL broadcast of the "
                              "LR latents per step on a side stream, overlapped with the previous step's decode") if world > 1 else "single GPU"}

    if args.impl == "reference":
        # The reference's CPU implementation of the path (oracle port), all host threads, rank 0 only.  Every step is a
        # BOUNDED SAMPLE of the workload (1/16 of the pixels, all 7 timestamps: ~1.5 s per step on 16 cores), the steps
        # and warm-ups run are the ones asked for, and config.workload names the crop -- so this line describes what ran.
        if rank != 0:
            return
        steps, warm = max(1, args.steps), max(0, args.warmup)
        r = cpu_reference_run(CPU_SAMPLE, steps, warm, CPU_SAMPLE_NOTE)
        ch, cw, chh, cww, ctimes = CPU_SAMPLE
        ref_config = {"workload": f"{args.workload} CROPPED to LR {ch}x{cw} -> HR {chh}x{cww} (1/16 of the pixels), {len(ctimes)} timestamps, B=1, "
                                  "synthetic latents + synthetic best.pth-layout weights; bounded CPU sample of the GPU arm's workload",
                      "l2": "n/a (host)", "parallelism": f"{r['cores']} host threads, rank 0 only"}
        line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": warm,
                "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": ref_config, "gpu_launches": 0,
                "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": r["sample"]},
                "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        emit(line)
        return

    import torch.distributed as dist

    from motif_b200 import _lib, sharding
    from motif_b200.decoder import SpaceTimeDecoder
    from motif_b200.softsplat_cp import FunctionSoftsplat

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if n_ranks > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()

    params = synthetic.synthetic_params(seed=0)
    dec = SpaceTimeDecoder(params, device=dev, precision=args.precision)
    feat_h, ff_h, res_h = [t.pin_memory() for t in synthetic.synthetic_latents(B, H, W, seed=rank if per_clip else 0)]
    if rank == 0 or per_clip:
        feat, ff, res = feat_h.to(dev), ff_h.to(dev), res_h.to(dev)
    else:
        feat, ff, res = [torch.empty_like(t, device=dev) for t in (feat_h, ff_h, res_h)]
    tt = torch.tensor([times])
    # N > 1: the second sharding axis of SURVEY 8e -- every rank decodes a band of destination rows of ALL timestamps (7
    # timestamps do not divide over 2 / 4 / 8 ranks, and imnet / the LR tables would be replicated); N == 1: the whole frame
    n0, n1 = 0, N
    r0, r1 = sharding.partition_rows(HH, world)[rank] if world > 1 else (0, HH)
    band = {"row_range": (r0, r1), "halo": HALO} if world > 1 else {}
    out_h = torch.empty(N, B, 3, max(r1 - r0, 1), WW, dtype=torch.float32).pin_memory()
    stat = torch.zeros(64, dtype=torch.float32, device=dev)

    from motif_b200.clip_stream import ClipStream

    # e2e at N > 1: the clip sits in pinned host memory that every rank can read (here: every rank pins the same synthetic clip);
    # each rank copies 1 / N of it over its own PCIe link and the parts are all-gathered over NVLink (ClipStream.sliced_copy_in)
    # (with destination row bands even that exchange is unnecessary: every rank pulls only the LR rows its band and halo read,
    #  ClipStream.band_copy_in -- no collective on the end-to-end arm)
    stream = ClipStream(dec, depth=2, distributed=world > 1, src=0, return_flow=True, sliced_copy_in=world > 1, band_copy_in=world > 1)
    lat_shapes = tuple(tuple(t.shape) for t in (feat_h, ff_h, res_h))
    rgb_dev = torch.empty(N, B, 3, HH, WW, dtype=torch.float32, device=dev)
    exchange = sharding.LatentExchange(lat_shapes, dev, src=0) if world > 1 else None

    def step(from_host: bool):
        """One clip.  Resident arm: latents already in HBM on rank 0 (broadcast + decode only).  Host arm (e2e): the
        public host-buffer API -- pinned latents copied in, frames copied out, double-buffered against the
        neighbouring clips' decode (motif_b200/clip_stream.py); every clip's copies happen inside the timed region."""
        if from_host:
            stream.submit(feat_h, ff_h, res_h, tt, (HH, WW), out_h, n_range=(n0, n1), **band)
            return
        if world > 1:
            f2, g2, r2 = exchange.take()                       # this step's broadcast (started during the previous step)
            exchange.start((feat, ff, res) if rank == 0 else None)   # next step's, behind this step's decode
        else:
            f2, g2, r2 = feat, ff, res
        if r1 > r0:
            dec.decode(f2, g2, r2, tt, (HH, WW), n_range=(n0, n1), return_flow=True, out=rgb_dev, flow_y_max=stat if world > 1 else None, **band)  # both outputs of the forward (Ours.py:858)
        if world > 1:
            exchange.release()

    def timed(from_host: bool, steps: int, profile: bool):
        if n_ranks > 1:
            dist.barrier()
        if world > 1:
            if not from_host:
                exchange.start((feat, ff, res) if rank == 0 else None)   # prime the pipeline: the timed region does `steps` broadcasts
        torch.cuda.synchronize()
        if profile:
            _lib.prof_enable(True)
        lib.motif_reset_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            step(from_host)
        if from_host:  # the copy-out stream's tail belongs to the timed region
            torch.cuda.current_stream().wait_stream(stream.s_out)
        e1.record()
        torch.cuda.synchronize()
        if world > 1 and not from_host:
            exchange.take()   # drain the broadcast started by the last step (outside the timed region: it belongs to the next clip)
            exchange.release()
            torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        launches = lib.motif_launch_count()
        if n_ranks > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
            dist.barrier()
        return ms, launches

    if world > 1:
        exchange.start((feat, ff, res) if rank == 0 else None)
    for _ in range(max(args.warmup, 3)):
        step(False)
    if world > 1:
        exchange.take()
        exchange.release()
    torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    kernel_names = ["imnet_kernel", "flow_splat_kernel", "synth_kernel", "imnet_tc_kernel", "flow_splat_tc_kernel", "synth_tc_kernel",
                    "imnet_f16_kernel", "flow_bin_f16_kernel", "synth_f16_kernel", "gather_l0_kernel"]
    ms, launches = timed(False, args.steps, profile=True)
    prof = _lib.prof_collect(kernel_names)
    _lib.prof_enable(False)
    clocks = sampler.stop() if rank == 0 else None

    for _ in range(3):
        step(True)
    stream.synchronize()
    ms_e2e, _ = timed(True, args.steps, profile=False)
    stream.synchronize()

    halo_check = None
    if world > 1:  # the band decodes were exact iff no source moved farther than the halo allows
        fy_resident = sharding.check_halo(stat, HALO)
        fy_host = sharding.check_halo(stream.flow_y_max, HALO)
        assert max(fy_resident, fy_host) < HALO - 1, f"source halo of {HALO} rows violated: max |flow_y| = {max(fy_resident, fy_host)}"
        halo_check = {"halo_rows": HALO, "max_abs_flow_y_px": max(fy_resident, fy_host)}
    if rank != 0:
        if n_ranks > 1:
            dist.destroy_process_group()
        return

    peaks = load_peaks()
    clips = n_ranks if per_clip else 1
    units_per_step = N * qs
    value = clips * units_per_step * args.steps / (ms * 1e-3)
    e2e_value = clips * units_per_step * args.steps / (ms_e2e * 1e-3)
    h2d = clips * (feat_h.numel() + ff_h.numel() + res_h.numel()) * 4
    if world > 1:  # every rank pulled the LR rows of its own band and halo (the halos overlap, so the sum exceeds one clip)
        rows = sum(b - a for a, b in (SpaceTimeDecoder.lr_rows_of_band(H, HH, band_r, HALO) for band_r in sharding.partition_rows(HH, world) if band_r[1] > band_r[0]))
        h2d = 5 * B * 64 * rows * W * 4
    d2h = clips * N * B * 3 * qs * 4  # all ranks together (each copies its own rows / clip out)

    # ---- roofline of the dominant kernel of the step (per launch, CUDA events on the launch stream) ----
    # Algorithmic work per launch.  Tensor-bound kernels: dense MACs of the reference's layers
    # counted ONCE (the three split-product passes are not extra algorithmic FLOPs, so an error-compensated path
    # cannot exceed 1/3 of the peak by construction).  HBM-bound kernel (gather_l0), bytes that MUST cross HBM per destination
    # pixel and timestamp: its list entries (8 x 8 B on average), side/zmax/count read + re-arm (48 B), the 256-byte fp16
    # hi/lo A operand written for synth_net, and the per-source rows of both reference frames (2 x 256 B) ONCE PER CLIP -- the
    # timestamps of a group share them through L2 (band-major CTA order) -- i.e. 512 / N bytes per timestamp.  (Round 1
    # counted the rows once per timestamp, 880 B, which exceeded the measured DRAM traffic and flattered the kernel.)
    GATHER_BYTES_PER_DST = 64 + 48 + 256 + (2 * 256) / N
    work = {
        "imnet_kernel": ("tensor", FLOP_IMNET_ROW * 2 * B * qs), "imnet_tc_kernel": ("tensor", FLOP_IMNET_ROW * 2 * B * qs),
        "flow_splat_kernel": ("tensor", FLOP_FLOW_IMNET_ROW * 2 * qs), "flow_splat_tc_kernel": ("tensor", FLOP_FLOW_IMNET_ROW * 2 * qs),
        "synth_kernel": ("tensor", FLOP_SYNTH_ROW * qs), "synth_tc_kernel": ("tensor", FLOP_SYNTH_ROW * qs),
        "imnet_f16_kernel": ("tensor", FLOP_IMNET_ROW * 2 * B * qs), "flow_bin_f16_kernel": ("tensor", FLOP_FLOW_IMNET_ROW * 2 * qs),
        "synth_f16_kernel": ("tensor", FLOP_SYNTH_ROW * qs), "gather_l0_kernel": ("hbm", GATHER_BYTES_PER_DST * qs),
    }
    live = {k: v for k, v in prof.items() if v[1] > 0}
    # tensor peak of the operand type the MMAs run in: kind::f16 = the measured bf16/fp16 dense figure, kind::tf32 = half of it
    # The timed region decides which measured peak applies (B200_PROFILING.md): a region shorter than ~1 s runs at burst
    # clocks, a seconds-long one under the power cap.  Both fractions are printed; `frac` uses the regime that ran.
    burst_regime = ms < 1000.0
    base_peak = peaks["bf16_burst"] if burst_regime else peaks["bf16_sustained"]
    regime = f"{'burst' if burst_regime else 'sustained'} peak (timed region {ms / 1e3:.2f} s)"
    if args.precision == "f16x3":
        tensor_peak, peak_name, peak_div = base_peak, f"fp16 dense = bf16 {regime}", 1.0
    else:
        tensor_peak, peak_name, peak_div = base_peak / 2.0, f"TF32 dense = 0.5 x bf16 {regime}", 2.0
    traffic = load_traffic() if world == 1 else {}  # the committed ncu capture is of the single-GPU launch (all 7 timestamps per launch)
    kernels = {}
    once_per_clip = ("imnet_kernel", "imnet_tc_kernel", "imnet_f16_kernel")
    for k, (tot_ms, cnt) in live.items():
        avg = tot_ms / cnt
        bound, amount = work[k]
        if k not in once_per_clip:  # `work` is per timestamp; a launch covers a group of timestamps (f16x3: all of this rank's)
            amount = amount * (n1 - n0) * args.steps / cnt * ((r1 - r0) / HH)  # this rank's rows
        rate = amount / (avg * 1e-3) / (1e12 if bound == "tensor" else 1e9)
        peak = tensor_peak if bound == "tensor" else peaks["hbm_gbs"]
        kernels[k] = {"launches_per_step": cnt / args.steps, "avg_ms": avg, "share_of_step": tot_ms / ms, "bound": bound,
                      "achieved": rate, "unit": "TFLOP/s" if bound == "tensor" else "GB/s", "frac": rate / peak,
                      "traffic": traffic.get(k)}
    roofline = None
    if live:
        dom = max(live, key=lambda k: live[k][0])
        kd = kernels[dom]
        peak = tensor_peak if kd["bound"] == "tensor" else peaks["hbm_gbs"]
        note = (f"{peak_name}, {peaks['source']}; algorithmic FLOPs (K un-padded, one pass counted; 3 split-product passes run)" if kd["bound"] == "tensor"
                else f"HBM copy bandwidth, {peaks['source']}; {GATHER_BYTES_PER_DST:.0f} algorithmic bytes per destination pixel and timestamp")
        roofline = {"kernel": dom, "bound": kd["bound"], "achieved": kd["achieved"], "peak": peak, "unit": kd["unit"], "frac": kd["frac"],
                    "traffic": kd["traffic"], "peak_note": note}
        step_flops = (2 * FLOP_FLOW_IMNET_ROW + FLOP_SYNTH_ROW) * N * qs + FLOP_IMNET_ROW * 2 * B * qs
        roofline["whole_step_tflops"] = step_flops * args.steps / (ms * 1e-3) / 1e12
        roofline["whole_step_tensor_frac"] = roofline["whole_step_tflops"] / tensor_peak
        if kd["bound"] == "tensor":
            roofline["frac_of_burst_peak"] = kd["achieved"] / (peaks["bf16_burst"] / peak_div)
            roofline["frac_of_sustained_peak"] = kd["achieved"] / (peaks["bf16_sustained"] / peak_div)

    # ---- HBM roofline of the stand-alone softmax splat operator (C=130, one 720x1280 reference frame) ----
    torch.manual_seed(0)
    x = torch.randn(1, 130, HH, WW, device=dev)
    low = torch.randn(1, 2, HH // 64, WW // 64, device=dev) * 6  # smooth flow field, ~10% local stretch
    fl = torch.nn.functional.interpolate(low, size=(HH, WW), mode="bilinear", align_corners=False).contiguous()
    z = -torch.rand(1, 1, HH, WW, device=dev)
    for _ in range(3):
        FunctionSoftsplat(x, fl, z, "softmax")
    torch.cuda.synchronize()
    _lib.prof_enable(True)
    reps = 10
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record()
    for _ in range(reps):
        FunctionSoftsplat(x, fl, z, "softmax")
    s1.record()
    torch.cuda.synchronize()
    sp = _lib.prof_collect(["splat_bin_kernel", "splat_gather_kernel", "splat_scatter_kernel"])
    _lib.prof_enable(False)
    op_ms = s0.elapsed_time(s1) / reps
    gather_ms = sp["splat_gather_kernel"][0] / max(sp["splat_gather_kernel"][1], 1)
    alg_bytes = SPLAT_BYTES_PER_SRC * qs
    # `frac` is the OPERATOR (bin + gather + surplus launches + memset, as a caller of FunctionSoftsplat sees it); the gather kernel alone is kernel_frac
    roofline_splat = {"kernel": "FunctionSoftsplat operator (splat_bin + splat_gather + surplus)", "bound": "hbm", "achieved": alg_bytes / (op_ms * 1e-3) / 1e9,
                      "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": alg_bytes / (op_ms * 1e-3) / 1e9 / peaks["hbm_gbs"], "traffic": traffic.get("splat_gather_kernel"),
                      "kernel_gbs": alg_bytes / (gather_ms * 1e-3) / 1e9, "kernel_frac": alg_bytes / (gather_ms * 1e-3) / 1e9 / peaks["hbm_gbs"],
                      "operator_gbs": alg_bytes / (op_ms * 1e-3) / 1e9, "operator_ms": op_ms,
                      "bin_ms": sp["splat_bin_kernel"][0] / max(sp["splat_bin_kernel"][1], 1),
                      "note": f"FunctionSoftsplat softmax, [1,130,{HH},{WW}], 1056 B per source pixel; {peaks['source']}"}
    del x, fl, z

    cpu_baseline = None
    if n_ranks == 1 and not args.no_cpu_baseline:
        r = cpu_reference_run(CPU_SAMPLE, 2, 1, CPU_SAMPLE_NOTE)
        cpu_baseline = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": r["sample"]}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_ranks, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak" if per_clip else "strong", "vs_baseline": None,
        "dtype": {"f16x3": "f16x3 (two-piece fp16 split, fp32 accumulate; fp32-equivalent)", "tf32x3": "tf32x3 (fp32-equivalent)", "fp32": "f32"}[args.precision],
        "data": "synthetic", "config": config,
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e / args.steps,
                "api": ("ClipStream.submit: copy-in, decode and copy-out of consecutive clips overlap on three streams (depth 2); every clip's own copies are inside the timed region"
                        + ("; every rank pulls the LR rows of its own band (+ halo) out of the pinned clip over its own PCIe link (strided copy, no collective) and copies its own band of the frames out" if world > 1 else ""))},
        "gpu_launches": launches, "halo_check": halo_check,
        "roofline": roofline, "roofline_splat": roofline_splat, "kernels": kernels,
        "cpu_baseline": cpu_baseline,
    }
    emit(line)
    if n_ranks > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
