#!/bin/bash
# pipeline trace of synth_fused_kernel (needs the MOTIF_TRACE build shipped in the snapshot)
mkdir -p gpurun_out
MOTIF_TRACE_KERNEL=2 TRACE_NO_ISSUER=1 TRACE_FIRST=${1:-6} TRACE_ITERS=${2:-2} timeout 300 python tools/trace_f16.py > gpurun_out/trace_fused.txt 2>&1; tail -150 gpurun_out/trace_fused.txt
