#!/bin/bash
# ncu --set full of selected decoder kernels on a short run.  usage: tools/gpu_prof.sh <tag> <kernel regex> [timestamps]
tag=${1:-prof}; rx=${2:-synth_fused}; nts=${3:-2}
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$rx" -s 0 -c 2 -o gpurun_out/${tag} -f python tools/run_decode.py --precision f16x3 --reps 2 --timestamps $nts > gpurun_out/${tag}.log 2>&1; echo "ncu exit $?"; tail -3 gpurun_out/${tag}.log
