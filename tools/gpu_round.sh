#!/bin/bash
# One GPU-box session: parity tests, smoke, both bench arms, ncu launch list and one ncu --set full capture.
# Outputs under gpurun_out/<tag>_*.   usage: tools/gpu_round.sh <tag> [skip-tests]
tag=${1:-r1}
mkdir -p gpurun_out
if [ -z "$2" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/${tag}_pytest.log
  timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2
fi
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_ref.json 2> gpurun_out/${tag}_bench_ref.err; echo "ref exit $?"; cat gpurun_out/${tag}_bench_ref.json
timeout 600 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench exit $?"; cat gpurun_out/${tag}_bench.json
timeout 300 python tools/bench_splat.py 2>&1 | tail -8 | tee gpurun_out/${tag}_splat.txt
timeout 300 python tools/bench_ref_gpu.py 2>&1 | tail -12 | tee gpurun_out/${tag}_ref_gpu.txt
timeout 300 python tools/bench_front.py 2>&1 | tail -4 | tee gpurun_out/${tag}_front.txt
timeout 300 python tools/bench_raft_corr.py 2>&1 | tail -3 | tee gpurun_out/${tag}_raft_corr.txt
timeout 300 python tools/bench_dcn.py 2>&1 | tail -2 | tee gpurun_out/${tag}_dcn.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${tag}_launches.log 2>&1; echo "ncu list exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'gather_l0_kernel|flow_bin_q_kernel|synth_q_kernel|imnet_f16_kernel|splat_gather' -s 4 -c 6 \
  -o gpurun_out/${tag}_full -f python tools/run_decode.py --precision f16x3 --reps 2 --splat > gpurun_out/${tag}_full.log 2>&1; echo "ncu full exit $?"
ls -la gpurun_out
