#!/bin/bash
# round-2 session A: parity (incl. the new full-size / regime tests), smoke, bench, reference-GPU timing
tag=${1:-r2a}
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest exit $?"; tail -15 gpurun_out/${tag}_pytest.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench exit $?"; cat gpurun_out/${tag}_bench.json
timeout 600 python tools/bench_gpu_reference.py --out gpurun_out/${tag}_gpu_reference.json 2>&1 | tail -3
cat gpurun_out/parity_report.jsonl
