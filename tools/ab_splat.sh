#!/bin/bash
# A/B of splat-operator build switches: rebuild, splat parity tests, stand-alone operator timing.
for defs in "$@"; do
  echo "=== defines '${defs}'"
  MOTIF_DEFINES="$defs" python -m motif_b200.build --force > /dev/null || { echo "build failed"; continue; }
  timeout 600 python -m pytest tests/test_splat_gpu.py tests/test_ref_gpu.py -m gpu -x -q 2>&1 | tail -1
  timeout 300 python tools/bench_splat.py 2>&1 | grep -E "operator"
done
