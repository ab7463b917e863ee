#!/bin/bash
# A/B of tuning switches on the GPU box: for each argument "DEFINES|ENV=.." (space-separated -D names, optional environment)
# rebuild the library, run the decoder parity tests and a short bench.   usage: tools/ab.sh "" "MOTIF_OUT3_CONST" ...
mkdir -p gpurun_out
i=0
for spec in "$@"; do
  defs="${spec%%|*}"; envs=""; [[ "$spec" == *"|"* ]] && envs="${spec#*|}"   # "DEFINES|ENV=1 ENV2=3"
  echo "=== variant $i: defines '${defs}' env '${envs}'"
  MOTIF_DEFINES="$defs" python -m motif_b200.build --force > /dev/null || { echo "build failed"; continue; }
  for kv in $envs; do export "$kv"; done
  timeout 600 python -m pytest tests/test_decoder_gpu.py tests/test_splat_gpu.py tests/test_ref_gpu.py -m gpu -x -q 2>&1 | tail -2
  timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2> gpurun_out/ab_$i.err > gpurun_out/ab_$i.json || tail -5 gpurun_out/ab_$i.err
  python - "$i" <<'PY'
import json, sys
d = json.load(open(f"gpurun_out/ab_{sys.argv[1]}.json"))
print("ms/step", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["ms_per_step"], 3), "clocks", d["clocks"]["sm_mhz"])
print({k: round(v["avg_ms"], 3) for k, v in d["kernels"].items()})
print("splat op", round(d["roofline_splat"]["operator_ms"], 3), "ms, operator frac", round(d["roofline_splat"]["frac"], 3), "gather kernel frac", round(d["roofline_splat"]["kernel_frac"], 3))
PY
  for kv in $envs; do unset "${kv%%=*}"; done
  i=$((i+1))
done
