#!/bin/bash
# tests (decoder + tc only unless ALL=1) then bench; prints ms/step and per-kernel ms
if [ -n "$ALL" ]; then python -m pytest tests -m gpu -x -q 2>&1 | tail -4; else python -m pytest tests/test_decoder_gpu.py -m gpu -x -q 2>&1 | tail -4; fi
python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>gpurun_out/qb.err > gpurun_out/qb.json || tail -5 gpurun_out/qb.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/qb.json"))
print("ms/step", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["ms_per_step"], 3), "clocks", d["clocks"]["sm_mhz"])
print({k: round(v["avg_ms"], 3) for k, v in d["kernels"].items()})
print("splat", round(d["roofline_splat"]["operator_ms"], 3), "ms, operator frac", round(d["roofline_splat"]["frac"], 3), "gather kernel frac", round(d["roofline_splat"]["kernel_frac"], 3))
PY
