#!/bin/bash
# quick GPU check of the decoder: parity tests (decoder + full size), then the bench line
tag=${1:-q}
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
timeout 900 python -m pytest tests/test_decoder_gpu.py tests/test_decoder_fullsize_gpu.py -m gpu -q -x > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest exit $?"; tail -25 gpurun_out/${tag}_pytest.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench exit $?"; tail -3 gpurun_out/${tag}_bench.err
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${tag}_bench.json"))
    print("ms/step", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["ms_per_step"], 3), "clocks", d["clocks"]["sm_mhz"])
    print({k: round(v["avg_ms"], 3) for k, v in d["kernels"].items()})
    print("splat", round(d["roofline_splat"]["operator_ms"], 3), "ms, operator frac", round(d["roofline_splat"]["frac"], 3), "gather kernel frac", round(d["roofline_splat"]["kernel_frac"], 3))
except Exception as e:
    print("no bench line", e)
PY
cat gpurun_out/parity_report.jsonl 2>/dev/null | cut -c1-700
