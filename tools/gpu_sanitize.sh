#!/bin/bash
# compute-sanitizer memcheck / racecheck over the small-size GPU parity tests (one box session).  usage: tools/gpu_sanitize.sh <tag>
tag=${1:-san}
mkdir -p gpurun_out
run() {  # tool, label, pytest args...
  local tool=$1 label=$2; shift 2
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 77 --print-limit 5 python -m pytest "$@" -m gpu -x -q -p no:cacheprovider > gpurun_out/${tag}_${tool}_${label}.log 2>&1
  echo "$tool $label exit $? : $(grep -c 'ERROR SUMMARY: 0 errors' gpurun_out/${tag}_${tool}_${label}.log) clean summaries; $(grep -E 'passed|failed' gpurun_out/${tag}_${tool}_${label}.log | tail -1)"
  grep -E "Invalid|Race|hazard|ERROR SUMMARY" gpurun_out/${tag}_${tool}_${label}.log | sort | uniq -c | head -8
}
run memcheck ops tests/test_splat_gpu.py tests/test_correlation_gpu.py tests/test_dcn_v2_gpu.py tests/test_raft_corr_gpu.py tests/test_flow_front_gpu.py tests/test_metrics_gpu.py -k "not adobe and not device and not cuda_core"
run memcheck decoder tests/test_decoder_gpu.py -k "golden or stages or spill or ensemble"
run racecheck ops tests/test_splat_gpu.py tests/test_correlation_gpu.py tests/test_dcn_v2_gpu.py -k "not adobe and not device and not cuda_core"
run racecheck decoder tests/test_decoder_gpu.py -k "golden and f16x3"
