#!/bin/bash
# bench.py on N GPUs of this box (torchrun), summary line.  usage: tools/gpu_multi.sh N [steps]
n=${1:-2}; steps=${2:-20}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps $steps --warmup 3 > gpurun_out/r2_n$n.json 2> gpurun_out/r2_n$n.err
echo "torchrun exit $? ; stdout bytes $(wc -c < gpurun_out/r2_n$n.json)"
grep -v "^\*\*\*\|OMP_NUM" gpurun_out/r2_n$n.err | tail -15
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r2_n$n.json"))
    print(d["n_gpus"], "GPUs:", round(d["ms_per_step"], 3), "ms/clip, e2e", round(d["e2e"]["ms_per_step"], 3), d.get("halo_check"), {k: round(v["avg_ms"], 3) for k, v in d["kernels"].items()})
except Exception as e:
    print("no line:", e)
PY
