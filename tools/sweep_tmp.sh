timeout 300 python -m pytest tests/test_decoder_gpu.py -m gpu -x -q 2>&1 | tail -3
run() { echo "== $*"; env "$@" python bench.py --no-cpu-baseline --steps 10 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],3), {k: round(v['avg_ms'],3) for k,v in d['kernels'].items()})"; }
run A=1
