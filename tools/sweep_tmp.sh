timeout 600 python -m pytest tests/test_splat_gpu.py -m gpu -x -q 2>&1 | tail -2
python tools/bench_splat.py 2>&1 | tail -8
python bench.py --no-cpu-baseline --steps 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(d['roofline_splat'])"
