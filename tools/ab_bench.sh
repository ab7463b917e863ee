#!/bin/bash
# A/B of two trees on the same box: the current tree and a copy under _cmp/ (kernel times of bench.py)
for d in . _cmp .; do
  (cd $d && python bench.py --no-cpu-baseline --steps 20 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.readline()); print('$d', round(d['ms_per_step'],3), {k: round(v['avg_ms'],3) for k,v in d['kernels'].items()})")
done
