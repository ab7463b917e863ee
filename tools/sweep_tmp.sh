timeout 300 python -m pytest tests/test_decoder_gpu.py -m gpu -x -q 2>&1 | tail -3
python bench.py --no-cpu-baseline > gpurun_out/s5_bench4.json 2> gpurun_out/s5_bench4.err; tail -3 gpurun_out/s5_bench4.err
python -c "
import json
d=json.load(open('gpurun_out/s5_bench4.json'))
print(d['ms_per_step'], d['e2e'])
print({k: round(v['avg_ms'],3) for k,v in d['kernels'].items()})"
