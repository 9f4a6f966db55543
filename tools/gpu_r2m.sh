#!/bin/bash
# round 2: full GPU test suite after the tracer / column / DSS changes, config-4 line at N=1
out=gpurun_out
mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 > $out/r2m_pytest_gpu.txt; cat $out/r2m_pytest_gpu.txt
timeout 600 python bench.py --ne 60 --tracers 5 --steps 20 --warmup 5 --no-cpu-baseline 2> $out/r2m_bench_cfg4_n1.err | grep "^{" > $out/r2m_bench_cfg4_n1.json
tail -2 $out/r2m_bench_cfg4_n1.err | cut -c1-300
python - <<PY
import json
d=json.load(open('gpurun_out/r2m_bench_cfg4_n1.json'))
print(d['ms_per_step'], d['value'], d['parity'], d['e2e'], d['roofline']['column_solve'], d['gpu_launches'])
PY
