#!/bin/bash
out=gpurun_out
mkdir -p $out
timeout 1200 python -m pytest tests/test_dropin.py tests/test_parity.py -m gpu -q -x 2>&1 | tail -6 > $out/r2f_pytest_gpu.txt
cat $out/r2f_pytest_gpu.txt
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2> $out/r2f_bench_n1.err | grep "^{" > $out/r2f_bench_n1.json
tail -3 $out/r2f_bench_n1.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2f_bench_n1.json'))
print(d['ms_per_step'], d['value'], d['parity'], d['e2e'])
PY
