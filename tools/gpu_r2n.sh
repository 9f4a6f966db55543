#!/bin/bash
# round 2: device-side set-up on the GPU (parity tests, bench line with the device-evaluated
# initial state = new one-GPU anchor)
out=gpurun_out
mkdir -p $out
timeout 900 python -m pytest tests/test_setup.py tests/test_configs.py tests/test_physics.py -m gpu -q -k "setup or device or held or jw_initial or geometry" 2>&1 | tail -6 > $out/r2n_pytest_setup.txt; cat $out/r2n_pytest_setup.txt
timeout 900 python bench.py --steps 20 --warmup 5 2> $out/r2n_bench_n1.err | grep "^{" > $out/r2n_bench_n1.json
tail -2 $out/r2n_bench_n1.err | cut -c1-300
python - <<PY
import json
d=json.load(open('gpurun_out/r2n_bench_n1.json'))
print(d['ms_per_step'], d['value'], d['parity'], d['e2e'], d['roofline']['column_solve'], d['gpu_launches'], d['setup_seconds'])
PY
