#!/bin/bash
# round 2, third series: DSS fused into the stage / hyperdiffusion kernels
out=gpurun_out
mkdir -p $out
timeout 900 python -m pytest tests -m gpu -q -x -k "fused or l30 or nonhydro_steps or configs" 2>&1 | tail -8 > $out/r2c_pytest_gpu.txt
cat $out/r2c_pytest_gpu.txt
timeout 300 python tools/kbench.py 2>&1 | grep -v "^{" > $out/r2c_kbench_fused.txt
TB200_DSS_FUSED=0 timeout 300 python tools/kbench.py 2>&1 | grep -v "^{" > $out/r2c_kbench_unfused.txt
cat $out/r2c_kbench_fused.txt $out/r2c_kbench_unfused.txt
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2> $out/r2c_bench_n1.err | grep "^{" > $out/r2c_bench_n1.json
tail -3 $out/r2c_bench_n1.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2c_bench_n1.json'))
print(d['ms_per_step'], d['value'], d['parity'], d['e2e'])
PY
