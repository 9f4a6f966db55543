#!/bin/bash
# round 2: tracers on the column-constant path - parity on the GPU, config-4 dry
# stand-in (ne=60 L30, 5 tracers) at N=1 with the fast and the general kernels,
# launch list of two steps
out=gpurun_out
mkdir -p $out
timeout 900 python -m pytest tests/test_parity.py tests/test_parity_l30.py -m gpu -q -k "tracer" 2>&1 | tail -6 > $out/r2j_pytest_tracers.txt
cat $out/r2j_pytest_tracers.txt
timeout 600 python bench.py --ne 60 --tracers 5 --steps 20 --warmup 5 --no-cpu-baseline 2> $out/r2j_bench_cfg4_n1.err | grep "^{" > $out/r2j_bench_cfg4_n1.json
TB200_TRACER_KERNEL=generic timeout 600 python bench.py --ne 60 --tracers 5 --steps 20 --warmup 5 --no-cpu-baseline --no-e2e 2> $out/r2j_bench_cfg4_n1_generic.err | grep "^{" > $out/r2j_bench_cfg4_n1_generic.json
for f in r2j_bench_cfg4_n1 r2j_bench_cfg4_n1_generic; do
  tail -2 $out/$f.err | cut -c1-300
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/$f.json'))
    print('$f', d['ms_per_step'], d['value'], d['parity'], d.get('e2e'), d['roofline']['kernel_ms'], d['roofline']['frac'], d['roofline']['column_solve'])
except Exception as e:
    print('$f', 'no line', e)
PY
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 600 --csv --log-file $out/r2j_launches_cfg4.csv python bench.py --ne 60 --tracers 5 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > $out/r2j_ncu_bench.log 2>&1
python tools/launch_summary.py $out/r2j_launches_cfg4.csv 2>&1 | tail -30
timeout 600 python -m pytest tests/test_physics.py tests/test_dropin.py -m gpu -q -k "held_suarez" 2>&1 | tail -8 > $out/r2j_pytest_hs.txt
cat $out/r2j_pytest_hs.txt
