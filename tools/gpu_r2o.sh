#!/bin/bash
# round 2, final one-GPU series: tests, bench lines, per-operation timings, launch list,
# one full ncu capture per kernel that changed this round
out=gpurun_out
mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -4 > $out/r2o_pytest_gpu.txt; cat $out/r2o_pytest_gpu.txt
timeout 900 python bench.py --steps 20 --warmup 5 2> $out/r2o_bench_n1.err | grep "^{" > $out/r2o_bench_n1.json
timeout 600 python bench.py --ne 60 --tracers 5 --steps 20 --warmup 5 --no-cpu-baseline 2> $out/r2o_bench_cfg4_n1.err | grep "^{" > $out/r2o_bench_cfg4_n1.json
python tools/kbench.py 2>&1 | grep -v "^{" > $out/r2o_kbench_ne120.txt
python tools/kbench.py --ne 60 --tracers 5 2>&1 | grep -v "^{" > $out/r2o_kbench_cfg4.txt
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 500 --csv \
    --log-file $out/r2o_launches_ne120.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > $out/r2o_ncu_bench.log 2>&1
python tools/launch_summary.py $out/r2o_launches_ne120.csv > $out/r2o_launches_ne120_summary.txt
# (the .ncu-rep files are reduced to their key metrics on the box: gpurun_out/ is
# limited to 64 MiB)
for k in k_dss_fast k_column_fast; do
    ncu --set full --clock-control none --import-source on -k regex:$k --launch-skip 4 -c 1 \
        -o /tmp/r2o_prof_$k -f python tools/kbench.py --reps 1 > /tmp/r2o_ncu_$k.log 2>&1
    python tools/ncu_key.py /tmp/r2o_prof_$k.ncu-rep > $out/r2o_${k}_key.txt 2>&1
    python tools/ncu_sass_top.py /tmp/r2o_prof_$k.ncu-rep > $out/r2o_${k}_sass_top.txt 2>&1
done
for k in k_tracer_stage_pipe k_column_tracers_fast k_tracer_hyper; do
    ncu --set full --clock-control none --import-source on -k regex:$k --launch-skip 4 -c 1 \
        -o /tmp/r2o_prof_$k -f python tools/kbench.py --ne 60 --tracers 5 --reps 1 > /tmp/r2o_ncu_$k.log 2>&1
    python tools/ncu_key.py /tmp/r2o_prof_$k.ncu-rep > $out/r2o_${k}_key.txt 2>&1
    python tools/ncu_sass_top.py /tmp/r2o_prof_$k.ncu-rep > $out/r2o_${k}_sass_top.txt 2>&1
done
for f in r2o_bench_n1 r2o_bench_cfg4_n1; do
python - <<PY
import json
d=json.load(open('gpurun_out/$f.json'))
print('$f', d['ms_per_step'], d['value'], d['parity']['ok'], d['e2e'] and d['e2e']['ms_per_step'], d['roofline']['frac'], d['roofline']['column_solve']['ms'], d['gpu_launches'], d['cpu_baseline'] and d['cpu_baseline']['value'])
PY
done
cat $out/r2o_kbench_ne120.txt; cat $out/r2o_kbench_cfg4.txt; head -14 $out/r2o_launches_ne120_summary.txt
