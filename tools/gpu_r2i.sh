#!/bin/bash
out=gpurun_out
mkdir -p $out
# unfused 3 S stage kernel under ncu: 10 stage launches of the two set-up steps + 1 warm-up call
ncu --set full --clock-control none --import-source on -k regex:k_nh_stage_pipe --launch-skip 11 -c 1 \
    -o $out/r2_prof_stage_tma -f python tools/kbench.py --reps 1 --only "stage(copy0)" > $out/r2i_ncu_stage.log 2>&1
tail -2 $out/r2i_ncu_stage.log
timeout 600 python bench.py --steps 20 --warmup 5 2> $out/r2i_bench_n1.err | grep "^{" > $out/r2i_bench_n1.json
tail -2 $out/r2i_bench_n1.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2i_bench_n1.json'))
print(d['ms_per_step'], d['value'], d['parity']['ok'], d['e2e'], d['roofline']['frac'], d['roofline']['kernels'], d['cpu_baseline'])
PY
