#!/bin/bash
# round 2: decomposition independence after the topography-derivative fix, persistent
# column solve against one block per batch, headline bench line at N=1
out=gpurun_out
mkdir -p $out
python tools/decomp_ops.py 12 30 cuda 2>&1 | tail -16 > $out/r2l_decomp_ops.txt; cat $out/r2l_decomp_ops.txt
python tools/kbench.py --only implicit --reps 5 2>&1 | tail -3 > $out/r2l_kbench_column_persistent.txt; cat $out/r2l_kbench_column_persistent.txt
TB200_COLUMN_PERSISTENT=0 python tools/kbench.py --only implicit --reps 5 2>&1 | tail -3 > $out/r2l_kbench_column_chunked.txt; cat $out/r2l_kbench_column_chunked.txt
timeout 900 python bench.py --steps 20 --warmup 5 2> $out/r2l_bench_n1.err | grep "^{" > $out/r2l_bench_n1.json
tail -2 $out/r2l_bench_n1.err | cut -c1-300
python - <<PY
import json
d=json.load(open('gpurun_out/r2l_bench_n1.json'))
print(d['ms_per_step'], d['value'], d['parity'], d['e2e'], d['roofline']['column_solve'], d['gpu_launches'])
PY
