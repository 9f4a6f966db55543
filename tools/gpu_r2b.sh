#!/bin/bash
# round 2, second series: new diagnostics tests, HBM-sizing run ne=240 L60,
# config-4 dry stand-in (ne=60, 5 tracers) on one GPU with the general kernels
out=gpurun_out
mkdir -p $out
timeout 900 python -m pytest tests -m gpu -q -x -k "diagnostics or lean or l30" 2>&1 | tail -5 > $out/r2b_pytest_gpu.txt
timeout 1500 python bench.py --ne 240 --levels 60 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e \
    2> $out/r2b_bench_ne240_l60_n1.err | grep "^{" > $out/r2b_bench_ne240_l60_n1.json
timeout 900 python bench.py --ne 60 --tracers 5 --steps 10 --warmup 3 --no-cpu-baseline \
    2> $out/r2b_bench_cfg4_n1.err | grep "^{" > $out/r2b_bench_cfg4_n1.json
cat $out/r2b_pytest_gpu.txt
tail -3 $out/r2b_bench_ne240_l60_n1.err; head -c 2500 $out/r2b_bench_ne240_l60_n1.json; echo
tail -3 $out/r2b_bench_cfg4_n1.err; head -c 2500 $out/r2b_bench_cfg4_n1.json
