#!/bin/bash
# round 2, first GPU series: parity tests, bench baseline with parity block,
# kernel table, tracer (config-4 stand-in) timing on the general kernels
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=name,memory.total --format=csv > $out/r2a_gpu.txt
free -g | head -2 >> $out/r2a_gpu.txt; nproc >> $out/r2a_gpu.txt
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -40 > $out/r2a_pytest_gpu.txt
timeout 600 python bench.py --steps 20 --warmup 5 2> $out/r2a_bench_n1.err | grep "^{" > $out/r2a_bench_n1.json
timeout 300 python tools/kbench.py 2>&1 | grep -v "^{" > $out/r2a_kbench_ne120.txt
timeout 300 python tools/kbench.py --ne 60 --tracers 5 2>&1 | grep -v "^{" > $out/r2a_kbench_ne60_tr5.txt
tail -n 15 $out/r2a_pytest_gpu.txt
cat $out/r2a_bench_n1.json | head -c 3000
cat $out/r2a_kbench_ne120.txt $out/r2a_kbench_ne60_tr5.txt
