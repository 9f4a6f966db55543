#!/bin/bash
out=gpurun_out
mkdir -p $out
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 > $out/r2h_pytest_gpu.txt
cat $out/r2h_pytest_gpu.txt
timeout 300 python tools/kbench.py 2>&1 | grep -v "^{" > $out/r2h_kbench_tma.txt
cat $out/r2h_kbench_tma.txt
