#!/bin/bash
out=gpurun_out
mkdir -p $out
timeout 900 python -m pytest tests -m gpu -q -x -k "fused or l30" 2>&1 | tail -4 > $out/r2d_pytest_gpu.txt
cat $out/r2d_pytest_gpu.txt
timeout 300 python tools/kbench.py 2>&1 | grep -v "^{" > $out/r2d_kbench_fused.txt
TB200_DSS_FUSED=0 timeout 300 python tools/kbench.py 2>&1 | grep -v "^{" > $out/r2d_kbench_unfused.txt
cat $out/r2d_kbench_fused.txt $out/r2d_kbench_unfused.txt
