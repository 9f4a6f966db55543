#!/bin/bash
out=gpurun_out
mkdir -p $out
# fused stage kernel (3S) under ncu: 10 stage launches of the two set-up steps, 1 warm-up call
ncu --set full --clock-control none --import-source on -k regex:k_nh_stage_pipe --launch-skip 11 -c 1 \
    -o $out/r2e_prof_stage_fused -f python tools/kbench.py --reps 1 --only "stage 3S + dss" > $out/r2e_ncu_stage.log 2>&1
tail -3 $out/r2e_ncu_stage.log
