#!/bin/bash
# 8 GPUs: DSS block order A/B (rows fastest = default, groups fastest)
out=gpurun_out
mkdir -p $out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
for rep in 1 2; do
timeout 300 $TR --master-port 2970$rep bench.py --gpus 8 --steps 40 --warmup 5 --no-e2e 2> /dev/null | grep "^{" > $out/r2q_rows_$rep.json
TB200_DSS_ORDER=groups timeout 300 $TR --master-port 2971$rep bench.py --gpus 8 --steps 40 --warmup 5 --no-e2e 2> /dev/null | grep "^{" > $out/r2q_groups_$rep.json
done
python - <<PY
import json
for f in ['rows_1','groups_1','rows_2','groups_2']:
    try:
        d=json.load(open('gpurun_out/r2q_%s.json'%f))
        print(f, d['ms_per_step'], d['roofline']['kernels']['k_dss_fast (1.5 S algorithmic; DRAM floor 2 S)']['ms'], d['roofline']['kernel_ms'], d['roofline']['column_solve']['ms'])
    except Exception as e:
        print(f, e)
PY
