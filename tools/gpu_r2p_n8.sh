#!/bin/bash
# round 2, final 8-GPU bench lines (headline, config-4 dry stand-in)
out=gpurun_out
mkdir -p $out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29701 bench.py --gpus 8 --steps 20 --warmup 5 2> $out/r2p_bench_n8.err | grep "^{" > $out/r2p_bench_n8.json
timeout 600 $TR --master-port 29703 bench.py --gpus 8 --ne 60 --tracers 5 --steps 20 --warmup 5 2> $out/r2p_bench_cfg4_n8.err | grep "^{" > $out/r2p_bench_cfg4_n8.json
for f in r2p_bench_n8 r2p_bench_cfg4_n8; do
  tail -2 $out/$f.err | cut -c1-300
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/$f.json'))
    print('$f', d['ms_per_step'], d['value'], d['parity']['ok'], d['parity'].get('vs_one_gpu'), d['e2e'] and d['e2e']['ms_per_step'], d.get('halo_exchange'), d['roofline']['kernel_ms'])
except Exception as e:
    print('$f', 'no line', e)
PY
done
