#!/bin/bash
# round 2, 8-GPU series after the tracer fast path: multi-rank parity (24 patches on 8
# ranks, tracers on 2), bench lines at N=8 (headline with its parity block, config-4
# dry stand-in on the column-constant path)
out=gpurun_out
mkdir -p $out
timeout 900 python -m pytest tests/test_multirank.py -m gpu -q -k "tracers or bits or (24_patches and 8)" 2>&1 | tail -6 > $out/r2k_pytest_multirank.txt
cat $out/r2k_pytest_multirank.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29701 bench.py --gpus 8 --steps 20 --warmup 5 2> $out/r2k_bench_n8.err | grep "^{" > $out/r2k_bench_n8.json
timeout 600 $TR --master-port 29703 bench.py --gpus 8 --ne 60 --tracers 5 --steps 20 --warmup 5 2> $out/r2k_bench_cfg4_n8.err | grep "^{" > $out/r2k_bench_cfg4_n8.json
for f in r2k_bench_n8 r2k_bench_cfg4_n8; do
  tail -2 $out/$f.err | cut -c1-300
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/$f.json'))
    print('$f', d['ms_per_step'], d['value'], d['parity']['ok'], d['parity'].get('vs_one_gpu'), d.get('e2e'), d.get('halo_exchange'), d['roofline']['kernel_ms'])
except Exception as e:
    print('$f', 'no line', e)
PY
done
