#!/bin/bash
# round 2, 8-GPU series: multi-rank parity on the 24-patch decomposition, bench
# lines at N=8 (headline, NCCL-callback variant, config-4 dry stand-in, ne=240 L60)
out=gpurun_out
mkdir -p $out
nvidia-smi topo -m > $out/r2g_topo.txt 2>&1
timeout 900 python -m pytest tests/test_multirank.py -m gpu -q 2>&1 | tail -6 > $out/r2g_pytest_multirank.txt
cat $out/r2g_pytest_multirank.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29701 bench.py --gpus 8 --steps 20 --warmup 5 2> $out/r2g_bench_n8.err | grep "^{" > $out/r2g_bench_n8.json
TB200_EXCHANGE=nccl timeout 600 $TR --master-port 29702 bench.py --gpus 8 --steps 20 --warmup 5 --no-e2e 2> $out/r2g_bench_n8_nccl.err | grep "^{" > $out/r2g_bench_n8_nccl.json
timeout 600 $TR --master-port 29703 bench.py --gpus 8 --ne 60 --tracers 5 --steps 20 --warmup 5 2> $out/r2g_bench_cfg4_n8.err | grep "^{" > $out/r2g_bench_cfg4_n8.json
timeout 900 $TR --master-port 29704 bench.py --gpus 8 --ne 240 --levels 60 --steps 10 --warmup 3 --no-e2e 2> $out/r2g_bench_ne240_l60_n8.err | grep "^{" > $out/r2g_bench_ne240_l60_n8.json
for f in r2g_bench_n8 r2g_bench_n8_nccl r2g_bench_cfg4_n8 r2g_bench_ne240_l60_n8; do
  tail -2 $out/$f.err | cut -c1-300
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/$f.json'))
    print('$f', d['ms_per_step'], d['value'], d['parity']['ok'], d['parity'].get('vs_one_gpu'), d.get('e2e'), d.get('halo_exchange'), d['roofline']['kernels'])
except Exception as e:
    print('$f', 'no line', e)
PY
done
