#!/bin/bash
# One-GPU measurement series of a round (run through gpurun; everything lands in
# gpurun_out/, copy what should be judged into profiles/):
#   gpurun --timeout 1200 -- 'bash tools/measure_round.sh r2'
# Numbers printed by runs under ncu are never bench values.
tag=${1:-rN}
out=gpurun_out
mkdir -p $out
python -m pytest tests -m gpu -q 2>&1 | tail -3 > $out/${tag}_pytest_gpu.txt
python bench.py --steps 20 --warmup 5 2> $out/${tag}_bench_n1.err | grep "^{" > $out/${tag}_bench_n1.json
python bench.py --impl reference --steps 4 --warmup 0 2> /dev/null | grep "^{" > $out/${tag}_bench_reference.json
python tools/kbench.py 2>&1 | grep -v "^{" > $out/${tag}_kbench_ne120.txt
# launch list of three timed steps (own kernels only)
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 500 --csv \
    --log-file $out/${tag}_launches_ne120.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > $out/${tag}_ncu_bench.log 2>&1
python tools/launch_summary.py $out/${tag}_launches_ne120.csv 81 23 > $out/${tag}_launches_ne120_summary.txt
# one full capture per hot kernel (kbench launches each operation alone)
for k in k_nh_stage_pipe k_dss_fast k_column_fast k_hyper_pipe; do
    ncu --set full --clock-control none --import-source on -k regex:$k --launch-skip 7 -c 1 \
        -o $out/${tag}_prof_$k -f python tools/kbench.py --reps 1 > $out/${tag}_ncu_$k.log 2>&1
done
tail -n 3 $out/${tag}_pytest_gpu.txt
cat $out/${tag}_kbench_ne120.txt
