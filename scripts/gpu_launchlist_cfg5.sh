#!/bin/bash
mkdir -p gpurun_out
S5='python bench.py --rows 25000000 --k 100 --batch 1024 --steps 3 --warmup 3 --no-cpu-baseline --no-parity --sweep 64'
$S5 > gpurun_out/plain5.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 1250 -c 900 --csv --log-file gpurun_out/r02_launches_k100_b1024_25m.csv $S5 > gpurun_out/ncu5a.log 2>&1
echo "launch list k100 rc=$?"
S4='python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity --sweep 1,64'
$S4 > gpurun_out/plain4.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 4550 -c 900 --csv --log-file gpurun_out/r02_launches_bench_cfg4_short.csv $S4 > gpurun_out/ncu4a.log 2>&1
echo "launch list cfg4 rc=$?"
