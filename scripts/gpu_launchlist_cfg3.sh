#!/bin/bash
mkdir -p gpurun_out
S3='python bench.py --config cfg3 --batch 64 --steps 3 --warmup 3 --no-cpu-baseline --no-parity --sweep 1'
$S3 > gpurun_out/plain3.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02_launches_cfg3_b64.csv $S3 > gpurun_out/ncu3a.log 2>&1
echo "launch list rc=$?"
