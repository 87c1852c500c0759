#!/bin/bash
# ncu launch list of the short bench command (the library's kernels only: the corpus build alone is 800 launches)
set -u
OUT=gpurun_out
SHORT="python bench.py --steps 2 --warmup 3 --sweep 1,64 --no-cpu-baseline --no-parity"
$SHORT > $OUT/plain3_final.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'scan_|merge_|rescore|prep_q|retry|rrf|norm_max' -c 600 --csv \
    --log-file $OUT/r02_launches_bench_cfg4_short_final.csv $SHORT > $OUT/ncu_launch_final.log 2>&1
echo "launch list rc=$?"
