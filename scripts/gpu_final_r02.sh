#!/bin/bash
# Final measurement pass of round 2 on ONE B200 (under gpurun): GPU tests, the default bench line (cfg4), cfg2 / cfg3, the
# 12.5M-row shard of an 8-GPU run, then the ncu launch list and one full capture of the headline kernel on the same (short)
# bench command -- each ncu run only after the command has exited 0 without it.  Outputs under gpurun_out/.
set -u
OUT=gpurun_out
mkdir -p $OUT
python -m pytest tests -m gpu -x -q > $OUT/r02_t_gpu_final.log 2>&1; echo "pytest rc=$?" >> $OUT/r02_t_gpu_final.log
tail -n 3 $OUT/r02_t_gpu_final.log
python bench.py > $OUT/r02_bench_cfg4_n1_final.json 2> $OUT/bench_cfg4_final.err; echo "cfg4 rc=$?"
python bench.py --config cfg2 > $OUT/r02_bench_cfg2_final.json 2> $OUT/bench_cfg2_final.err; echo "cfg2 rc=$?"
python bench.py --config cfg3 > $OUT/r02_bench_cfg3_final.json 2> $OUT/bench_cfg3_final.err; echo "cfg3 rc=$?"
python bench.py --rows 12500000 --no-cpu-baseline --sweep 1,1024 > $OUT/r02_bench_12m5_rows_final.json 2> $OUT/bench_12m5_final.err; echo "12.5M rc=$?"
SHORT="python bench.py --steps 2 --warmup 3 --sweep 1,64 --no-cpu-baseline --no-parity"
$SHORT > $OUT/plain_final.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/r02_launches_bench_cfg4_short_final.csv $SHORT > $OUT/ncu_launch_final.log 2>&1
echo "launch list rc=$?"
$SHORT > $OUT/plain2_final.log 2>&1 &&
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'scan_mma_kernel<.int.1, .int.2' -s 6 -c 1 -f \
    -o $OUT/r02_prof_scan_mma_cg2_co8_100m_final $SHORT > $OUT/ncu_full_final.log 2>&1
echo "full capture rc=$?"
