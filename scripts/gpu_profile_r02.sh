#!/bin/bash
# Round-2 measurement pass on ONE B200 (run under gpurun): the other BASELINE configs through bench.py, then the ncu
# launch list and one full capture of the dominant kernels -- each ncu run only after the same command exited 0 plain.
mkdir -p gpurun_out
S4='python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity --sweep 1,64'
S3='python bench.py --config cfg3 --batch 64 --steps 2 --warmup 3 --no-cpu-baseline --no-parity --sweep 1'
python bench.py --config cfg2 --steps 20 --warmup 5 > gpurun_out/r02_bench_cfg2.json 2> gpurun_out/r02_bench_cfg2.err; echo "cfg2 rc=$?"
python bench.py --config cfg3 --steps 20 --warmup 5 > gpurun_out/r02_bench_cfg3.json 2> gpurun_out/r02_bench_cfg3.err; echo "cfg3 rc=$?"
$S4 > gpurun_out/plain4.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_bench_cfg4_short.csv $S4 > gpurun_out/ncu4a.log 2>&1
echo "launch list rc=$?"
$S4 > gpurun_out/plain4b.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:scan_mma_kernel -s 6 -c 2 -f -o gpurun_out/r02_prof_scan_mma_cg2_co8_100m $S4 > gpurun_out/ncu4b.log 2>&1
echo "full capture cfg4 rc=$?"
$S3 > gpurun_out/plain3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:scan_mma_small_kernel -s 8 -c 2 -f -o gpurun_out/r02_prof_scan_mma_small_64q_k128_10m $S3 > gpurun_out/ncu3.log 2>&1
echo "full capture cfg3 b64 rc=$?"
tail -n 2 gpurun_out/ncu4a.log gpurun_out/ncu4b.log gpurun_out/ncu3.log
