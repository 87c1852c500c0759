#!/bin/bash
# Run on the GPU box (via gpurun): GPU tests, the default bench line, then the ncu launch list and one
# full capture of the dominant kernel on the same (short) bench command.  Outputs under gpurun_out/.
set -u
TAG=${1:-s3}
OUT=gpurun_out
mkdir -p $OUT
python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu_$TAG.log
tail -3 $OUT/pytest_gpu_$TAG.log
python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref_$TAG.json 2> $OUT/bench_ref_$TAG.err; echo "ref rc=$?"
SHORT="python bench.py --steps 2 --warmup 3 --sweep 1,1s,8,128 --no-cpu-baseline"
$SHORT > $OUT/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'scan_|merge_|rescore|prep_q|ingest|rrf' -c 600 \
    --csv --log-file $OUT/launches_$TAG.csv $SHORT > $OUT/ncu_launch_$TAG.log 2>&1
echo "launch list rc=$?"
$SHORT > $OUT/plain2_$TAG.log 2>&1 &&
# (the second-chance launches share the function name: select the CTA-pair instance of the first pass by its template arguments)
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'scan_mma_kernel<.int.1, .int.2' -s 6 -c 2 -o $OUT/prof_mma_bench_$TAG -f \
    $SHORT > $OUT/ncu_full_$TAG.log 2>&1
echo "full capture rc=$?"
ncu --set full --clock-control none --import-source on -k regex:scan_mma_small -s 6 -c 2 -o $OUT/prof_mma_small_bench_$TAG -f \
    $SHORT > $OUT/ncu_full_small_$TAG.log 2>&1
echo "full capture (K2s) rc=$?"
