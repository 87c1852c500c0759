#!/bin/bash
# A/B: companions in the multi-way compaction (mma_debug 1024 = off), 32 pending slots + 4-stage ring at k' = 32 (variant build)
set -x
mkdir -p gpurun_out
V=financial_rag_b200/libfrb200_pend32.so
K="mma_path or first_tile or co_resident or scheduling or second_chance or deletes"
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$K" > gpurun_out/t_parity_ab2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t_parity_ab2.log
FRB200_LIB=$V timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$K" > gpurun_out/t_parity_ab2_variant.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t_parity_ab2_variant.log
tail -3 gpurun_out/t_parity_ab2.log gpurun_out/t_parity_ab2_variant.log
O=gpurun_out/sweep_rows_fixed_cost_ab2.jsonl
: > $O
for rep in 1 2; do
echo '{"lib": "default"}' >> $O
FR_SWEEP_SIZES=390625,12500000 FR_SWEEP_DBG=0,1024,768 timeout 300 python scripts/sweep_rows.py 4096 8 6 >> $O 2>> gpurun_out/sweep_ab2.err
echo '{"lib": "pend32"}' >> $O
FRB200_LIB=$V FR_SWEEP_SIZES=390625,12500000 FR_SWEEP_DBG=0,1024 timeout 300 python scripts/sweep_rows.py 4096 8 6 >> $O 2>> gpurun_out/sweep_ab2.err
done
echo '{"lib": "default"}' >> $O
FR_SWEEP_SIZES=10000000 FR_SWEEP_DBG=0,1024,768 timeout 300 python scripts/sweep_rows.py 1024 4 6 >> $O 2>> gpurun_out/sweep_ab2.err
FR_SWEEP_SIZES=10000000,40000000 FR_SWEEP_DBG=0,1024,768 timeout 300 python scripts/sweep_rows.py 128 1 6 >> $O 2>> gpurun_out/sweep_ab2.err
FR_SWEEP_SIZES=10000000,40000000 FR_SWEEP_DBG=0,768 timeout 300 python scripts/sweep_rows.py 256 1 6 >> $O 2>> gpurun_out/sweep_ab2.err
echo '{"lib": "pend32"}' >> $O
FRB200_LIB=$V FR_SWEEP_SIZES=10000000 FR_SWEEP_DBG=0 timeout 300 python scripts/sweep_rows.py 1024 4 6 >> $O 2>> gpurun_out/sweep_ab2.err
FRB200_LIB=$V FR_SWEEP_SIZES=10000000,40000000 FR_SWEEP_DBG=0 timeout 300 python scripts/sweep_rows.py 128 1 6 >> $O 2>> gpurun_out/sweep_ab2.err
FRB200_LIB=$V FR_SWEEP_SIZES=10000000,40000000 FR_SWEEP_DBG=0 timeout 300 python scripts/sweep_rows.py 256 1 6 >> $O 2>> gpurun_out/sweep_ab2.err
cat $O
