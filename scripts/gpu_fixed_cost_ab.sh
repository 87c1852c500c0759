#!/bin/bash
# A/B of the K2 launch's fixed cost (first-tile bound: mma_debug 256 switches it off; multi-way compaction: 512 = one by one)
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/t_parity_ab.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t_parity_ab.log
tail -3 gpurun_out/t_parity_ab.log
FR_SWEEP_SIZES=390625,12500000 FR_SWEEP_DBG=0,768,256,512,0,768 timeout 300 python scripts/sweep_rows.py 4096 8 6 > gpurun_out/sweep_rows_fixed_cost_ab.jsonl 2> gpurun_out/sweep_ab.err
FR_SWEEP_SIZES=10000000 FR_SWEEP_DBG=0,768,0,768 timeout 300 python scripts/sweep_rows.py 1024 4 6 >> gpurun_out/sweep_rows_fixed_cost_ab.jsonl 2>> gpurun_out/sweep_ab.err
FR_SWEEP_SIZES=10000000 FR_SWEEP_DBG=0,768,0,768 timeout 300 python scripts/sweep_rows.py 128 1 6 >> gpurun_out/sweep_rows_fixed_cost_ab.jsonl 2>> gpurun_out/sweep_ab.err
cat gpurun_out/sweep_rows_fixed_cost_ab.jsonl
