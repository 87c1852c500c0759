#!/bin/bash
# Round-2 pass on an 8-GPU box (gpurun --gpus 8): real multi-GPU parity tests, cfg4 and cfg5 through the product path.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
(timeout 600 python -m pytest tests/test_gpu_group.py -q -k "multi" 2>&1 | tail -12) > gpurun_out/r02_t_multi_n8.log; cat gpurun_out/r02_t_multi_n8.log
(NCCL_DEBUG=INFO timeout 600 $TR --master-port 29531 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02_bench_cfg4_n8.json 2> gpurun_out/r02_bench_cfg4_n8.err; echo "cfg4 n8 rc=$?")
(timeout 900 $TR --master-port 29532 bench.py --config cfg5 --gpus 8 --steps 10 --warmup 3 > gpurun_out/r02_bench_cfg5_n8.json 2> gpurun_out/r02_bench_cfg5_n8.err; echo "cfg5 n8 rc=$?")
(timeout 600 python bench.py --gpus 8 --single-process --steps 20 --warmup 5 > gpurun_out/r02_bench_cfg4_n8_single_process.json 2> gpurun_out/r02_bench_cfg4_n8_single_process.err; echo "cfg4 n8 single-process (peer stores) rc=$?")
(timeout 600 python bench.py --gpus 8 --single-process --exchange nccl --steps 10 --warmup 3 --no-parity --sweep 1,64 > gpurun_out/r02_bench_cfg4_n8_single_process_nccl.json 2> gpurun_out/r02_bench_cfg4_n8_single_process_nccl.err; echo "cfg4 n8 single-process (nccl) rc=$?")
grep -h "nranks" gpurun_out/r02_bench_cfg4_n8.err | grep "Init COMPLETE" | cut -c1-200 | head -20 > gpurun_out/r02_nccl_info_n8.txt
