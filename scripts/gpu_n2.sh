#!/bin/bash
# 2-GPU sanity of the end-of-round code: the driver's launch line for N = 2, and the NCCL all-reduce test (run under gpurun --gpus 2)
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 50 --warmup 10 --no-extras --no-eager-baseline --no-cpu-baseline 2>&1 | grep '^{' | tail -1 | tee gpurun_out/n2_bench.json
timeout 300 python -m pytest tests/test_allreduce_gpu.py -x -q -m gpu 2>&1 | tail -3 | tee gpurun_out/n2_allreduce.log
