#!/bin/bash
# 2-GPU job: NCCL check of the gradient all-reduce, then the config-4 train step (run under gpurun --gpus 2)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_allreduce_gpu.py -x -q -s 2>&1 | tail -8 | tee gpurun_out/allreduce_gpu.log
N=${N:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 scripts/train_step.py --steps ${STEPS:-5} --warmup 2 2>&1 | grep -v Warning | tail -4 | tee gpurun_out/config4_n$N.json
