#!/bin/bash
# Runs on the GPU box under gpurun. Each stage in its own process + timeout so one poisoned CUDA
# context (a trapped kernel) cannot hide the other results.  Logs land in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
run() { name=$1; shift; echo "=== $name"; timeout 600 "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" | tee -a gpurun_out/$name.log; tail -n 25 gpurun_out/$name.log; }
run conv     python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "conv or gemm" -x --tb=short
run movers   python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "not conv and not gemm" --tb=short
run model    python -m pytest tests/test_model_gpu.py -q -m gpu --tb=short -s
run smoke    python __graft_entry__.py smoke
