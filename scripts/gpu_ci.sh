#!/bin/bash
# Runs on the GPU box under gpurun. Each stage in its own process + timeout so one poisoned CUDA
# context (a trapped kernel) cannot hide the other results.  Logs land in gpurun_out/.
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" | tee -a gpurun_out/$name.log; tail -n ${TAILN:-25} gpurun_out/$name.log | cut -c1-600; }
run kernels  python -m pytest tests/test_kernels_gpu.py -q -m gpu --tb=short
CFT_CONV_CTAS=2 run conv_2cta python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "conv or gemm" --tb=short
CFT_CONV_CTAS=1 run conv_1cta python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "conv or gemm" --tb=short
CFT_ATTENTION_SIMT=1 run attn_simt python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "attention" --tb=short
run model    python -m pytest tests/test_model_gpu.py -q -m gpu --tb=short -s
run smoke    python __graft_entry__.py smoke
TAILN=60 run layers   python scripts/profile_layers.py 32
run shapes   python scripts/prof_shapes.py --time
run bench    python bench.py --steps 20 --warmup 5
