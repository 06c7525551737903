#!/bin/bash
# Runs on the GPU box under gpurun. Each stage in its own process + timeout so one poisoned CUDA
# context (a trapped kernel) cannot hide the other results.  Logs land in gpurun_out/.
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" | tee -a gpurun_out/$name.log; tail -n ${TAILN:-25} gpurun_out/$name.log | cut -c1-600; }
TAILN=12 run kernels  python -m pytest tests/test_kernels_gpu.py -q -m gpu --tb=line
CFT_CONV_CTAS=2 TAILN=12 run conv_2cta python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "conv or gemm" --tb=line
CFT_CONV_CTAS=1 TAILN=12 run conv_1cta python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "conv or gemm" --tb=line
CFT_NO_ROW_REUSE=1 TAILN=12 run conv_norowreuse python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "conv or gemm" --tb=line
CFT_ATTENTION_SIMT=1 TAILN=6 run attn_simt python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "attention" --tb=line
TAILN=16 run model    python -m pytest tests/test_model_gpu.py -q -m gpu --tb=line -s
TAILN=4 run smoke    python __graft_entry__.py smoke
CFT_NO_BRES=1 TAILN=12 run conv_nobres python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "conv or gemm" --tb=line
CFT_STAGE8K=1 TAILN=12 run conv_stage8k python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "conv or gemm" --tb=line
TAILN=60 run step_1s  python scripts/trace_step.py 32 --one-stream
TAILN=8 run step_2s  python scripts/trace_step.py 32
CFT_STAGE8K=1 TAILN=60 run step_1s_stage8k  python scripts/trace_step.py 32 --one-stream
if [ -n "$CI_FULL" ]; then
TAILN=60 run layers   python scripts/profile_layers.py 32
run shapes   python scripts/prof_shapes.py --time
fi
run bench    python bench.py --steps 20 --warmup 5
