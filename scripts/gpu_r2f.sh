#!/bin/bash
# 256-thread attention core: validation + bench + ncu (run under gpurun)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/f_gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/f_smoke.log
timeout 600 python bench.py 2>&1 | tail -1 | tee gpurun_out/f_bench.json
CASES="attention_d1024 attention_d512" bash scripts/gpu_ncu_all.sh 2>&1 | tail -5
