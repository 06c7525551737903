#!/bin/bash
# un-pool at 2 CTAs / SM, LayerNorm templated on the row length: validation + bench + ncu of both (run under gpurun)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/e_gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/e_smoke.log
timeout 600 python bench.py 2>&1 | tail -1 | tee gpurun_out/e_bench.json
CASES="unpool_p3_add2_add layernorm_d1024" bash scripts/gpu_ncu_all.sh 2>&1 | tail -5
timeout 300 python scripts/trace_step.py 32 > gpurun_out/e_timeline_v6.txt 2>&1; head -2 gpurun_out/e_timeline_v6.txt
