#!/bin/bash
# full GPU regression: every -m gpu test, smoke(), bench line (run under gpurun)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 2>&1 | tail -3 | tee gpurun_out/bench_n1.json
