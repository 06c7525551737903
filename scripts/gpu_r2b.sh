#!/bin/bash
# round 2, late: full GPU suite with the switch arms, then the config 3 / config 5 A/Bs of the row-reuse guard and the
# 32-wide K tail (run under gpurun)
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -x -q -m gpu 2>&1 | tail -25 | tee gpurun_out/gpu_tests.log
timeout 300 python scripts/sweep_configs.py --only config3 --steps 10 --warmup 3 2>&1 | tail -1 | tee gpurun_out/c3_guard.json
CFT_NO_ROW_REUSE=1 timeout 300 python scripts/sweep_configs.py --only config3 --steps 10 --warmup 3 2>&1 | tail -1 | tee gpurun_out/c3_norowreuse.json
timeout 300 python scripts/sweep_configs.py --only config5_x_640_b32 --steps 10 --warmup 3 2>&1 | tail -1 | tee gpurun_out/c5_b32_default.json
CFT_KTAIL32=1 timeout 300 python scripts/sweep_configs.py --only config5_x_640_b32 --steps 10 --warmup 3 2>&1 | tail -1 | tee gpurun_out/c5_b32_ktail32.json
timeout 300 python scripts/ab_step.py --steps 30 --tag default 2>&1 | tail -1 | tee gpurun_out/ab_default.json
