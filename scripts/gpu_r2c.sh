#!/bin/bash
# batch-spanning tiles + un-pool prefetch: parity, A/B, per-shape timelines (run under gpurun)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py tests/test_block_gpu.py -x -q -m gpu 2>&1 | tail -8 | tee gpurun_out/c_tests.log
timeout 300 python scripts/ab_step.py --steps 30 --tag tiles3d 2>&1 | tail -1 | tee gpurun_out/c_ab_tiles3d.json
CFT_NO_BATCH_TILES=1 timeout 300 python scripts/ab_step.py --steps 30 --tag tiles2d 2>&1 | tail -1 | tee gpurun_out/c_ab_tiles2d.json
timeout 300 python scripts/ab_step.py --steps 30 --tag tiles3d_b 2>&1 | tail -1 | tee gpurun_out/c_ab_tiles3d_b.json
timeout 300 python scripts/trace_step.py 32 > gpurun_out/c_timeline_v5.txt 2>&1; head -3 gpurun_out/c_timeline_v5.txt
CFT_CONV_CTAS=1 timeout 300 python scripts/trace_step.py 32 > gpurun_out/c_timeline_v5_ctas1.txt 2>&1; head -3 gpurun_out/c_timeline_v5_ctas1.txt
timeout 300 python scripts/sweep_configs.py --only config3 --steps 10 --warmup 3 2>&1 | tail -1 | tee gpurun_out/c_c3.json
timeout 300 python scripts/sweep_configs.py --only config5_x_640_b32 --steps 10 --warmup 3 2>&1 | tail -1 | tee gpurun_out/c_c5_b32.json
