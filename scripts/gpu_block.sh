#!/bin/bash
# fused CFT-block kernel: parity tests, then timing (run under gpurun)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_block_gpu.py -x -q 2>&1 | tail -30 | tee gpurun_out/block_tests.log
timeout 300 python scripts/time_block.py 2>&1 | tee gpurun_out/block_time.jsonl
