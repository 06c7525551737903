#!/bin/bash
# fused CFT-block kernel: parity tests, timing, in-kernel timeline, one ncu capture (run under gpurun)
mkdir -p gpurun_out
if [ -z "$SKIP_TESTS" ]; then timeout 600 python -m pytest tests/test_block_gpu.py -x -q 2>&1 | tail -30 | tee gpurun_out/block_tests.log; fi
timeout 300 python scripts/time_block.py --dims 256 512 2>&1 | tee gpurun_out/block_time.jsonl
timeout 120 python scripts/trace_block.py --d 256 2>&1 | tee gpurun_out/block_trace_256.txt
timeout 120 python scripts/trace_block.py --d 512 2>&1 | tee gpurun_out/block_trace_512.txt

timeout 120 python scripts/trace_block.py --d 512 --batch 4 2>&1 | tee gpurun_out/block_trace_512_b4.txt
timeout 120 python scripts/time_block.py --dims 256 512 --batch 4 2>&1 | tee gpurun_out/block_time_b4.jsonl
