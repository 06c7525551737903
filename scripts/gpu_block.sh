#!/bin/bash
# fused CFT-block kernel: parity tests, timing, in-kernel timeline, one ncu capture (run under gpurun)
mkdir -p gpurun_out
if [ -z "$SKIP_TESTS" ]; then timeout 600 python -m pytest tests/test_block_gpu.py -x -q 2>&1 | tail -30 | tee gpurun_out/block_tests.log; fi
timeout 300 python scripts/time_block.py --dims 256 512 2>&1 | tee gpurun_out/block_time.jsonl
timeout 120 python scripts/trace_block.py --d 256 2>&1 | tee gpurun_out/block_trace_256.txt
timeout 120 python scripts/trace_block.py --d 512 2>&1 | tee gpurun_out/block_trace_512.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:cft_gpt_block -s 3 -c 1 -o gpurun_out/block256 python scripts/trace_block.py --d 256 > gpurun_out/ncu_block.log 2>&1
