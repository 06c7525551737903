#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "focus" --tb=short > gpurun_out/focus_tests.log 2>&1; echo "focus tests exit $?"; tail -30 gpurun_out/focus_tests.log | cut -c1-300
