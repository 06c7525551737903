#!/bin/bash
mkdir -p gpurun_out
for s in "$@"; do
  timeout 300 python scripts/trace_conv.py $s 32 > gpurun_out/trace_$s.log 2>&1; echo "$s exit $?"
  grep -v 'CTA [0-9]' gpurun_out/trace_$s.log | head -12
  grep 'CTA 0:' gpurun_out/trace_$s.log | head -1 | cut -c1-300
done
