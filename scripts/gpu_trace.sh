#!/bin/bash
mkdir -p gpurun_out
for s in "$@"; do
  timeout 300 python scripts/trace_conv.py $s > gpurun_out/trace_$s.log 2>&1; echo "$s exit $?"
  grep -v 'CTA [0-9]' gpurun_out/trace_$s.log
done
