#!/bin/bash
mkdir -p gpurun_out
for s in c3_p3_1x1_128 c3_p3_3x3_128 focus_16_64; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:cft_conv_tcgen05 -s 1 -c 1 -f -o gpurun_out/prof_$s python scripts/prof_shapes.py $s > gpurun_out/ncu_$s.log 2>&1
  echo "$s exit $?"
done
ls -la gpurun_out/*.ncu-rep
