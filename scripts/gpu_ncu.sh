#!/bin/bash
mkdir -p gpurun_out
for s in c3_p3_1x1_128 c3_p4_3x3_256; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:cft_conv_tcgen05 -s 1 -c 1 -f -o gpurun_out/prof3_$s python scripts/prof_shapes.py $s > gpurun_out/ncu3_$s.log 2>&1
  echo "$s exit $?"
done
bash scripts/gpu_ncu_step.sh
