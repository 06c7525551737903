#!/bin/bash
# Launch list of ONE eager forward step (all kernels) with duration + DRAM bytes, for profiles/.
mkdir -p gpurun_out
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
  --clock-control none --csv --log-file gpurun_out/step_launches.csv \
  python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --ncu-range > gpurun_out/ncu_step.log 2>&1
echo "exit $?"; wc -l gpurun_out/step_launches.csv
