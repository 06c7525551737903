#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" | tee -a gpurun_out/$name.log; tail -n ${TAILN:-25} gpurun_out/$name.log | cut -c1-600; }
run kernels  python -m pytest tests/test_kernels_gpu.py -q -m gpu --tb=short
CFT_CONV_CTAS=2 run conv_2cta python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "conv or gemm" --tb=short
run model    python -m pytest tests/test_model_gpu.py -q -m gpu --tb=short
TAILN=60 run layers   python scripts/profile_layers.py 32
run shapes   python scripts/prof_shapes.py --time
run bench    python bench.py --steps 20 --warmup 5
for s in focus_16_64 c3_p2_3x3_64 c3_p3_1x1_128; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:cft_conv_tcgen05 -s 1 -c 1 -f -o gpurun_out/prof2_$s python scripts/prof_shapes.py $s > gpurun_out/ncu2_$s.log 2>&1
done
