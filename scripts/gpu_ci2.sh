#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" | tee -a gpurun_out/$name.log; tail -n ${TAILN:-25} gpurun_out/$name.log; }
CFT_CONV_CTAS=2 run conv_2cta python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "conv or gemm" --tb=short
CFT_CONV_CTAS=2 run model_2cta python -m pytest tests/test_model_gpu.py -q -m gpu --tb=short -k "golden or l_640 or batch32"
CFT_CONV_CTAS=2 TAILN=20 run shapes_2cta python scripts/prof_shapes.py --time
CFT_CONV_CTAS=1 TAILN=20 run shapes_1cta python scripts/prof_shapes.py --time
TAILN=50 run layers   python scripts/profile_layers.py 32
run bench    python bench.py --steps 20 --warmup 5
