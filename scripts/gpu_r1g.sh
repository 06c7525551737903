#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; to=$2; shift 2; echo "=== $name"; t0=$SECONDS; timeout $to "$@" > gpurun_out/$name.log 2>&1; echo "exit $? ($((SECONDS-t0)) s)" | tee -a gpurun_out/$name.log; tail -n ${TAILN:-12} gpurun_out/$name.log | cut -c1-${CUT:-400}; }
TAILN=6 run engine 400 python -m pytest tests/test_model_gpu.py tests/test_nms_gpu.py -q -m gpu --tb=short -k "engine"
TAILN=3 CUT=8000 run bench 600 python bench.py --steps 20 --warmup 5
TAILN=3 CUT=3000 run bench_ref 600 python bench.py --impl reference --steps 2 --warmup 1
