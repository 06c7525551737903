#!/bin/bash
# GPU call 4: full suite after the NMS round-loop rewrite and the mixed-chunk attention; timings; sweep; bench line.
mkdir -p gpurun_out
run() { name=$1; to=$2; shift 2; echo "=== $name"; t0=$SECONDS; timeout $to "$@" > gpurun_out/$name.log 2>&1; echo "exit $? ($((SECONDS-t0)) s)" | tee -a gpurun_out/$name.log; tail -n ${TAILN:-12} gpurun_out/$name.log | cut -c1-${CUT:-400}; }
TAILN=6 run nms 300 python -m pytest tests/test_nms_gpu.py -q -m gpu --tb=short -x
TAILN=6 run suite 600 python -m pytest tests -q -m gpu --tb=line --deselect tests/test_nms_gpu.py
TAILN=3 run smoke 300 python __graft_entry__.py smoke
TAILN=3 CUT=900 run ab_nms 200 python scripts/ab_step.py --tag v15 --nms --steps 20
TAILN=20 run sweep 600 python scripts/sweep_configs.py --out gpurun_out/sweep.jsonl
TAILN=3 CUT=8000 run bench 600 python bench.py --steps 20 --warmup 5
