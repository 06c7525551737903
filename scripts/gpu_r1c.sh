#!/bin/bash
# Round-1 late session, GPU call 2: NMS tests after the sort rewrite, GPT/GELU kernel tests with the new default,
# NMS timing, the other BASELINE configs (sweep), the PyTorch-eager GPU baseline, the bench line.
mkdir -p gpurun_out
run() { name=$1; to=$2; shift 2; echo "=== $name"; t0=$SECONDS; timeout $to "$@" > gpurun_out/$name.log 2>&1; echo "exit $? ($((SECONDS-t0)) s)" | tee -a gpurun_out/$name.log; tail -n ${TAILN:-12} gpurun_out/$name.log | cut -c1-${CUT:-400}; }
TAILN=6 run nms 300 python -m pytest tests/test_nms_gpu.py -q -m gpu --tb=short -x
TAILN=6 run suite 600 python -m pytest tests -q -m gpu --tb=line --deselect tests/test_nms_gpu.py
TAILN=3 CUT=900 run ab_base 200 python scripts/ab_step.py --tag v14_default --nms
TAILN=20 run sweep 600 python scripts/sweep_configs.py --out gpurun_out/sweep.jsonl
TAILN=5 run eager 400 python scripts/eager_baseline.py
TAILN=3 CUT=6000 run bench 600 python bench.py --steps 20 --warmup 5
