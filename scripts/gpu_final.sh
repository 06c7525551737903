#!/bin/bash
# Round-end check on one GPU: what the driver runs (GPU suite, smoke, bench both arms).
mkdir -p gpurun_out
run() { name=$1; to=$2; shift 2; echo "=== $name"; t0=$SECONDS; timeout $to "$@" > gpurun_out/$name.log 2>&1; echo "exit $? ($((SECONDS-t0)) s)" | tee -a gpurun_out/$name.log; tail -n ${TAILN:-12} gpurun_out/$name.log | cut -c1-${CUT:-400}; }
TAILN=5 run suite 900 python -m pytest tests -x -q -m gpu
TAILN=3 run smoke 300 python -c "import __graft_entry__ as g; g.smoke()"
TAILN=2 CUT=1500 run bench 600 python bench.py
