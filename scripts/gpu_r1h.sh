#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; to=$2; shift 2; echo "=== $name"; t0=$SECONDS; timeout $to "$@" > gpurun_out/$name.log 2>&1; echo "exit $? ($((SECONDS-t0)) s)" | tee -a gpurun_out/$name.log; tail -n ${TAILN:-12} gpurun_out/$name.log | cut -c1-${CUT:-400}; }
TAILN=2 CUT=1400 run bench4 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras
TAILN=2 CUT=1400 run bench3 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras --slots 3
TAILN=2 CUT=1400 run bench2 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras --slots 2
