#!/bin/bash
# same-box A/B of one environment switch of the library (run under gpurun):  SWITCH="CFT_NO_BATCH_TILES=1" bash scripts/gpu_ab.sh
# -> three ab_step.py processes (default, switch, default again: the repeat shows the run-to-run spread of the lease)
mkdir -p gpurun_out
timeout 300 python scripts/ab_step.py --steps 30 --tag default 2>&1 | tail -1 | tee gpurun_out/ab_default.json
env ${SWITCH:?set SWITCH=NAME=VALUE} timeout 300 python scripts/ab_step.py --steps 30 --tag "$SWITCH" 2>&1 | tail -1 | tee gpurun_out/ab_switch.json
timeout 300 python scripts/ab_step.py --steps 30 --tag default_again 2>&1 | tail -1 | tee gpurun_out/ab_default_again.json
