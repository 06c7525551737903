#!/bin/bash
# Round-1 late session, GPU call 1: full GPU suite (default switches), the suite again with the three new switches on,
# then A/B step timings of each switch.  Every stage has its own timeout; logs land in gpurun_out/.
mkdir -p gpurun_out
run() { name=$1; to=$2; shift 2; echo "=== $name"; t0=$SECONDS; timeout $to "$@" > gpurun_out/$name.log 2>&1; echo "exit $? ($((SECONDS-t0)) s)" | tee -a gpurun_out/$name.log; tail -n ${TAILN:-12} gpurun_out/$name.log | cut -c1-400; }
TAILN=14 run nms 300 python -m pytest tests/test_nms_gpu.py -q -m gpu --tb=short -x
TAILN=10 run suite 600 python -m pytest tests -q -m gpu --tb=line --deselect tests/test_nms_gpu.py
export CFT_PDL_ALL=1 CFT_GELU_FAST=1 CFT_ONE_TEAM=1
TAILN=10 run suite_flags 600 python -m pytest tests -q -m gpu --tb=line
unset CFT_PDL_ALL CFT_GELU_FAST CFT_ONE_TEAM
TAILN=3 run ab_base 200 python scripts/ab_step.py --tag base --nms
CFT_PDL_ALL=1 TAILN=3 run ab_pdl 200 python scripts/ab_step.py --tag pdl_all
CFT_GELU_FAST=1 TAILN=3 run ab_gelu 200 python scripts/ab_step.py --tag gelu_fast
CFT_ONE_TEAM=1 TAILN=3 run ab_team 200 python scripts/ab_step.py --tag one_team
CFT_PDL_ALL=1 CFT_GELU_FAST=1 CFT_ONE_TEAM=1 TAILN=3 run ab_all 200 python scripts/ab_step.py --tag all
CFT_PDL_ALL=1 CFT_GELU_FAST=1 CFT_ONE_TEAM=1 CFT_ONE_STREAM=1 TAILN=3 run ab_all_1s 200 python scripts/ab_step.py --tag all_one_stream
CFT_ONE_STREAM=1 TAILN=3 run ab_base_1s 200 python scripts/ab_step.py --tag base_one_stream
