#!/bin/bash
# GPU call 3: NMS tests (args as __grid_constant__), ncu evidence for v14: launch list of one step, --set full captures
# of the NMS, attention and fused-Focus kernels.
mkdir -p gpurun_out
run() { name=$1; to=$2; shift 2; echo "=== $name"; t0=$SECONDS; timeout $to "$@" > gpurun_out/$name.log 2>&1; echo "exit $? ($((SECONDS-t0)) s)" | tee -a gpurun_out/$name.log; tail -n ${TAILN:-12} gpurun_out/$name.log | cut -c1-${CUT:-400}; }
TAILN=4 run nms 300 python -m pytest tests/test_nms_gpu.py -q -m gpu --tb=short -x
TAILN=8 run attn 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu --tb=short -k attention
TAILN=6 run model_x 400 python -m pytest tests/test_model_gpu.py -q -m gpu --tb=short -k "x_flir or yolov5x or golden"
TAILN=3 CUT=600 run sweep_x 300 python scripts/sweep_configs.py --only config5_x_640_b32
TAILN=3 CUT=900 run ab_nms 200 python scripts/ab_step.py --tag v14b --nms --steps 10
TAILN=3 run ncu_nms 300 ncu --set full --clock-control none --import-source on -k regex:nms_kernel -s 1 -c 1 -f -o gpurun_out/prof_nms python scripts/prof_nms.py
TAILN=3 run ncu_step 600 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/step_launches_v14.csv python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-extras --ncu-range
TAILN=3 run ncu_attn 300 ncu --set full --clock-control none --import-source on -k regex:cft_attention_tcgen05 -s 40 -c 1 -f -o gpurun_out/prof_attn python scripts/ab_step.py --tag ncu --steps 1
ls -la gpurun_out | tail -12
