#!/bin/bash
# N-GPU job (gpurun --gpus N): BASELINE config 5 (yolov5x-x3 @640) strong-scaling points -- global batch 8 / 32 / 128 split
# over the N GPUs -- and, with CONFIG4=1, the config-4 train step.   N=<gpus> bash scripts/gpu_scale.sh
mkdir -p gpurun_out
N=${N:-1}
run() {  # args: per-GPU batch
  if [ "$N" = "1" ]; then python bench.py --gpus 1 "$@"; else
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N "$@"; fi
}
for G in ${GLOBALS:-8 32 128}; do
  b=$((G / N)); [ $b -lt 1 ] && continue
  timeout 600 bash -c "$(declare -f run); N=$N run --cfg yolov5x_fusion_transformerx3_FLIR_aligned --batch $b --steps 30 --warmup 5 --no-extras --no-eager-baseline --no-cpu-baseline" 2>&1 | grep '^{' | python -c "
import json,sys
for ln in sys.stdin:
    d=json.loads(ln); print(json.dumps({'config':5,'n_gpus':d['n_gpus'],'global_batch':d['config']['global_batch'],'per_gpu_batch':$b,'pairs_per_s':round(d['value'],1),'ms_per_step':round(d['ms_per_step'],3),'e2e_pairs_per_s':round(d['e2e']['value'],1),'whole_forward_tensor_frac':round(d['roofline']['whole_forward_tensor_frac'],4) if d.get('roofline') else None}))
" | tee -a gpurun_out/config5_scaling_n$N.jsonl
done
if [ -n "$CONFIG4" ]; then
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 scripts/train_step.py --steps 5 --warmup 2 2>&1 | grep '^{' | tee gpurun_out/config4_n$N.json
fi
