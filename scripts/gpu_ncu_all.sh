#!/bin/bash
# per-kernel `ncu --set full` captures (one warm launch each) -> gpurun_out/ncu_r02/<case>.ncu-rep ; read them here with
# scripts/summarize_ncu.py.  Under gpurun, one GPU.   CASES="a b" bash scripts/gpu_ncu_all.sh
mkdir -p gpurun_out/ncu_r02
declare -A K=( [conv]=cft_conv_tcgen05 [focus]=cft_focus_tcgen05 [attention]=cft_attention_tcgen05 [layernorm]=layernorm_kernel
               [pool]=pool_tokens [unpool]=unpool_rows [spp]=maxpool_cascade [upsample2x]=upsample2x [detect]=detect_decode [gpt]=cft_gpt_block )
ALL=$(python - <<'PY'
import re,sys
src=open("scripts/prof_kernels.py").read()
print(" ".join(re.findall(r'^    "([a-z0-9_]+)": lambda', src, re.M)))
PY
)
for c in ${CASES:-$ALL}; do
  key=${c%%_*}
  kn=${K[$key]}
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$kn -s 2 -c 1 -f -o gpurun_out/ncu_r02/$c \
      python scripts/prof_kernels.py $c > gpurun_out/ncu_r02/$c.log 2>&1
  tail -1 gpurun_out/ncu_r02/$c.log
done
ls -la gpurun_out/ncu_r02 | head -40
