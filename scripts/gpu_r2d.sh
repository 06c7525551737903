#!/bin/bash
# refresh the ncu captures of the kernels that changed late in round 2 + Focus variants + a bench line (run under gpurun)
mkdir -p gpurun_out
timeout 200 python scripts/time_focus_variants.py 2>&1 | tail -6 | tee gpurun_out/d_focus_variants.jsonl
CASES="conv_c3_p4_3x3_256_pairs_n256 conv_c3_p5_3x3_512_tiles_4x4x8 conv_c3_p3_3x3_128_chain_1x1 unpool_p3_add2_add spp_maxpool_cascade upsample2x_p4 layernorm_d1024 pool_tokens_p3 detect_decode_p3" bash scripts/gpu_ncu_all.sh 2>&1 | tail -12
timeout 600 python bench.py 2>&1 | tail -1 | tee gpurun_out/d_bench.json
