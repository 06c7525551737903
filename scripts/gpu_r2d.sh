#!/bin/bash
# late round 2: full validation first (every -m gpu test, smoke, bench), then Focus variants and fresh ncu captures of the
# kernels that changed late (run under gpurun)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 | tee gpurun_out/d_gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/d_smoke.log
timeout 600 python bench.py 2>&1 | tail -1 | tee gpurun_out/d_bench.json
timeout 200 python scripts/time_focus_variants.py 2>&1 | tail -6 | tee gpurun_out/d_focus_variants.jsonl
CASES="conv_c3_p4_3x3_256_pairs_n256 conv_c3_p5_3x3_512_tiles_4x4x8 conv_c3_p3_3x3_128_chain_1x1 unpool_p3_add2_add spp_maxpool_cascade upsample2x_p4 layernorm_d1024 pool_tokens_p3 detect_decode_p3" bash scripts/gpu_ncu_all.sh 2>&1 | tail -12
timeout 300 python scripts/ab_step.py --steps 30 --tag default 2>&1 | tail -1 | tee gpurun_out/d_ab_default.json
CFT_FUSED_BLOCK_MAX_D=512 timeout 300 python scripts/ab_step.py --steps 30 --tag fused512 2>&1 | tail -1 | tee gpurun_out/d_ab_fused512.json
