#!/bin/bash
mkdir -p gpurun_out
S="c3_p3_1x1_128 c3_p2_1x1_64 c3_p3_3x3_128 c3_p4_3x3_256 c3_p5_3x3_512 c3_p2_3x3_64 focus_16_64 gpt_p5_down_4096_1024 gpt_p5_up_1024_4096"
echo "== base";            python scripts/prof_shapes.py --time $S
echo "== cbufs=2";         CFT_STAGE_BUFS=2 python scripts/prof_shapes.py --time $S
echo "== silu tanh";       CFT_SILU_TANH=1 python scripts/prof_shapes.py --time $S
echo "== model parity (exp2/rcp SiLU)"; python -m pytest tests/test_model_gpu.py -q -m gpu -s -k "golden or l_640" 2>&1 | grep -E "raw0|passed|failed"
echo "== model parity (tanh SiLU)"; CFT_SILU_TANH=1 python -m pytest tests/test_model_gpu.py -q -m gpu -s -k "golden or l_640" 2>&1 | grep -E "raw0|passed|failed"
python scripts/profile_layers.py 32 > gpurun_out/layers.log 2>&1; head -30 gpurun_out/layers.log
python bench.py --steps 20 --warmup 5 > gpurun_out/bench.log 2>&1; tail -1 gpurun_out/bench.log | cut -c1-400
