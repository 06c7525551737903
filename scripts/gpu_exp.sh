#!/bin/bash
mkdir -p gpurun_out
S="c3_p3_1x1_128 c3_p2_1x1_64 c3_p3_3x3_128 c3_p4_3x3_256 c3_p2_3x3_64 focus_16_64 gpt_p5_down_4096_1024"
echo "== base (cbufs=1)";            python scripts/prof_shapes.py --time $S
echo "== cbufs=2";                   CFT_STAGE_BUFS=2 python scripts/prof_shapes.py --time $S
echo "== no tap grouping";           CFT_NO_TAP_GROUPING=1 python scripts/prof_shapes.py --time $S
echo "== skip store";                CFT_DEBUG_SKIP_STORE=1 python scripts/prof_shapes.py --time $S
echo "== no act";                    python scripts/prof_shapes.py --time --noact $S
echo "== no act, skip store";        CFT_DEBUG_SKIP_STORE=1 python scripts/prof_shapes.py --time --noact $S
