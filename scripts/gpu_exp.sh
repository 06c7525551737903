#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_model_gpu.py -q -m gpu --tb=line 2>&1 | tail -3
echo "== two streams"; python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'])"
echo "== one stream"; CFT_ONE_STREAM=1 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'])"
