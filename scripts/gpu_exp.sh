#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_kernels_gpu.py -q -m gpu --tb=line 2>&1 | tail -3
python -m pytest tests/test_model_gpu.py -q -m gpu --tb=line 2>&1 | tail -3
echo "== PDL on";  python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['kernel_ms_per_step'])"
echo "== PDL off"; CFT_NO_PDL=1 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['kernel_ms_per_step'])"
echo "== e2e debug"; python scripts/e2e_debug.py 2>&1 | grep -v Warn
