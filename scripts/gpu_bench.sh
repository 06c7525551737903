#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" | tee -a gpurun_out/$name.log; tail -n 30 gpurun_out/$name.log; }
run model    python -m pytest tests/test_model_gpu.py -q -m gpu --tb=short -k "golden"
run shapes   python scripts/prof_shapes.py --time
run bench_eager python bench.py --steps 10 --warmup 3 --no-graph --no-cpu-baseline
run bench    python bench.py --steps 20 --warmup 5
run launches ncu --metrics gpu__time_duration.sum --clock-control none -s 1300 -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline
run ncu_full ncu --set full --clock-control none --import-source on -k regex:cft_conv_tcgen05 -c 30 -o gpurun_out/prof_conv -f python scripts/prof_shapes.py
ls -la gpurun_out
