#!/bin/bash
# bench (plain) + ncu launch list of a short bench run
mkdir -p gpurun_out
timeout 600 python bench.py --steps 200 --warmup 5 2>&1 | tail -1 | tee gpurun_out/bench.log | python -c "import sys,json; l=json.loads(sys.stdin.read()); print(l['value'], l['ms_per_step'], l['e2e']['value'], l['gpu_launches'])"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 700 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 > gpurun_out/bench_under_ncu.log 2>&1
tail -2 gpurun_out/bench_under_ncu.log | cut -c1-200
nvidia-smi --query-gpu=name,clocks.sm,temperature.gpu --format=csv
