#!/bin/bash
mkdir -p gpurun_out
echo "== bench single acc"; MORL_GEMM_SPLIT_ACC=0 timeout 900 python bench.py --steps 100 --warmup 5 2>&1 | tail -1 > gpurun_out/bench_single.log; python -c "
import json; l=json.loads(open('gpurun_out/bench_single.log').read()); print('single', l['value'], l['ms_per_step'], l['e2e']['value'], l['roofline']['us_per_launch'])"
echo "== bench split acc"; timeout 900 python bench.py --steps 100 --warmup 5 2>&1 | tail -1 > gpurun_out/bench.log; python -c "
import json; l=json.loads(open('gpurun_out/bench.log').read()); print('split', l['value'], l['ms_per_step'], l['e2e']['value'], l['roofline']['us_per_launch'])"
echo "== ncu launch list (split)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 500 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 > gpurun_out/bench_under_ncu.log 2>&1
echo "== ncu launch list (single)"
MORL_GEMM_SPLIT_ACC=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 500 --csv --log-file gpurun_out/launches_single.csv python bench.py --steps 3 --warmup 3 > gpurun_out/bench_under_ncu_single.log 2>&1
ls -la gpurun_out/*.csv
