#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 700 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 > gpurun_out/bench_under_ncu.log 2>&1
tail -2 gpurun_out/bench_under_ncu.log | cut -c1-300
