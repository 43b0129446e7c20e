#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/bench_dyna.py 2>&1 | grep -v Warn | tail -1 | tee gpurun_out/bench_dyna.log
timeout 900 python scripts/bench_algos.py 2>&1 | grep -v Warn > gpurun_out/bench_algos.log; tail -60 gpurun_out/bench_algos.log
