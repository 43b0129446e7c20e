#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_algos_update_gpu.py -q --timeout 300 2>&1 | grep -v "^E   \(  \|$\)" | tail -60 | tee gpurun_out/algos.log
timeout 600 python bench.py --steps 100 --warmup 5 2>&1 | tail -1 | tee gpurun_out/bench_quick.log | cut -c1-300
