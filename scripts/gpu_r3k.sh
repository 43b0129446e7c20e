#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_dyna_gpu.py -q --maxfail 4 --timeout 200 2>&1 | tail -12 | tee gpurun_out/pytest_dyna.log
timeout 300 python scripts/bench_dyna.py 2>&1 | grep -v Warn | tail -1 | tee gpurun_out/bench_dyna_graph.log
MORL_DYNA_FIT_GRAPH=0 timeout 300 python scripts/bench_dyna.py 2>&1 | grep -v Warn | tail -1 | tee gpurun_out/bench_dyna_eager.log
