#!/bin/bash
mkdir -p gpurun_out
echo "== golden diag"; timeout 600 python scripts/golden_diag.py 2>&1 | grep -v Warning | tee gpurun_out/golden_diag.log
echo "== bench --impl reference"; MORL_CPU_BUDGET_S=200 timeout 900 python bench.py --impl reference --steps 20 --warmup 5 2>&1 | tail -1 | tee gpurun_out/bench_reference.log
