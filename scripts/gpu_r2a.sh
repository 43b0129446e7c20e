#!/bin/bash
# round-2 first call: new whole-update goldens + full GPU suite + bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
lscpu | head -20 > gpurun_out/cpu.txt; nproc >> gpurun_out/cpu.txt
echo "== new tests first"; timeout 900 python -m pytest tests/test_envelope_update_golden_gpu.py tests/test_gemm_gpu.py tests/test_envelope_update_gpu.py -m gpu -q --timeout 600 2>&1 | tail -60 | tee gpurun_out/pytest_new.log
echo "== pytest -m gpu" ; timeout 1500 python -m pytest tests -m gpu -q --maxfail 12 --timeout 900 --durations=8 2>&1 | tail -40 | tee gpurun_out/pytest_gpu.log
echo "== smoke" ; timeout 300 python __graft_entry__.py smoke 2>&1 | grep -v Warning | tail -5 | tee gpurun_out/smoke.log
echo "== bench" ; timeout 900 python bench.py --steps 200 --warmup 5 2>&1 | tail -1 | tee gpurun_out/bench.log
