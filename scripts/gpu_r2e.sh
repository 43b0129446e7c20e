#!/bin/bash
mkdir -p gpurun_out
echo "== pytest -m gpu" ; timeout 1800 python -m pytest tests -m gpu -q --maxfail 15 --timeout 900 --durations=6 2>&1 | tail -30 | tee gpurun_out/pytest_gpu.log
echo "== bench default (fwd split / bwd single)"; timeout 900 python bench.py --steps 100 --warmup 5 2>&1 | tail -1 > gpurun_out/bench.log; python -c "
import json; l=json.loads(open('gpurun_out/bench.log').read()); print('default', l['value'], l['ms_per_step'], l['e2e']['value'], l['roofline']['us_per_launch'], l['gpu_launches'])"
echo "== bench single acc"; MORL_GEMM_SPLIT_ACC=0 timeout 900 python bench.py --steps 100 --warmup 5 2>&1 | tail -1 > gpurun_out/bench_single.log; python -c "
import json; l=json.loads(open('gpurun_out/bench_single.log').read()); print('single', l['value'], l['ms_per_step'], l['e2e']['value'], l['roofline']['us_per_launch'])"
echo "== memcheck"; timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python scripts/sanitize_all.py 2>&1 | grep -v Warning | tail -25 | tee gpurun_out/sanitize_memcheck.log
