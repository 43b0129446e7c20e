#!/bin/bash
mkdir -p gpurun_out
echo "== pytest -m gpu" ; timeout 2400 python -m pytest tests -m gpu -q --maxfail 20 --timeout 900 --durations=6 2>&1 | tail -40 | tee gpurun_out/pytest_gpu.log
echo "== bench"; timeout 900 python bench.py --steps 100 --warmup 5 2>&1 | tail -1 > gpurun_out/bench.log; python -c "
import json; l=json.loads(open('gpurun_out/bench.log').read()); print('bench', l['value'], l['ms_per_step'], l['e2e']['value'], l['roofline']['us_per_launch'], l['gpu_launches'], l['config']['ms_eval_round_rank0'])"
MORL_GEMM_SPLIT_ACC=0 timeout 900 python bench.py --steps 100 --warmup 5 2>&1 | tail -1 > gpurun_out/bench_single.log; python -c "
import json; l=json.loads(open('gpurun_out/bench_single.log').read()); print('bench single', l['value'], l['ms_per_step'], l['e2e']['value'], l['roofline']['us_per_launch'])"
echo "== golden diag"; timeout 600 python scripts/golden_diag.py north_star 2>&1 | grep -v Warning | grep "tc=True graph=True" -A 11 | tee gpurun_out/golden_diag_f16x2.log
