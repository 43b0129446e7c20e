#!/bin/bash
mkdir -p gpurun_out
for m in f16x2 bf16x3 notc; do timeout 300 python scripts/hv_debug.py $m 1500 2>&1 | grep "step" ; done | tee gpurun_out/hv_debug.log
echo "== sumtree + per"; timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q --timeout 500 -k "sumtree or device_per" 2>&1 | tail -12 | tee gpurun_out/pytest_sumtree.log
echo "== population"; timeout 600 python -m pytest tests/test_algos_update_gpu.py -m gpu -q --timeout 500 -k "population" 2>&1 | tail -12 | tee gpurun_out/pytest_pop.log
echo "== envelope"; timeout 900 python -m pytest tests/test_envelope_update_gpu.py tests/test_envelope_update_golden_gpu.py -m gpu -q --timeout 500 2>&1 | tail -12 | tee gpurun_out/pytest_env.log
echo "== bench"; timeout 900 python bench.py --steps 100 --warmup 5 2>&1 | tail -1 > gpurun_out/bench.log; python -c "
import json; l=json.loads(open('gpurun_out/bench.log').read()); print('bench', l['value'], l['ms_per_step'], l['e2e']['value'], l['roofline']['us_per_launch'], l['gpu_launches'], l['config']['ms_eval_round_rank0'])"
MORL_GEMM_SPLIT_ACC=0 timeout 900 python bench.py --steps 100 --warmup 5 2>&1 | tail -1 > gpurun_out/bench_single.log; python -c "
import json; l=json.loads(open('gpurun_out/bench_single.log').read()); print('bench single', l['value'], l['ms_per_step'], l['e2e']['value'], l['roofline']['us_per_launch'])"
