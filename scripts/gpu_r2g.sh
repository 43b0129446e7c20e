#!/bin/bash
mkdir -p gpurun_out
echo "== single tile"; timeout 600 python -m pytest tests/test_envelope_update_gpu.py -m gpu -q -x --timeout 500 -k "single_tile" 2>&1 | tail -30 | tee gpurun_out/pytest_single_tile.log
echo "== hv parity"; timeout 600 python -m pytest tests/test_hv_parity_gpu.py -m gpu -q -x -s --timeout 500 -k "one_percent" 2>&1 | grep "seed\|mean\|passed\|failed" | tee gpurun_out/pytest_hv.log
echo "== population + sumtree"; timeout 600 python -m pytest tests/test_algos_update_gpu.py tests/test_kernels_gpu.py -m gpu -q --timeout 500 -k "population or sumtree or device_per" 2>&1 | tail -12 | tee gpurun_out/pytest_pop.log
echo "== device per envelope + goldens"; timeout 900 python -m pytest tests/test_envelope_update_gpu.py tests/test_envelope_update_golden_gpu.py -m gpu -q --timeout 500 2>&1 | tail -12 | tee gpurun_out/pytest_env.log
echo "== bench"; timeout 900 python bench.py --steps 100 --warmup 5 2>&1 | tail -1 > gpurun_out/bench.log; python -c "
import json; l=json.loads(open('gpurun_out/bench.log').read()); print('bench', l['value'], l['ms_per_step'], l['e2e']['value'], l['roofline']['us_per_launch'], l['gpu_launches'], l['config']['ms_eval_round_rank0'])"
