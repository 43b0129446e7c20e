#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemm_gpu.py tests/test_qhead_envelope_gpu.py -q --maxfail 3 --timeout 120 2>&1 | tail -4 | tee gpurun_out/pytest_mn.log
if grep -q "failed\|error" gpurun_out/pytest_mn.log; then echo "tests failed: skipping the rest"; exit 0; fi
timeout 900 python -m pytest tests/test_envelope_update_golden_gpu.py tests/test_envelope_update_gpu.py -q --maxfail 6 --timeout 600 2>&1 | tail -3 | tee gpurun_out/pytest_mn_update.log
for v in "MORL_MN_MULTICAST=0" "MORL_MN_MULTICAST=1" "MORL_MN_MULTICAST=0" "MORL_MN_MULTICAST=1"; do env $v MORL_SKIP_CPU_BASELINE=1 timeout 600 python bench.py --steps 300 --warmup 5 2>&1 | tail -1 | python -c "
import json,sys; l=json.loads(sys.stdin.read()); print('$v', round(l['value'],1), round(l['ms_per_step'],4), round(l['e2e']['value'],1))"; done | tee gpurun_out/bench_ab9.log
