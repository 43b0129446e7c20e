#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gemm_gpu.py tests/test_qhead_envelope_gpu.py -q -k "chain" --maxfail 3 --timeout 120 2>&1 | tail -15 | tee gpurun_out/pytest_chain.log
if grep -q "failed\|error" gpurun_out/pytest_chain.log; then echo "chain test failed: skipping the rest"; exit 0; fi
MORL_GEMM_CHAIN_BWD=1 timeout 900 python -m pytest tests/test_envelope_update_golden_gpu.py tests/test_envelope_update_gpu.py tests/test_qhead_envelope_gpu.py tests/test_hv_parity_gpu.py -q --maxfail 6 --timeout 600 2>&1 | tail -3 | tee gpurun_out/pytest_chain_update.log
for v in "MORL_GEMM_CHAIN_BWD=0" "MORL_GEMM_CHAIN_BWD=1" "MORL_GEMM_CHAIN_BWD=0" "MORL_GEMM_CHAIN_BWD=1"; do env $v MORL_SKIP_CPU_BASELINE=1 timeout 600 python bench.py --steps 300 --warmup 5 2>&1 | tail -1 | python -c "
import json,sys; l=json.loads(sys.stdin.read()); print('$v', round(l['value'],1), round(l['ms_per_step'],4), round(l['e2e']['value'],1))"; done | tee gpurun_out/bench_ab5.log
MORL_GEMM_CHAIN_BWD=1 timeout 300 python scripts/kernel_timeline.py 8 2>&1 | grep -v Warn > gpurun_out/kernel_timeline4.log; head -14 gpurun_out/kernel_timeline4.log
