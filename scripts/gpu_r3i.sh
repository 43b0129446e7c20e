#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gemm_gpu.py tests/test_qhead_envelope_gpu.py -q -k "chain" --maxfail 3 --timeout 120 2>&1 | tail -4 | tee gpurun_out/pytest_chain.log
if grep -q "failed\|error" gpurun_out/pytest_chain.log; then echo "chain test failed: skipping the rest"; exit 0; fi
for v in "MORL_X=0" "MORL_HEAD_REVERSE=0" "MORL_X=0" "MORL_HEAD_REVERSE=0"; do env $v MORL_SKIP_CPU_BASELINE=1 timeout 600 python bench.py --steps 300 --warmup 5 2>&1 | tail -1 | python -c "
import json,sys; l=json.loads(sys.stdin.read()); print('$v', round(l['value'],1), round(l['ms_per_step'],4), round(l['e2e']['value'],1), round(l['roofline']['us_per_launch'],1), round(l['roofline']['frac'],3))"; done | tee gpurun_out/bench_ab7.log
