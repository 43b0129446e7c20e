#!/bin/bash
mkdir -p gpurun_out
for v in "MORL_TWO_STREAMS=0" "MORL_TWO_STREAMS=1" "MORL_THREE_STREAMS=1" "MORL_TWO_STREAMS=0" "MORL_TWO_STREAMS=1" "MORL_THREE_STREAMS=1"; do env $v MORL_SKIP_CPU_BASELINE=1 timeout 600 python bench.py --steps 300 --warmup 5 2>&1 | tail -1 | python -c "
import json,sys; l=json.loads(sys.stdin.read()); print('$v', round(l['value'],1), round(l['ms_per_step'],4), round(l['e2e']['value'],1))"; done | tee gpurun_out/bench_ab3.log
MORL_THREE_STREAMS=1 timeout 900 python -m pytest tests/test_envelope_update_golden_gpu.py tests/test_envelope_update_gpu.py tests/test_qhead_envelope_gpu.py -q --maxfail 6 --timeout 600 2>&1 | tail -3 | tee gpurun_out/pytest_three.log
