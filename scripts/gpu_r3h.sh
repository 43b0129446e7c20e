#!/bin/bash
mkdir -p gpurun_out
for v in "MORL_HEAD_REVERSE=0" "MORL_HEAD_REVERSE=1" "MORL_PRE_REFRESH=1" "MORL_HEAD_REVERSE=0" "MORL_HEAD_REVERSE=1" "MORL_PRE_REFRESH=1" "MORL_THREE_STREAMS=0"; do env $v MORL_SKIP_CPU_BASELINE=1 timeout 600 python bench.py --steps 300 --warmup 5 2>&1 | tail -1 | python -c "
import json,sys; l=json.loads(sys.stdin.read()); print('$v', round(l['value'],1), round(l['ms_per_step'],4), round(l['e2e']['value'],1))"; done | tee gpurun_out/bench_ab6.log
MORL_THREE_STREAMS=0 timeout 300 python scripts/kernel_timeline.py 8 2>&1 | grep -v Warn > gpurun_out/kernel_timeline5.log; head -34 gpurun_out/kernel_timeline5.log
