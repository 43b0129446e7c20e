#!/bin/bash
mkdir -p gpurun_out
run() { echo "== $1"; env $1 timeout 600 python bench.py --steps 200 --warmup 5 2>&1 | tail -1 | python -c "import sys,json; l=json.loads(sys.stdin.read()); print(l['value'], l['ms_per_step'], l['e2e']['value'], l['clocks'])"; }
run "A=1"
run "MORL_TC_SNAKE=0"
run "MORL_TC_MULTI_SPLIT=0"
run "MORL_TC_SNAKE=0 MORL_TC_MULTI_SPLIT=0"
run "A=1"
