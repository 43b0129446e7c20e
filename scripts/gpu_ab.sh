#!/bin/bash
# A/B of an environment toggle on the whole bench: usage  VAR=MORL_GEMM_L2HINT bash scripts/gpu_ab.sh
mkdir -p gpurun_out
V=${VAR:-MORL_GEMM_L2HINT}
for val in 0 1 0 1; do
  echo "== $V=$val"
  env $V=$val timeout 600 python bench.py --steps 150 --warmup 5 2>&1 | tail -1 | python -c "import sys,json; l=json.loads(sys.stdin.read()); print(l['value'], l['ms_per_step'], l['e2e']['value'], l['roofline']['us_per_launch'])"
done 2>&1 | tee gpurun_out/ab.log
