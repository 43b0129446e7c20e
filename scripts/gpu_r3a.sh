#!/bin/bash
# health check of HEAD after the container was re-created: GPU tests (no -x: list every failure), smoke, a bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail 10 --timeout 300 2>&1 | tail -25 | tee gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee gpurun_out/smoke.log
MORL_CPU_BASELINE_BUDGET_S=20 timeout 600 python bench.py --steps 200 --warmup 5 2>&1 | tail -1 > gpurun_out/bench.log
python - <<'PY'
import json
l = json.loads(open('gpurun_out/bench.log').read())
print('bench', l['value'], l['ms_per_step'], l['e2e'], l['clocks'], l['roofline']['bound'], l['roofline']['frac'], l['roofline']['us_per_launch'], l['roofline_envelope']['frac'], l['roofline_envelope']['us_per_launch'])
PY
timeout 200 python scripts/gemm_time.py "" 2>&1 | tail -12 | tee gpurun_out/gemm_time.log
timeout 300 python scripts/kernel_timeline.py 8 2>&1 | grep -v Warn | tee gpurun_out/kernel_timeline.log | head -70
