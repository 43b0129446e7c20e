#!/bin/bash
# bring-up of the fused output-layer + envelope kernel, then the health check of the whole tree
mkdir -p gpurun_out
MORL_RUN_UNVALIDATED=1 timeout 600 python -m pytest tests/test_qhead_envelope_gpu.py tests/test_dyna_gpu.py -q --maxfail 6 --timeout 200 2>&1 | tail -30 | tee gpurun_out/pytest_qhead.log
timeout 200 python scripts/qhead_time.py 2>&1 | grep -v Warn | tee gpurun_out/qhead_time.log
timeout 900 python -m pytest tests -m gpu -q --maxfail 10 --timeout 300 --deselect tests/test_qhead_envelope_gpu.py 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee gpurun_out/smoke.log
MORL_FUSED_HEAD=1 MORL_SKIP_CPU_BASELINE=1 timeout 600 python bench.py --steps 200 --warmup 5 2>&1 | tail -1 > gpurun_out/bench.log
python - <<'PY'
import json
l = json.loads(open('gpurun_out/bench.log').read())
print('bench', l['value'], l['ms_per_step'], l['e2e'], l['clocks'], l['roofline']['bound'], l['roofline']['frac'], l['roofline']['us_per_launch'], l['roofline_envelope']['frac'], l['roofline_envelope']['us_per_launch'])
PY
MORL_FUSED_HEAD=0 MORL_SKIP_CPU_BASELINE=1 timeout 600 python bench.py --steps 200 --warmup 5 2>&1 | tail -1 | python -c "
import json,sys; l=json.loads(sys.stdin.read()); print('MORL_FUSED_HEAD=0', l['value'], l['ms_per_step'], l['e2e']['value'])" | tee gpurun_out/bench_unfused.log
MORL_FUSED_HEAD=1 timeout 300 python scripts/kernel_timeline.py 8 2>&1 | grep -v Warn | tee gpurun_out/kernel_timeline.log | head -70
MORL_FUSED_HEAD=1 MORL_TWO_STREAMS=1 MORL_SKIP_CPU_BASELINE=1 timeout 600 python bench.py --steps 200 --warmup 5 2>&1 | tail -1 | python -c "
import json,sys; l=json.loads(sys.stdin.read()); print('MORL_TWO_STREAMS=1', l['value'], l['ms_per_step'], l['e2e']['value'])" | tee gpurun_out/bench_two_streams.log
timeout 300 python scripts/gemm_time.py "" MORL_GEMM_SKIPB=1 MORL_GEMM_STAGES=2 "MORL_GEMM_SKIPB=1 MORL_GEMM_STAGES=2" 2>&1 | tee gpurun_out/gemm_time.log
