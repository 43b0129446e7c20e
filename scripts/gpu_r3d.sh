#!/bin/bash
# second optimisation batch: ReLU-bit stores batched, tree branch deferred, narrow output-layer kernel for the training pass, two-stream no-grad chains
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_qhead_envelope_gpu.py tests/test_gemm_gpu.py -q --maxfail 6 --timeout 200 2>&1 | tail -8 | tee gpurun_out/pytest_qhead.log
MORL_NARROW_HEAD=1 timeout 900 python -m pytest tests/test_envelope_update_golden_gpu.py tests/test_envelope_update_gpu.py tests/test_hv_parity_gpu.py -q --maxfail 6 --timeout 600 2>&1 | tail -5 | tee gpurun_out/pytest_narrow.log
for v in "MORL_DEFER_TREE=0" "MORL_DEFER_TREE=1" "MORL_NARROW_HEAD=1" "MORL_NARROW_HEAD=1 MORL_TWO_STREAMS=1" "MORL_TWO_STREAMS=1"; do env $v MORL_SKIP_CPU_BASELINE=1 timeout 600 python bench.py --steps 200 --warmup 5 2>&1 | tail -1 | python -c "
import json,sys; l=json.loads(sys.stdin.read()); print('$v', round(l['value'],1), round(l['ms_per_step'],4), round(l['e2e']['value'],1), round(l['roofline']['us_per_launch'],2), round(l['roofline_envelope_fused']['us_per_launch'],2))"; done | tee gpurun_out/bench_ab2.log
timeout 300 python scripts/gemm_time.py "" MORL_GEMM_SKIPB=1 MORL_GEMM_STAGES=2 "MORL_GEMM_SKIPB=1 MORL_GEMM_STAGES=2" 2>&1 | tee gpurun_out/gemm_time.log
MORL_NARROW_HEAD=1 timeout 300 python scripts/kernel_timeline.py 8 2>&1 | grep -v Warn > gpurun_out/kernel_timeline2.log; head -30 gpurun_out/kernel_timeline2.log
timeout 900 python -m pytest tests -m gpu -q --maxfail 10 --timeout 300 2>&1 | tail -6 | tee gpurun_out/pytest_gpu.log
