#!/bin/bash
# new split-operand GEMM: correctness, error statistics, role stats, whole-update goldens, bench
mkdir -p gpurun_out
echo "== gemm tests"; timeout 900 python -m pytest tests/test_gemm_gpu.py -m gpu -q -x --timeout 300 2>&1 | tail -25 | tee gpurun_out/pytest_gemm.log
echo "== error probe (split acc)"; timeout 300 python scripts/gemm_error_probe.py 2>&1 | grep -v Warn | tee gpurun_out/gemm_error_split.log
echo "== error probe (single acc)"; MORL_GEMM_SPLIT_ACC=0 timeout 300 python scripts/gemm_error_probe.py 2>&1 | grep -v Warn | tee gpurun_out/gemm_error_single.log
echo "== timing"; timeout 120 python scripts/gemm_probe.py --all 2>&1 | tail -6 | tee gpurun_out/gemm_probe_split.log
MORL_GEMM_SPLIT_ACC=0 timeout 120 python scripts/gemm_probe.py --all 2>&1 | tail -6 | tee gpurun_out/gemm_probe_single.log
echo "== role stats"; timeout 120 python scripts/gemm_stats.py f16x2 2>&1 | tail -5 | tee gpurun_out/gemm_stats_split.log
MORL_GEMM_SPLIT_ACC=0 timeout 120 python scripts/gemm_stats.py f16x2 2>&1 | tail -5 | tee gpurun_out/gemm_stats_single.log
echo "== golden diag f16x2"; timeout 600 python scripts/golden_diag.py north_star config2 2>&1 | grep -v Warning | grep "tc=True graph=True" -A 11 | tee gpurun_out/golden_diag_f16x2.log
echo "== golden diag f16x2 single acc"; MORL_GEMM_SPLIT_ACC=0 timeout 600 python scripts/golden_diag.py north_star 2>&1 | grep -v Warning | grep "tc=True graph=True" -A 11 | tee gpurun_out/golden_diag_f16x2_single.log
echo "== golden diag bf16x3"; MORL_TC_FMT=bf16x3 timeout 600 python scripts/golden_diag.py north_star 2>&1 | grep -v Warning | grep "tc=True graph=True" -A 11 | tee gpurun_out/golden_diag_bf16x3.log
echo "== update tests"; timeout 900 python -m pytest tests/test_envelope_update_gpu.py tests/test_kernels_gpu.py -m gpu -q --timeout 300 2>&1 | tail -15 | tee gpurun_out/pytest_update.log
echo "== bench" ; timeout 900 python bench.py --steps 100 --warmup 5 2>&1 | tail -1 | tee gpurun_out/bench.log
