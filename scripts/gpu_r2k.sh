#!/bin/bash
mkdir -p gpurun_out
echo "== failing test + new ones"; timeout 900 python -m pytest tests/test_hv_parity_gpu.py tests/test_gemm_gpu.py -m gpu -q --timeout 600 2>&1 | tail -6 | tee gpurun_out/pytest_part.log
echo "== bench"; timeout 900 python bench.py --steps 100 --warmup 5 2>&1 | tail -1 > gpurun_out/bench.log; python -c "
import json; l=json.loads(open('gpurun_out/bench.log').read()); print('bench', l['value'], l['ms_per_step'], l['e2e']['value'], l['roofline']['frac'], l['roofline']['us_per_launch'], l['gpu_launches'])"
echo "== launch list"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 700 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 > gpurun_out/bench_under_ncu.log 2>&1
for z in 0 1; do echo "== initcheck SAN_ZERO_PLANES=$z"; SAN_ZERO_PLANES=$z CUDA_MODULE_LOADING=EAGER timeout 900 compute-sanitizer --tool initcheck --print-limit 3 python scripts/sanitize_all.py gemm layer1 2>&1 | grep -v Warning | (head -30; echo ...; tail -4) | tee gpurun_out/sanitize_initcheck_zero$z.log; done
echo "== racecheck (head)"; CUDA_MODULE_LOADING=EAGER timeout 900 compute-sanitizer --tool racecheck --print-limit 4 python scripts/sanitize_all.py gemm 2>&1 | grep -v Warning | (head -40; echo ...; tail -3) | tee gpurun_out/sanitize_racecheck.log
echo "== memcheck (head)"; CUDA_MODULE_LOADING=EAGER timeout 900 compute-sanitizer --tool memcheck --print-limit 4 python scripts/sanitize_all.py 2>&1 | grep -v Warning | (head -30; echo ...; tail -3) | tee gpurun_out/sanitize_memcheck.log
