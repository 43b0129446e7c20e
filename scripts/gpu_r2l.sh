#!/bin/bash
# round 2, session 3: ReLU bit masks + runtime TMA ring depth -- correctness, bench A/B, launch list
mkdir -p gpurun_out
echo "== gemm tests"; timeout 1200 python -m pytest tests/test_gemm_gpu.py -m gpu -q -x --timeout 600 2>&1 | tail -8 | tee gpurun_out/pytest_gemm.log
echo "== update / golden / checkpoint tests"; timeout 1500 python -m pytest tests/test_envelope_update_golden_gpu.py tests/test_envelope_update_gpu.py -m gpu -q --timeout 900 2>&1 | tail -8 | tee gpurun_out/pytest_update.log
echo "== bench"; timeout 900 python bench.py --steps 100 --warmup 5 2>&1 | tail -1 > gpurun_out/bench.log; python -c "
import json; l=json.loads(open('gpurun_out/bench.log').read()); print('bench', l['value'], l['ms_per_step'], l['e2e']['value'], l['roofline']['frac'], l['roofline']['us_per_launch'], l['gpu_launches'])"
echo "== bench MORL_GEMM_STAGES=3"; MORL_GEMM_STAGES=3 timeout 900 python bench.py --steps 100 --warmup 5 2>&1 | tail -1 > gpurun_out/bench_stages3.log; python -c "
import json; l=json.loads(open('gpurun_out/bench_stages3.log').read()); print('bench', l['value'], l['ms_per_step'], l['e2e']['value'])"
echo "== launch list"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 700 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 > gpurun_out/bench_under_ncu.log 2>&1
python scripts/summarize_profiles.py r02l > /dev/null; sed -n '/one gradient update/,$p' profiles/r02l_launches.txt | cut -c1-150; cp profiles/r02l_launches.txt gpurun_out/
