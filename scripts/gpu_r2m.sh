#!/bin/bash
# round 2, session 3: PDL on every kernel of the update, batched loads in the layer-1 backward reductions, H=256 first-layer kernel
mkdir -p gpurun_out
echo "== gemm + kernel tests"; timeout 1500 python -m pytest tests/test_gemm_gpu.py tests/test_kernels_gpu.py -m gpu -q -x --timeout 600 2>&1 | tail -8 | tee gpurun_out/pytest_gemm.log
echo "== update / golden tests"; timeout 1500 python -m pytest tests/test_envelope_update_golden_gpu.py tests/test_envelope_update_gpu.py -m gpu -q --timeout 900 2>&1 | tail -8 | tee gpurun_out/pytest_update.log
for pdl in 1 0; do
echo "== bench MORL_PDL=$pdl"; MORL_PDL=$pdl timeout 900 python bench.py --steps 100 --warmup 5 2>&1 | tail -1 > gpurun_out/bench_pdl$pdl.log; python -c "
import json; l=json.loads(open('gpurun_out/bench_pdl$pdl.log').read()); print('bench', l['value'], l['ms_per_step'], l['e2e']['value'], l['roofline']['frac'], l['roofline']['us_per_launch'], l['gpu_launches'])"
done
cp gpurun_out/bench_pdl1.log gpurun_out/bench.log
echo "== launch list (default cache control)"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 700 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 > gpurun_out/bench_under_ncu.log 2>&1
python scripts/summarize_profiles.py r02m > /dev/null; sed -n '/one gradient update/,$p' profiles/r02m_launches.txt | cut -c1-150; cp profiles/r02m_launches.txt gpurun_out/
echo "== launch list (warm caches: --cache-control none)"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s 600 -c 700 --csv --log-file gpurun_out/launches_warm.csv python bench.py --steps 3 --warmup 3 > gpurun_out/bench_under_ncu2.log 2>&1
