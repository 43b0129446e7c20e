#!/bin/bash
# One gpurun call: GPU parity tests, smoke, bench, ncu launch list + full capture of the envelope kernel.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
echo "== pytest -m gpu" ; timeout 900 python -m pytest tests -m gpu -q --maxfail 12 --timeout 600 --ignore tests/test_gemm_gpu.py 2>&1 | tail -40 | tee gpurun_out/pytest_gpu.log
echo "== gemm"; timeout 300 python -m pytest tests/test_gemm_gpu.py -q -x --timeout 200 2>&1 | tail -15 | tee gpurun_out/gemm_test.log; timeout 120 python scripts/gemm_probe.py 2>&1 | tail -5 | tee gpurun_out/gemm_probe.log
echo "== smoke" ; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -15 | tee gpurun_out/smoke.log
echo "== bench" ; timeout 900 python bench.py --steps ${BENCH_STEPS:-100} --warmup 5 2>&1 | tail -15 | tee gpurun_out/bench.log
if [ "${DO_NCU:-1}" = "1" ]; then
echo "== ncu launch list (bench, short)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 > gpurun_out/bench_under_ncu.log 2>&1
tail -3 gpurun_out/bench_under_ncu.log
echo "== ncu full capture of the envelope kernel"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:envelope_td -s 4 -c 2 -o gpurun_out/prof_envelope -f python scripts/profile_envelope.py 4 > gpurun_out/ncu_envelope.log 2>&1
tail -3 gpurun_out/ncu_envelope.log
fi
ls -la gpurun_out
