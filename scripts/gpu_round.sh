#!/bin/bash
# One gpurun call: GPU parity tests, smoke, bench (both arms), ncu launch list + full captures of the envelope and GEMM kernels.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
echo "== pytest -m gpu" ; timeout 1500 python -m pytest tests -m gpu -q --maxfail 12 --timeout 900 --durations=8 2>&1 | tail -40 | tee gpurun_out/pytest_gpu.log
echo "== smoke" ; timeout 300 python __graft_entry__.py smoke 2>&1 | grep -v Warning | tail -5 | tee gpurun_out/smoke.log
echo "== bench" ; timeout 900 python bench.py --steps ${BENCH_STEPS:-200} --warmup 5 2>&1 | tail -1 | tee gpurun_out/bench.log
echo "== bench --impl reference" ; timeout 900 python bench.py --impl reference --steps 5 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_reference.log
[ -x morl_baselines_b200/csrc/_build/ffma2_rate ] && { echo "== ffma2 rate" ; timeout 120 morl_baselines_b200/csrc/_build/ffma2_rate 2>&1 | tee gpurun_out/ffma2_rate.log | tail -22 ; }
if [ "${DO_NCU:-1}" = "1" ]; then
echo "== ncu launch list (bench, short)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 700 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 > gpurun_out/bench_under_ncu.log 2>&1
echo "== ncu full capture of the envelope kernels"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:envelope_td -s 4 -c 2 -o gpurun_out/prof_envelope -f python scripts/profile_envelope.py 4 > gpurun_out/ncu_envelope.log 2>&1
echo "== ncu full capture of the GEMM kernels"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16x3 -s 2 -c 2 -o gpurun_out/prof_gemm -f python scripts/gemm_probe.py > gpurun_out/ncu_gemm.log 2>&1
fi
ls gpurun_out
