#!/bin/bash
# Round-2 artefact run (one gpurun call): GPU parity tests, smoke, bench (both arms + MORL/D workload), A/B of the GEMM switches, error
# probes, ncu launch list + full captures of the envelope and GEMM kernels, compute-sanitizer over every kernel.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
echo "== pytest -m gpu" ; timeout 2400 python -m pytest tests -m gpu -q --maxfail 20 --timeout 900 --durations=6 2>&1 | tail -30 | tee gpurun_out/pytest_gpu.log
echo "== smoke" ; timeout 300 python __graft_entry__.py smoke 2>&1 | grep -v Warning | tail -3 | tee gpurun_out/smoke.log
echo "== bench" ; timeout 900 python bench.py --steps ${BENCH_STEPS:-200} --warmup 5 2>&1 | tail -1 > gpurun_out/bench.log; python -c "
import json; l=json.loads(open('gpurun_out/bench.log').read()); print('bench', l['value'], l['ms_per_step'], l['e2e']['value'], l['roofline'], l['roofline_envelope']['frac'], l['gpu_launches'], l['config']['ms_eval_round_rank0'])"
echo "== bench --impl reference" ; MORL_CPU_BUDGET_S=${CPU_BUDGET:-100} timeout 900 python bench.py --impl reference --steps 20 --warmup 5 2>&1 | tail -1 | tee gpurun_out/bench_reference.log | cut -c1-400
echo "== bench A/B"; for v in "MORL_GEMM_PDL=0" "MORL_GEMM_SPLIT_ACC=1" "MORL_TC_FMT=bf16x3"; do env $v timeout 600 python bench.py --steps 100 --warmup 5 2>&1 | tail -1 | python -c "
import json,sys; l=json.loads(sys.stdin.read()); print('$v', round(l['value'],1), round(l['ms_per_step'],4), round(l['e2e']['value'],1), round(l['roofline']['us_per_launch'],2))"; done | tee gpurun_out/bench_ab.log
echo "== morld workload"; timeout 600 python bench.py --workload morld --steps 20 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_morld.log | cut -c1-300
MORL_POPULATION_GRAPH=0 timeout 600 python bench.py --workload morld --steps 20 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_morld_serial.log | cut -c1-200
echo "== error probes"; timeout 300 python scripts/gemm_error_probe.py 2>&1 | grep -v Warn | tee gpurun_out/gemm_error_single.log; MORL_GEMM_SPLIT_ACC=1 timeout 300 python scripts/gemm_error_probe.py 2>&1 | grep -v Warn | tee gpurun_out/gemm_error_split.log
timeout 600 python scripts/golden_diag.py 2>&1 | grep -v Warning > gpurun_out/golden_diag.log; grep -c "tc=" gpurun_out/golden_diag.log
timeout 120 python scripts/gemm_stats.py f16x2 2>&1 | tail -5 | tee gpurun_out/gemm_stats.log
if [ "${DO_NCU:-1}" = "1" ]; then
echo "== ncu launch list (bench, short)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 700 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 > gpurun_out/bench_under_ncu.log 2>&1
echo "== ncu full capture of the envelope kernels"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:envelope_td -s 4 -c 2 -o gpurun_out/prof_envelope -f python scripts/profile_envelope.py 4 > gpurun_out/ncu_envelope.log 2>&1
echo "== ncu full capture of the GEMM kernels"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_planes -s 4 -c 3 -o gpurun_out/prof_gemm -f python scripts/gemm_probe.py > gpurun_out/ncu_gemm.log 2>&1
fi
if [ "${DO_SAN:-1}" = "1" ]; then
for tool in memcheck racecheck synccheck initcheck; do echo "== compute-sanitizer $tool"; CUDA_MODULE_LOADING=EAGER timeout 1200 compute-sanitizer --tool $tool --print-limit 10 python scripts/sanitize_all.py 2>&1 | grep -v Warning | tail -14 | tee gpurun_out/sanitize_$tool.log; done
fi
ls gpurun_out | wc -l
