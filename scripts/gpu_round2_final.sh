#!/bin/bash
# round-2 final artefacts (one gpurun call): GPU tests incl. the new ones, smoke, bench (both arms), A/B, MORL/D workload, ncu launch list +
# full captures of the fused head / envelope / GEMM kernels, compute-sanitizer over the new kernels
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
echo "== pytest -m gpu"; MORL_RUN_UNVALIDATED=1 timeout 1500 python -m pytest tests -m gpu -q --maxfail 20 --timeout 600 --durations=6 2>&1 | tail -30 | tee gpurun_out/pytest_gpu.log
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | grep -v Warning | tail -3 | tee gpurun_out/smoke.log
echo "== bench"; timeout 900 python bench.py --steps ${BENCH_STEPS:-200} --warmup 5 2>&1 | tail -1 > gpurun_out/bench.log; python -c "
import json; l=json.loads(open('gpurun_out/bench.log').read()); print('bench', l['value'], l['ms_per_step'], l['e2e']['value'], l['roofline']['frac'], l['roofline_envelope']['frac'], l['roofline_envelope_operator']['frac'], l['gpu_launches'], l['cpu_baseline']['value'])"
echo "== bench --impl reference"; MORL_CPU_BUDGET_S=${CPU_BUDGET:-80} timeout 900 python bench.py --impl reference --steps 20 --warmup 5 2>&1 | tail -1 | tee gpurun_out/bench_reference.log | cut -c1-300
echo "== bench A/B"; for v in "MORL_FUSED_HEAD=0 MORL_GEMM_CHAIN=0 MORL_GEMM_CHAIN_BWD=0 MORL_THREE_STREAMS=0 MORL_TWO_STREAMS=0 MORL_NARROW_HEAD=0 MORL_DEFER_TREE=0" "MORL_GEMM_CHAIN=0 MORL_GEMM_CHAIN_BWD=0" "MORL_GEMM_CHAIN_BWD=0" "MORL_THREE_STREAMS=0" "MORL_GEMM_PDL=0"; do env $v MORL_SKIP_CPU_BASELINE=1 timeout 600 python bench.py --steps 200 --warmup 5 2>&1 | tail -1 | python -c "
import json,sys; l=json.loads(sys.stdin.read()); print('$v', round(l['value'],1), round(l['ms_per_step'],4), round(l['e2e']['value'],1), round(l['roofline']['us_per_launch'],2))"; done | tee gpurun_out/bench_ab.log
echo "== qhead timing"; timeout 200 python scripts/qhead_time.py 2>&1 | grep -v Warn | tee gpurun_out/qhead_time.log
echo "== kernel timeline"; timeout 300 python scripts/kernel_timeline.py 8 2>&1 | grep -v Warn > gpurun_out/kernel_timeline.log; head -40 gpurun_out/kernel_timeline.log
echo "== morld workload"; timeout 600 python bench.py --workload morld --steps 20 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_morld.log | cut -c1-200
timeout 600 python scripts/golden_diag.py 2>&1 | grep -v Warning > gpurun_out/golden_diag.log; grep -c "tc=" gpurun_out/golden_diag.log
if [ "${DO_NCU:-1}" = "1" ]; then
echo "== ncu launch list (bench, short)"
MORL_SKIP_CPU_BASELINE=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 700 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 > gpurun_out/bench_under_ncu.log 2>&1
echo "== ncu full capture: fused head"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:qhead_envelope -s 2 -c 2 -o gpurun_out/prof_qhead -f python scripts/qhead_time.py > gpurun_out/ncu_qhead.log 2>&1
echo "== ncu full capture: chained hidden layers"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_chain -s 3 -c 2 -o gpurun_out/prof_chain -f python scripts/chain_probe.py > gpurun_out/ncu_chain.log 2>&1
echo "== ncu full capture: envelope operator"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:envelope_td -s 4 -c 2 -o gpurun_out/prof_envelope -f python scripts/profile_envelope.py 4 > gpurun_out/ncu_envelope.log 2>&1
echo "== ncu full capture: GEMM"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_planes -s 4 -c 3 -o gpurun_out/prof_gemm -f python scripts/gemm_probe.py > gpurun_out/ncu_gemm.log 2>&1
fi
if [ "${DO_SAN:-1}" = "1" ]; then
for tool in memcheck racecheck synccheck; do echo "== compute-sanitizer $tool (new kernels)"; CUDA_MODULE_LOADING=EAGER timeout 900 compute-sanitizer --tool $tool --print-limit 10 python scripts/sanitize_all.py qhead dyna chain 2>&1 | grep -v Warning | tail -12 | tee gpurun_out/sanitize_new_$tool.log; done
fi
ls gpurun_out | wc -l
