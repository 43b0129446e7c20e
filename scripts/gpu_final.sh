#!/bin/bash
# record run of the final build of round 2: GPU tests, smoke, the bench line, the in-situ kernel timeline
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q --maxfail 20 --timeout 600 2>&1 | tail -6 | tee gpurun_out/pytest_gpu.log
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | grep -v Warning | tail -2 | tee gpurun_out/smoke.log
echo "== bench"; timeout 900 python bench.py --steps 200 --warmup 5 2>&1 | tail -1 > gpurun_out/bench.log; python -c "
import json; l=json.loads(open('gpurun_out/bench.log').read()); print('bench', l['value'], l['ms_per_step'], l['e2e']['value'], l['roofline']['frac'], l['roofline']['us_per_launch'], l['roofline_envelope']['frac'], l['roofline_envelope']['us_per_launch'], l['roofline_envelope_operator']['frac'], l['gpu_launches'], l['cpu_baseline']['value'], l['roofline']['traffic'], l['roofline_envelope']['traffic'])"
echo "== kernel timeline"; timeout 300 python scripts/kernel_timeline.py 8 2>&1 | grep -v Warn > gpurun_out/kernel_timeline.log; head -12 gpurun_out/kernel_timeline.log
