#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py --steps 200 --warmup 5 2>&1 | tail -1 | tee gpurun_out/bench.log | python -c "import sys,json; l=json.loads(sys.stdin.read()); print(l['value'], l['ms_per_step'], l['e2e'], l['clocks'], l['roofline']['frac'], l['roofline_gemm']['us_per_launch'])"
