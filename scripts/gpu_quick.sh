#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemm_gpu.py tests/test_envelope_update_gpu.py -q -x --timeout 300 2>&1 | tail -6
timeout 120 python scripts/gemm_probe.py 2>&1 | tail -3
timeout 600 python bench.py --steps 100 --warmup 5 2>&1 | tail -1 | cut -c1-260 | tee gpurun_out/bench_quick.log
