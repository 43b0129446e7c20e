#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail 5 --timeout 600 2>&1 | tail -3 | tee gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke 2>&1 | grep -v Warning | tail -1 | tee gpurun_out/smoke.log
