#!/bin/bash
mkdir -p gpurun_out
CUDA_MODULE_LOADING=EAGER timeout 140 compute-sanitizer --tool memcheck --print-limit 10 python scripts/sanitize_all.py chain qhead 2>&1 | grep -v Warning | tail -14 | tee gpurun_out/sanitize_new_memcheck.log
