#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/gemm_stats.py 2>&1 | tail -5 | tee gpurun_out/gemm_stats.log
