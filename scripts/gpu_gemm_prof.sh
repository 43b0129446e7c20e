#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16x3 -s 2 -c 2 -o gpurun_out/prof_gemm -f python scripts/gemm_probe.py > gpurun_out/ncu_gemm.log 2>&1
tail -3 gpurun_out/ncu_gemm.log
