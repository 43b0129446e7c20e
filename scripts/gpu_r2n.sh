#!/bin/bash
mkdir -p gpurun_out
python scripts/gemm_time.py "" MORL_GEMM_SKIPB=1 MORL_GEMM_STAGES=2 "MORL_GEMM_SKIPB=1 MORL_GEMM_STAGES=2" MORL_GEMM_PDL=0 "MORL_GEMM_SKIPB=1 MORL_GEMM_PDL=0" 2>&1 | tee gpurun_out/gemm_time.log
