#!/bin/bash
mkdir -p gpurun_out
for m in f16x2 notc; do timeout 400 python scripts/hv_debug.py $m 6000 1000 2>&1 | grep "step" ; done | tee gpurun_out/hv_debug.log
timeout 400 python scripts/hv_debug.py f16x2 6000 6000 2>&1 | grep "step" | tee -a gpurun_out/hv_debug.log
