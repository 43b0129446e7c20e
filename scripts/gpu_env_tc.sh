#!/bin/bash
# tensor-core envelope kernel: agreement sweep + timing, per-phase cycle accounting, one ncu full capture
mkdir -p gpurun_out
timeout 300 python scripts/envelope_ab.py 2>&1 | grep -v Warning | tail -12 | tee gpurun_out/envelope_ab.log
timeout 120 python scripts/envelope_stats.py 2>&1 | grep -v Warning | tail -12 | tee gpurun_out/envelope_stats.log
if [ "${DO_NCU:-1}" = "1" ]; then
timeout 300 ncu --set full --clock-control none --import-source on -k regex:envelope_td_tc -s 4 -c 2 -o gpurun_out/prof_envelope_tc -f python scripts/profile_envelope.py 4 > gpurun_out/ncu_envelope_tc.log 2>&1
tail -3 gpurun_out/ncu_envelope_tc.log
fi
