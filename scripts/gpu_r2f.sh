#!/bin/bash
mkdir -p gpurun_out
echo "== hv parity"; timeout 600 python -m pytest tests/test_hv_parity_gpu.py -m gpu -q -x -s --timeout 500 2>&1 | grep -v Warning | tail -25 | tee gpurun_out/pytest_hv.log
echo "== new tests"; timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_algos_update_gpu.py tests/test_envelope_update_golden_gpu.py -m gpu -q --timeout 500 -k "hypervolume or archive or population or adam or golden or unmodified or morld" 2>&1 | tail -15 | tee gpurun_out/pytest_new.log
echo "== morld bench (population graph)"; timeout 600 python bench.py --workload morld --steps 20 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_morld.log
echo "== morld bench (serial)"; MORL_POPULATION_GRAPH=0 timeout 600 python bench.py --workload morld --steps 20 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_morld_serial.log
echo "== ncu gemm (split + single)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_planes_kernel -s 6 -c 1 -o gpurun_out/prof_gemm -f python scripts/gemm_probe.py > gpurun_out/ncu_gemm.log 2>&1
MORL_GEMM_SPLIT_ACC=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_planes_kernel -s 6 -c 1 -o gpurun_out/prof_gemm_single -f python scripts/gemm_probe.py > gpurun_out/ncu_gemm_single.log 2>&1
ls -la gpurun_out/*.ncu-rep
