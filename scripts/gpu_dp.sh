#!/bin/bash
# DP-Envelope on 2 GPUs: parity against the single-GPU update, then the strong-scaling bench line (1 GPU and 2 GPUs, same code path)
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tests/dp_envelope_check.py 2>&1 | grep -v Warn | tail -8 | tee gpurun_out/dp_check.log
timeout 300 python bench.py --workload envelope_dp --steps 200 --warmup 5 2>&1 | tail -1 | tee gpurun_out/bench_dp_n1.log | cut -c1-400
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --workload envelope_dp --gpus 2 --steps 200 --warmup 5 2>&1 | tail -1 | tee gpurun_out/bench_dp_n2.log | cut -c1-400
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 --steps 200 --warmup 5 2>&1 | tail -1 | tee gpurun_out/bench_n2.log | cut -c1-300
