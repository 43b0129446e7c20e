#!/bin/bash
# usage: gpurun --gpus N -- bash scripts/gpu_multi.sh N
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv | tee gpurun_out/gpus_$N.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 100 --warmup 5 2>&1 | tail -3 | tee gpurun_out/bench_n$N.log | cut -c1-600
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 2 --warmup 1 2>&1 | tail -2 | tee gpurun_out/bench_ref_n$N.log | cut -c1-400
