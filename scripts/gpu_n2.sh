#!/bin/bash
# 2-GPU validation: Envelope replicas + the one-collective front exchange over NCCL, MORL/D population sharding
mkdir -p gpurun_out
N=${NGPU:-2}
echo "== bench --gpus $N"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 50 --warmup 5 2>&1 | grep -v Warning | tail -3 | tee gpurun_out/bench_n$N.log | cut -c1-600
echo "== morld --gpus $N"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --workload morld --gpus $N --steps 20 --warmup 3 2>&1 | grep -v Warning | tail -3 | tee gpurun_out/bench_morld_n$N.log | cut -c1-600
echo "== reference arm under torchrun"; MORL_CPU_BUDGET_S=40 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus $N --steps 2 --warmup 1 2>&1 | grep -v Warning | tail -2 | cut -c1-300
