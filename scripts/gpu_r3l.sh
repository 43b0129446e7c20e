#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"pair_layer1_grad|pairs_grad_reduce|sumtree_batch_set" -s 9 -c 3 -o gpurun_out/prof_tail -f python scripts/tail_probe.py > gpurun_out/ncu_tail.log 2>&1
tail -3 gpurun_out/ncu_tail.log
