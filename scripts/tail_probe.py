"""ncu target for the kernels of the backward tail of the Envelope update (pair_layer1_grad, pairs_grad_reduce, sumtree_batch_set): a few
captured updates at the north-star shape."""
import os, sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch as th

import bench
from morl_baselines_b200.testing import synthetic_store

dev = th.device("cuda:0")
agent = bench._make_agent(dev, 0, True)
bench._fill_store(agent.replay_buffer, synthetic_store(bench.STORE, bench.OBS, bench.A, bench.D, seed=0))
agent.global_step = 1
for _ in range(6):
    agent.update()
th.cuda.synchronize()
print("done")
