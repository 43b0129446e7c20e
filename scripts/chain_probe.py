"""ncu target of the chained hidden-layer launch (`-k regex:gemm_chain`): the 2-chain x 3-layer launch of the no-grad passes at the
north-star shape on rotating buffers, + its in-graph timing (bench.time_chain_kernel)."""
import os, sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch as th

import bench

dev = th.device("cuda:0")
t, n_prod, flops, nbytes = bench.time_chain_kernel(dev, replays=3)
print(f"gemm_chain_kernel 2 x 3 layers: {t * 1e6:.1f} us per launch = {t * 1e6 / n_prod:.1f} us per layer product; MMA issued {flops / t / 1e12:.0f} TFLOP/s, "
      f"algorithmic HBM bytes {nbytes / 1e6:.0f} MB -> {nbytes / t / 1e9:.0f} GB/s")
