"""Launch the fused envelope-TD kernel (and friends) a few times at the north-star shape -- the target of the ncu captures
(`ncu --set full -k regex:envelope_td ...`).  Not a benchmark: numbers printed under a profiler are never bench values."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch as th
from morl_baselines_b200 import ops

B, W, A, D = 1024, 64, 8, 3
dev = th.device("cuda:0")
g = th.Generator(device=dev).manual_seed(0)
q_on = th.randn(B, W, A, D, device=dev, generator=g)
q_tg = th.randn(B, W, A, D, device=dev, generator=g)
wset = th.rand(W, D, device=dev, generator=g); wset = wset / wset.sum(1, keepdim=True)
rew = th.randn(B, D, device=dev, generator=g)
done = th.zeros(B, device=dev)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 6):
    ops.envelope_td(q_on, q_tg, wset, rew, done, 0.99, ops.DOT_UNFUSED, ops.ROWS_BMAJOR, want_indices=False)
    ops.envelope_td(q_on, q_tg, wset, rew, done, 0.99, ops.DOT_UNFUSED, ops.ROWS_REFERENCE)
pts = th.randn(20000, 4, device=dev, generator=g)
ops.pareto_mask(pts, True)
th.cuda.synchronize()
print("done")
