"""Small launches of every envelope-TD kernel family (v1, v3, wp, tc) for compute-sanitizer (memcheck / racecheck / synccheck):
    compute-sanitizer --tool racecheck python scripts/sanitize_envelope.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch as th

from morl_baselines_b200 import ops

dev = th.device("cuda:0")
g = th.Generator(device=dev).manual_seed(0)
for (B, W, A, D) in [(40, 64, 8, 3), (9, 48, 4, 2)]:
    q_on = th.randn(B, W, A, D, device=dev, generator=g)
    q_tg = th.randn(B, W, A, D, device=dev, generator=g)
    wset = th.rand(W, D, device=dev, generator=g)
    rew = th.randn(B, D, device=dev, generator=g)
    done = th.zeros(B, device=dev)
    ref = None
    for path in ("v1", "v3", "wp", "tc"):
        os.environ["MORL_ENVELOPE_PATH"] = path
        out = ops.envelope_td(q_on, q_tg, wset, rew, done, 0.99)
        th.cuda.synchronize()
        if ref is None:
            ref = out
        else:
            assert all(th.equal(a, b) for a, b in zip(ref, out)), path
    os.environ.pop("MORL_ENVELOPE_PATH", None)
pts = th.randn(700, 3, device=dev, generator=g, dtype=th.float64)
ops.pareto_mask(pts, True)
th.cuda.synchronize()
print("sanitize run ok")
