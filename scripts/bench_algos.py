"""Secondary measurements: the other hot-path rows of SURVEY.md section 8 at the shapes of BASELINE.md section 2 (reference-CPU
numbers quoted there were taken on an 8-vCPU Xeon during the survey).  Not the headline metric (bench.py); one JSON object on stdout.
Each `update()` is timed end to end through the public class API (host sampling, H2D, kernels, PER write-back), median of N calls
after warm-up, wall clock with a device synchronize on both sides."""
import json
import os
import random
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch as th

from morl_baselines_b200.testing import FakeEnv  # spaces-only stand-in for a mo-gymnasium env

dev = th.device("cuda:0")
out = {}


def timed(fn, n=30, warm=5):
    for _ in range(warm):
        fn()
    th.cuda.synchronize()
    ts = []
    for _ in range(n):
        t0 = time.perf_counter()
        fn()
        th.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    return float(np.median(ts))


def fill(rb, n, obs, act_dim, d, discrete, rng):
    rb.obs[:n] = rng.standard_normal((n, obs)).astype(np.float32)
    rb.next_obs[:n] = rng.standard_normal((n, obs)).astype(np.float32)
    rb.actions[:n] = rng.integers(0, act_dim, size=(n, 1)).astype(rb.actions.dtype) if discrete else rng.uniform(-1, 1, (n, act_dim)).astype(np.float32)
    rb.rewards[:n] = rng.standard_normal((n, d)).astype(np.float32)
    rb.dones[:n] = (rng.random((n, 1)) < 0.02).astype(np.float32)
    rb.size, rb.ptr = n, 0
    rb.mark_all_dirty()
    if hasattr(rb, "tree"):
        rb.tree.batch_set(np.arange(n), np.full(n, 0.1))


rng = np.random.default_rng(0)
random.seed(0)
np.random.seed(0)
th.manual_seed(0)
N = 16384

# ---- Envelope at the config-2 scale (|W| = 32, B = 256) and at the reference default (|W| = 4) ----
from morl_baselines_b200.multi_policy.envelope.envelope import Envelope

for W, ref in ((32, "1.13 s (0.89/s)"), (4, "15.4 ms (65/s)")):
    ag = Envelope(FakeEnv(obs_dim=32, n_actions=8, reward_dim=3), batch_size=256, num_sample_w=W, per=True, buffer_size=N, log=False, seed=0, device=dev)
    fill(ag.replay_buffer, N, 32, 8, 3, True, rng)
    ag.global_step = 1
    t = timed(ag.update)
    out[f"Envelope.update |W|={W} B=256"] = {"ms": t * 1e3, "updates_per_s": 1 / t, "reference_cpu_8vcpu": ref}

# ---- GPI-PD (discrete) ----
from morl_baselines_b200.multi_policy.gpi_pd.gpi_pd import GPIPD

for B, ref in ((128, "37.3 ms (26.8/s)"), (1024, "236.7 ms (4.2/s)")):
    ag = GPIPD(FakeEnv(obs_dim=32, n_actions=8, reward_dim=3), batch_size=B, gradient_updates=1, per=True, gpi_pd=True, dyna=False, buffer_size=N,
               log=False, seed=0, device=dev)
    fill(ag.replay_buffer, N, 32, 8, 3, True, rng)
    M = list(rng.dirichlet(np.ones(3), 64).astype(np.float32))
    ag.set_weight_support(M)
    ag.global_step = 1
    w = th.tensor(M[0]).to(dev)
    t = timed(lambda: ag.update(w))
    out[f"GPIPD.update B={B} |M|=64 gpi_pd"] = {"ms": t * 1e3, "updates_per_s": 1 / t, "reference_cpu_8vcpu": ref}
    obs_t = th.randn(B, 32, device=dev)
    wrow = w.reshape(1, 3).expand(B, 3).contiguous()
    t = timed(lambda: ag._envelope_target(obs_t, wrow, ag._support_matrix()))
    out[f"GPIPD._envelope_target B={B} |M|=64"] = {"ms": t * 1e3, "reference_cpu_8vcpu": "130 ms" if B == 128 else "1,148 ms"}
o1 = th.randn(32, device=dev)
t = timed(lambda: ag.gpi_action(o1, w), n=100)
out["GPIPD.gpi_action |M|=64"] = {"ms": t * 1e3, "reference_cpu_8vcpu": "0.9-1.0 ms"}

# ---- GPI-LS continuous (hopper dims) ----
from morl_baselines_b200.multi_policy.gpi_pd.gpi_pd_continuous_action import GPILSContinuousAction

ag = GPILSContinuousAction(FakeEnv(obs_dim=11, continuous_action_dim=3, reward_dim=3), batch_size=128, gradient_updates=1, per=True, buffer_size=N,
                           log=False, seed=0, device=dev)
fill(ag.replay_buffer, N, 11, 3, 3, False, rng)
M = list(rng.dirichlet(np.ones(3), 64).astype(np.float32))
ag.set_weight_support(M)
w = th.tensor(M[0]).to(dev)
t = timed(lambda: ag.update(w))
out["GPILSContinuousAction.update hopper B=128"] = {"ms": t * 1e3, "updates_per_s": 1 / t, "reference_cpu_8vcpu": "13.4 ms (75/s)"}
ag.use_gpi = True
ob = np.zeros(11, np.float32)
t = timed(lambda: ag.eval(ob, M[1]), n=100)
out["GPILSContinuousAction.eval GPI |M|=64"] = {"ms": t * 1e3, "reference_cpu_8vcpu": "13.2 ms"}

# ---- CAPQL (halfcheetah dims) ----
from morl_baselines_b200.multi_policy.capql.capql import CAPQL

ag = CAPQL(FakeEnv(obs_dim=17, continuous_action_dim=6, reward_dim=2), batch_size=128, log=False, seed=0, device=dev)
for _ in range(2048):
    ag.replay_buffer.push(rng.standard_normal(17), rng.uniform(-1, 1, 6), rng.dirichlet(np.ones(2)), rng.standard_normal(2), rng.standard_normal(17), 0.0)
t = timed(ag.update)
out["CAPQL.update halfcheetah B=128"] = {"ms": t * 1e3, "updates_per_s": 1 / t, "reference_cpu_8vcpu": "8.4 ms (120/s)"}

# ---- MOSAC (hopper dims; one MORL/D subproblem) ----
from morl_baselines_b200.single_policy.ser.mosac_continuous_action import MOSAC

ag = MOSAC(FakeEnv(obs_dim=11, continuous_action_dim=3, reward_dim=3), weights=np.array([0.3, 0.3, 0.4], np.float32), batch_size=128, log=False, seed=0,
           device=dev, buffer_size=N)
fill(ag.buffer, N, 11, 3, 3, False, rng)
ag.global_step = 0
t = timed(ag.update)
out["MOSAC.update hopper B=128"] = {"ms": t * 1e3, "updates_per_s": 1 / t, "reference_cpu_8vcpu": "13.2 ms (76/s)"}

# ---- Pareto prune ----
from morl_baselines_b200.common.pareto import filter_pareto_dominated

for n, d, ref in ((64, 3, "0.36 ms"), (600, 2, "20 ms"), (6000, 4, "2.03 s")):
    pts = rng.standard_normal((n, d))
    t = timed(lambda: filter_pareto_dominated(pts), n=20)
    out[f"filter_pareto_dominated N={n} d={d} (host array in, host array out)"] = {"ms": t * 1e3, "reference_cpu_8vcpu": ref}

print(json.dumps(out, indent=1))
