"""Hypervolume-parity reference run at BASELINE configs[1] hyper-parameters (|W| = 32, batch 256, net 4 x 256, per=True) on a budget
SHORT enough that the reference does NOT reach the true Pareto front (the round-1 fixture, hv_parity.json, saturates at the true front
for every seed and therefore cannot detect a moderate regression).

To keep the comparison a test of the UPDATE PATH rather than of chaotic exploration, the behaviour policy is uniformly random throughout
(initial_epsilon = final_epsilon = 1): the replay contents are then a function of the environment / action-sampling seeds only and are
identical for both engines; what differs is what each engine LEARNS from them.  The discounted returns of the greedy policy for the fixed
evaluation-weight list, the non-dominated front and its hypervolume after `TOTAL_STEPS` environment steps are frozen per seed, together
with intermediate checkpoints (informational).

    python tests/golden/make_golden_hv_config2.py         (build container only: needs /root/reference; ~1.1 s per update on 8 cores)
"""

from __future__ import annotations

import json
import os
import sys
import time

import numpy as np
import torch as th

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_harness as rh  # noqa: E402
from tests.golden.make_golden_hv import evaluate  # noqa: E402
from tests.golden.standin_env import HV_REF_POINT, TreasureChain, robust_eval_weights  # noqa: E402
from morl_baselines_b200.common.performance_indicators import hypervolume as hypervolume_exact  # noqa: E402

HP = dict(learning_rate=3e-4, initial_epsilon=1.0, final_epsilon=1.0, epsilon_decay_steps=None, tau=1.0, target_net_update_freq=200,
          buffer_size=4096, net_arch=[256, 256, 256, 256], batch_size=256, learning_starts=100, gradient_updates=1, gamma=0.98,
          max_grad_norm=1.0, envelope=True, num_sample_w=32, per=True, per_alpha=0.6, initial_homotopy_lambda=0.0,
          final_homotopy_lambda=1.0, homotopy_decay_steps=None)
CHECKPOINTS = [int(x) for x in os.environ.get("HV2_CHECKPOINTS", "200,300,400").split(",")]
SEEDS = [int(x) for x in os.environ.get("HV2_SEEDS", "0,1,2").split(",")]


def main():
    assert rh.reference_available()
    em = rh.import_reference("morl_baselines.multi_policy.envelope.envelope")
    em.equally_spaced_weights = lambda dim, n, seed=42: [w for w, _, _ in robust_eval_weights(HP["gamma"])]
    th.set_num_threads(len(os.sched_getaffinity(0)))
    ew = robust_eval_weights(HP["gamma"])
    weights = [w for w, _, _ in ew]
    path = os.path.join(HERE, "hv_parity_config2.json")
    out = {"hyper_parameters": HP, "checkpoints": CHECKPOINTS, "total_timesteps": CHECKPOINTS[-1], "ref_point": HV_REF_POINT.tolist(),
           "eval_weights": [list(map(float, w)) for w in weights], "env": "TreasureChain (tests/golden/standin_env.py)", "seeds": {}}
    out["true_front_hv"] = hypervolume_exact(HV_REF_POINT, TreasureChain().pareto_front(HP["gamma"]))
    for seed in SEEDS:
        th.manual_seed(seed)
        np.random.seed(seed)
        env = TreasureChain(seed=seed)
        agent = em.Envelope(env, log=False, seed=seed, device="cpu", **HP)
        rec, done_steps = {}, 0
        for cp in CHECKPOINTS:
            t0 = time.time()
            agent.train(total_timesteps=cp - done_steps, reset_num_timesteps=False)
            done_steps = cp
            returns, front, hv = evaluate(agent, HP["gamma"], weights)
            rec[str(cp)] = {"hv": hv, "n_front": len(front), "returns": [list(map(float, p)) for p in returns]}
            print(f"seed {seed} @ {cp} steps: hv {hv:.4f} / true {out['true_front_hv']:.4f} ({100 * hv / out['true_front_hv']:.1f} %), |front| {len(front)}, "
                  f"{time.time() - t0:.0f} s", flush=True)
        out["seeds"][str(seed)] = rec
        json.dump(out, open(path, "w"), indent=1)
    last = str(CHECKPOINTS[-1])
    out["hv_mean"] = float(np.mean([v[last]["hv"] for v in out["seeds"].values()]))
    json.dump(out, open(path, "w"), indent=1)
    print("hv mean", out["hv_mean"], "true", out["true_front_hv"])


if __name__ == "__main__":
    main()
