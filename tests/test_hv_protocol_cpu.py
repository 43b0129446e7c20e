"""CPU checks of the hypervolume-parity protocol's fixtures: the stand-in MOMDP, its evaluation-weight list and
the frozen reference run (tests/golden/hv_parity.json, produced by the unmodified reference on CPU) are self-consistent, so the GPU test
(tests/test_hv_parity_gpu.py) compares against a meaningful target."""

import json
import os

import numpy as np

from tests.golden.standin_env import HV_REF_POINT, N_POS, TreasureChain, robust_eval_weights

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _rollout(policy, gamma):
    env = TreasureChain()
    obs, _ = env.reset()
    ret, g, done, steps = np.zeros(3), 1.0, False, 0
    while not done:
        obs, r, term, trunc, _ = env.step(policy(obs))
        ret += g * r
        g *= gamma
        done = term or trunc
        steps += 1
    return ret, steps


def test_standin_env_true_front_and_margins():
    from morl_baselines_b200.common.performance_indicators import hypervolume
    from oracle import oracle as orc  # CPU restatement of the reference's non-dominated filter (the product's prune is CUDA-only)

    gamma = 0.98
    env = TreasureChain()
    front = env.pareto_front(gamma)
    assert len(front) == 2 * N_POS
    # every "walk x steps, then collect A|B" policy reproduces its analytic discounted return, and episodes always terminate
    k = 0
    for x in range(N_POS):
        for collect in (1, 2):
            ret, steps = _rollout(lambda o, x=x, collect=collect: 0 if int(np.argmax(o[:N_POS])) < x else collect, gamma)
            np.testing.assert_allclose(ret, front[k], rtol=1e-6)
            assert steps == x + 1
            k += 1
    ret, steps = _rollout(lambda o: 0, gamma)  # walking off the end terminates empty-handed
    assert steps == N_POS and ret[0] == 0 and ret[1] == 0
    # the six policies are mutually non-dominated and each wins two evaluation weights with a comfortable margin
    assert int(orc.pareto_mask(np.array(front), True).sum()) == 2 * N_POS
    ew = robust_eval_weights(gamma)
    assert len(ew) == 4 * N_POS and sorted({p for _, _, p in ew}) == list(range(2 * N_POS))
    assert min(g for _, g, _ in ew) > 0.85
    assert all(abs(float(w.sum()) - 1.0) < 1e-6 and (w > 0).all() for w, _, _ in ew)
    assert hypervolume(HV_REF_POINT, front) > 0


def test_frozen_reference_run_is_consistent():
    from morl_baselines_b200.common.performance_indicators import hypervolume

    g = json.load(open(os.path.join(ROOT, "tests", "golden", "hv_parity.json")))
    gamma = g["hyper_parameters"]["gamma"]
    np.testing.assert_allclose(g["true_front_hv"], hypervolume(HV_REF_POINT, TreasureChain().pareto_front(gamma)), rtol=1e-12)
    np.testing.assert_allclose(np.array(g["eval_weights"]), np.array([w for w, _, _ in robust_eval_weights(gamma)]), rtol=1e-6)
    assert len(g["seeds"]) >= 3
    for rec in g["seeds"].values():
        np.testing.assert_allclose(rec["hv"], hypervolume(HV_REF_POINT, rec["front"]), rtol=1e-9)
        assert rec["hv"] <= g["true_front_hv"] * (1 + 1e-6)
    np.testing.assert_allclose(g["hv_mean"], np.mean([r["hv"] for r in g["seeds"].values()]), rtol=1e-12)
