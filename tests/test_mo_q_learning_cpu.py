"""BASELINE.json configs[0] -- tabular MOQLearning, the reference's CPU-runnable configuration ("plumbing, runs without a GPU").
The mirror class needs no CUDA device and must reproduce the reference's Q-tables BIT FOR BIT on the same environment, seeds and
hyper-parameters (float64 numpy arithmetic in the same order; tests/golden/make_golden_moql.py froze the unmodified reference)."""

import os
import time

import numpy as np
import pytest

from tests.golden.standin_env import TreasureChain

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = {"w_time": (np.array([0.2, 0.1, 0.7]), 0), "w_a": (np.array([0.8, 0.1, 0.1]), 1), "w_b": (np.array([0.1, 0.8, 0.1]), 2)}


@pytest.mark.parametrize("tag", sorted(CASES))
def test_moqlearning_matches_reference_tables(tag):
    from morl_baselines_b200.single_policy.ser.mo_q_learning import MOQLearning

    g = np.load(os.path.join(ROOT, "tests", "golden", "moql.npz"))
    w, seed = CASES[tag]
    env = TreasureChain(seed=seed)
    agent = MOQLearning(env, weights=w, learning_rate=0.1, gamma=0.98, initial_epsilon=1.0, final_epsilon=0.05, epsilon_decay_steps=2000, log=False,
                        seed=seed)
    agent.train(time.time(), total_timesteps=3000)
    keys = sorted(agent.q_table)
    assert np.array_equal(np.array(keys, dtype=np.float64), g[f"{tag}/keys"])
    assert np.array_equal(np.stack([agent.q_table[k] for k in keys]), g[f"{tag}/values"])  # bit-exact
    assert float(agent.epsilon) == float(g[f"{tag}/epsilon"]) and agent.num_episodes == int(g[f"{tag}/num_episodes"])
    obs, _ = env.reset()
    acts, done = [], False
    while not done:
        a = agent.eval(obs, w)
        acts.append(a)
        obs, _, term, trunc, _ = env.step(a)
        done = term or trunc
    assert acts == list(g[f"{tag}/greedy_actions"])
    # the learnt greedy policy is the optimal one of the deterministic chain for these weights
    front = np.array(env.pareto_front(0.98))
    best = int(np.argmax(front @ w))
    assert (len(acts) - 1, acts[-1]) == (best // 2, 1 + best % 2)


def test_moqlearning_api_surface():
    from morl_baselines_b200.single_policy.ser.mo_q_learning import MOQLearning

    env = TreasureChain()
    agent = MOQLearning(env, weights=np.array([0.3, 0.3, 0.4]), log=False, seed=0)
    assert agent.device.type == "cpu" and agent.action_dim == 3 and agent.reward_dim == 3
    obs, _ = env.reset()
    assert agent.scalarized_q_values(obs, agent.weights).shape == (3,)
    assert 0 <= agent.eval(obs, agent.weights) < 3
    cfg = agent.get_config()
    assert cfg["scalarization"] == "weighted_sum" and cfg["dyna"] is False
    with pytest.raises(NotImplementedError):
        MOQLearning(env, dyna=True, log=False)
