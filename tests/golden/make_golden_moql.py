"""Golden Q-tables of the reference's tabular MOQLearning (BASELINE.json configs[0], the CPU-runnable case) on the stand-in MOMDP:
    python tests/golden/make_golden_moql.py   ->  tests/golden/moql.npz     (build container only: needs /root/reference)"""

import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_harness as rh  # noqa: E402
from tests.golden.standin_env import TreasureChain  # noqa: E402

CASES = {"w_time": (np.array([0.2, 0.1, 0.7]), 0), "w_a": (np.array([0.8, 0.1, 0.1]), 1), "w_b": (np.array([0.1, 0.8, 0.1]), 2)}
STEPS = 3000


def main():
    assert rh.reference_available()
    mq = rh.import_reference("morl_baselines.single_policy.ser.mo_q_learning")
    out = {}
    for tag, (w, seed) in CASES.items():
        env = TreasureChain(seed=seed)
        agent = mq.MOQLearning(env, weights=w, learning_rate=0.1, gamma=0.98, initial_epsilon=1.0, final_epsilon=0.05, epsilon_decay_steps=2000,
                               log=False, seed=seed)
        agent.train(time.time(), total_timesteps=STEPS)
        keys = sorted(agent.q_table)
        out[f"{tag}/keys"] = np.array(keys, dtype=np.float64)
        out[f"{tag}/values"] = np.stack([agent.q_table[k] for k in keys])
        out[f"{tag}/epsilon"] = np.float64(agent.epsilon)
        out[f"{tag}/num_episodes"] = np.int64(agent.num_episodes)
        obs, _ = env.reset()
        acts = []
        done = False
        while not done:
            a = agent.eval(obs, w)
            acts.append(a)
            obs, _, term, trunc, _ = env.step(a)
            done = term or trunc
        out[f"{tag}/greedy_actions"] = np.array(acts, np.int64)
        print(tag, len(keys), "states, greedy", acts)
    np.savez_compressed(os.path.join(HERE, "moql.npz"), **out)


if __name__ == "__main__":
    main()
