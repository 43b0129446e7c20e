"""Golden vectors for WHOLE updates of GPIPDContinuousAction / GPILSContinuousAction (reference
multi_policy/gpi_pd/gpi_pd_continuous_action.py:373-452) and its GPI evaluation (:454-485), produced by the unmodified reference on
CPU (run in the build container only):

    python tests/golden/make_golden_gpipd_continuous.py   ->  tests/golden/gpipd_continuous.npz

Dropout is disabled on the constructed networks (p = 0: CPU and CUDA generators differ) and the TD3 target-policy noise
(th.randn_like, :55) is drawn from a seeded numpy stream.  Stored: initial state_dicts, the seeded replay contents, and the
parameters / priorities / actions after the updates.
"""

from __future__ import annotations

import os
import random
import sys

import numpy as np
import torch as th

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_harness as rh  # noqa: E402

OBS, ACT, D, B, N = 11, 3, 3, 16, 128  # mo-hopper-v4 dimensions (BASELINE.json configs[2])


def sd_to_npz(out, prefix, sd):
    for k, v in sd.items():
        out[f"{prefix}/{k}"] = v.detach().cpu().numpy().copy()


class NoiseStream:
    def __init__(self, seed):
        self.rng = np.random.default_rng(seed)

    def __call__(self, shape):
        return th.from_numpy(self.rng.standard_normal(tuple(shape)).astype(np.float32))


def fill(rb, rng):
    rb.obs[:] = rng.standard_normal((N, OBS)).astype(np.float32)
    rb.next_obs[:] = rng.standard_normal((N, OBS)).astype(np.float32)
    rb.actions[:] = rng.uniform(-1, 1, (N, ACT)).astype(np.float32)
    rb.rewards[:] = rng.standard_normal((N, D)).astype(np.float32)
    rb.dones[:] = (rng.random((N, 1)) < 0.1).astype(np.float32)
    rb.size, rb.ptr = N, 0


def gen(out, tag, n_support):
    gm = rh.import_reference("morl_baselines.multi_policy.gpi_pd.gpi_pd_continuous_action")
    th.manual_seed(0)
    env = rh.FakeEnv(obs_dim=OBS, continuous_action_dim=ACT, reward_dim=D)
    agent = gm.GPILSContinuousAction(env, batch_size=B, net_arch=[32, 32], num_q_nets=2, gradient_updates=3, per=True, buffer_size=N,
                                     log=False, seed=3, device="cpu")
    for net in agent.q_nets + agent.target_q_nets:
        for m in net.modules():
            if isinstance(m, th.nn.Dropout):
                m.p = 0.0
    rng = np.random.default_rng(41)
    rb = agent.replay_buffer
    fill(rb, rng)
    rb.tree.batch_set(np.arange(N), rng.random(N) + 0.1)
    support = rng.dirichlet(np.ones(D), n_support).astype(np.float32)
    agent.set_weight_support(list(support))
    sd_to_npz(out, f"{tag}/init_policy", agent.policy.state_dict())
    for i, q in enumerate(agent.q_nets):
        sd_to_npz(out, f"{tag}/init_q{i}", q.state_dict())
    for k in ("obs", "next_obs", "actions", "rewards", "dones"):
        out[f"{tag}/rb_{k}"] = getattr(rb, k).copy()
    out[f"{tag}/tree_leaves0"] = rb.tree.nodes[-1].copy()
    out[f"{tag}/support"] = support
    w = th.tensor(support[min(2, n_support - 1)])
    stream = NoiseStream(99)
    orig = th.randn_like
    th.randn_like = lambda t, **kw: stream(t.shape)
    try:
        random.seed(15)
        np.random.seed(16)
        agent.global_step = 5
        agent.update(w)
    finally:
        th.randn_like = orig
    sd_to_npz(out, f"{tag}/final_policy", agent.policy.state_dict())
    sd_to_npz(out, f"{tag}/final_target_policy", agent.target_policy.state_dict())
    for i, (q, tq) in enumerate(zip(agent.q_nets, agent.target_q_nets)):
        sd_to_npz(out, f"{tag}/final_q{i}", q.state_dict())
        sd_to_npz(out, f"{tag}/final_tq{i}", tq.state_dict())
    out[f"{tag}/tree_leaves1"] = rb.tree.nodes[-1].copy()
    out[f"{tag}/min_priority1"] = np.float64(rb.min_priority)
    # eval: plain policy and GPI over the support (:454-485)
    obs_eval = rng.standard_normal((6, OBS)).astype(np.float32)
    out[f"{tag}/eval_obs"] = obs_eval
    wq = support[0]
    agent.use_gpi = False
    out[f"{tag}/eval_plain"] = np.stack([agent.eval(o, wq) for o in obs_eval])
    agent.use_gpi = True
    out[f"{tag}/eval_gpi"] = np.stack([agent.eval(o, wq) for o in obs_eval])
    print(tag, "done")


def main():
    assert rh.reference_available()
    th.set_num_threads(os.cpu_count() or 1)
    out = {}
    gen(out, "m5", 5)   # |M| > 1: the doubled batch with random.choices over the support (:381-391)
    gen(out, "m1", 1)   # |M| = 1: plain batch
    path = os.path.join(HERE, "gpipd_continuous.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path))


if __name__ == "__main__":
    main()
