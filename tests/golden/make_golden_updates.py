"""Golden vectors for WHOLE updates of the other in-scope algorithms, produced by the unmodified reference on CPU
(run in the build container only):   python tests/golden/make_golden_updates.py   ->  tests/golden/updates.npz

  * GPIPD.update (multi_policy/gpi_pd/gpi_pd.py:416-562), gpi_pd = True and False (GPI-LS), dropout disabled
  * CAPQL.update (multi_policy/capql/capql.py:321-362) with the reparameterisation noise injected from a seeded numpy stream
  * MOSAC.update (single_policy/ser/mosac_continuous_action.py:429-507), same noise injection
Stored: initial state_dicts, the seeded inputs, and the parameters / priorities after the updates.
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


def sd_to_npz(out, prefix, sd):
    for k, v in sd.items():
        out[f"{prefix}/{k}"] = v.detach().cpu().numpy().copy()


def gen_gpipd(out):
    gm = rh.import_reference("morl_baselines.multi_policy.gpi_pd.gpi_pd")
    OBS, A, D, B, N = 10, 4, 3, 16, 256
    for gpi_pd in (True, False):
        tag = f"gpipd{int(gpi_pd)}"
        th.manual_seed(0)
        agent = gm.GPIPD(rh.FakeEnv(obs_dim=OBS, n_actions=A, reward_dim=D), batch_size=B, net_arch=[32, 32, 32], num_nets=2, gradient_updates=2,
                         dyna=False, per=True, gpi_pd=gpi_pd, drop_rate=0.0, layer_norm=True, buffer_size=N, log=False, seed=1, device="cpu",
                         target_net_update_freq=3)
        rng = np.random.default_rng(11)
        rb = agent.replay_buffer
        rb.obs[:] = rng.standard_normal((N, OBS)).astype(np.float32)
        rb.next_obs[:] = rng.standard_normal((N, OBS)).astype(np.float32)
        rb.actions[:] = rng.integers(0, A, size=(N, 1)).astype(np.uint8)
        rb.rewards[:] = rng.standard_normal((N, D)).astype(np.float32)
        rb.dones[:] = (rng.random((N, 1)) < 0.1).astype(np.float32)
        rb.size, rb.ptr = N, 0
        rb.tree.batch_set(np.arange(N), rng.random(N) + 0.01)
        support = rng.dirichlet(np.ones(D), 7).astype(np.float32)
        agent.set_weight_support(list(support))
        for i, net in enumerate(agent.q_nets):
            sd_to_npz(out, f"{tag}/init{i}", net.state_dict())
        for k in ("obs", "next_obs", "actions", "rewards", "dones"):
            out[f"{tag}/rb_{k}"] = getattr(rb, k).copy()
        out[f"{tag}/tree_leaves0"] = rb.tree.nodes[-1].copy()
        out[f"{tag}/support"] = support
        w = th.tensor(support[2])
        agent.global_step = agent.dynamics_rollout_starts  # use all `gradient_updates` (gpi_pd.py:419)
        agent.global_step = 3
        agent.dynamics_rollout_starts = 0
        random.seed(5)
        np.random.seed(6)
        for _ in range(2):
            agent.update(w)
            agent.global_step += 1
        for i, net in enumerate(agent.q_nets):
            sd_to_npz(out, f"{tag}/final{i}", net.state_dict())
        for i, net in enumerate(agent.target_q_nets):
            sd_to_npz(out, f"{tag}/final_target{i}", net.state_dict())
        out[f"{tag}/tree_leaves1"] = rb.tree.nodes[-1].copy()
        out[f"{tag}/min_priority1"] = np.float64(rb.min_priority)
        # gpi_action / eval on a few observations
        obs_eval = rng.standard_normal((8, OBS)).astype(np.float32)
        out[f"{tag}/eval_obs"] = obs_eval
        out[f"{tag}/eval_act"] = np.array([agent.eval(o, support[1]) for o in obs_eval], np.int32)
        # _reset_priorities
        agent._reset_priorities(w)
        out[f"{tag}/tree_leaves_reset"] = rb.tree.nodes[-1].copy()
        print(tag, "done")


class _NoiseStream:
    """Seeded standard-normal stream injected in place of Normal.rsample (CPU and CUDA generators differ)."""

    def __init__(self, seed):
        self.rng = np.random.default_rng(seed)

    def __call__(self, shape):
        return th.from_numpy(self.rng.standard_normal(tuple(shape)).astype(np.float32))


def _patch_rsample(stream):
    orig = th.distributions.Normal.rsample

    def rsample(self, sample_shape=th.Size()):
        eps = stream(self.loc.shape)
        return self.loc + eps * self.scale

    th.distributions.Normal.rsample = rsample
    return orig


def gen_capql(out):
    cm = rh.import_reference("morl_baselines.multi_policy.capql.capql")
    OBS, ACT, D, B = 9, 3, 2, 16
    th.manual_seed(0)
    env = rh.FakeEnv(obs_dim=OBS, continuous_action_dim=ACT, reward_dim=D)
    agent = cm.CAPQL(env, batch_size=B, net_arch=[32, 32], log=False, seed=2, device="cpu", gradient_updates=2)
    rng = np.random.default_rng(21)
    trans = []
    for _ in range(64):
        t = (rng.standard_normal(OBS).astype(np.float32), rng.uniform(-1, 1, ACT).astype(np.float32), rng.dirichlet(np.ones(D)).astype(np.float32),
             rng.standard_normal(D).astype(np.float32), rng.standard_normal(OBS).astype(np.float32), np.float32(rng.random() < 0.1))
        agent.replay_buffer.push(*t)
        trans.append(np.concatenate([np.asarray(x, np.float32).reshape(-1) for x in t]))
    out["capql/transitions"] = np.stack(trans)
    sd_to_npz(out, "capql/init_policy", agent.policy.state_dict())
    for i, q in enumerate(agent.q_nets):
        sd_to_npz(out, f"capql/init_q{i}", q.state_dict())
    orig = _patch_rsample(_NoiseStream(77))
    try:
        random.seed(9)
        agent.update()
    finally:
        th.distributions.Normal.rsample = orig
    sd_to_npz(out, "capql/final_policy", agent.policy.state_dict())
    for i, (q, tq) in enumerate(zip(agent.q_nets, agent.target_q_nets)):
        sd_to_npz(out, f"capql/final_q{i}", q.state_dict())
        sd_to_npz(out, f"capql/final_tq{i}", tq.state_dict())
    print("capql done")


def gen_mosac(out):
    mm = rh.import_reference("morl_baselines.single_policy.ser.mosac_continuous_action")
    OBS, ACT, D, B, N = 9, 3, 3, 16, 128
    th.manual_seed(0)
    env = rh.FakeEnv(obs_dim=OBS, continuous_action_dim=ACT, reward_dim=D)
    w = np.array([0.2, 0.5, 0.3], dtype=np.float32)
    agent = mm.MOSAC(env, weights=w, batch_size=B, net_arch=[32, 32], log=False, seed=4, device="cpu", buffer_size=N)
    rng = np.random.default_rng(31)
    buf = agent.buffer
    buf.obs[:] = rng.standard_normal((N, OBS)).astype(np.float32)
    buf.next_obs[:] = rng.standard_normal((N, OBS)).astype(np.float32)
    buf.actions[:] = rng.uniform(-1, 1, (N, ACT)).astype(np.float32)
    buf.rewards[:] = rng.standard_normal((N, D)).astype(np.float32)
    buf.dones[:] = (rng.random((N, 1)) < 0.1).astype(np.float32)
    buf.size, buf.ptr = N, 0
    for k in ("obs", "next_obs", "actions", "rewards", "dones"):
        out[f"mosac/rb_{k}"] = getattr(buf, k).copy()
    for name in ("actor", "qf1", "qf2"):
        sd_to_npz(out, f"mosac/init_{name}", getattr(agent, name).state_dict())
    orig = _patch_rsample(_NoiseStream(88))
    try:
        np.random.seed(12)
        for step in range(2):
            agent.global_step = 2 * step  # policy_freq = 2: actor updates on even steps
            agent.update()
    finally:
        th.distributions.Normal.rsample = orig
    for name in ("actor", "qf1", "qf2", "qf1_target", "qf2_target"):
        sd_to_npz(out, f"mosac/final_{name}", getattr(agent, name).state_dict())
    out["mosac/final_log_alpha"] = agent.log_alpha.detach().numpy().copy()
    print("mosac done")


def main():
    assert rh.reference_available()
    th.set_num_threads(os.cpu_count() or 1)
    out = {}
    gen_gpipd(out)
    gen_capql(out)
    gen_mosac(out)
    path = os.path.join(HERE, "updates.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path))


if __name__ == "__main__":
    main()
