"""Pin the PyTorch-CPU port of Envelope.update (oracle/envelope_update_port.py -- the CPU baseline and the whole-update
checker) against the UNMODIFIED reference.  Runs only where /root/reference is mounted (the build container)."""

import numpy as np
import pytest
import torch as th

from oracle import ref_harness as rh
from oracle.envelope_update_port import EnvelopeUpdatePort, synthetic_store

pytestmark = pytest.mark.skipif(not rh.reference_available(), reason="reference not mounted (GPU box)")


@pytest.mark.parametrize("per", [False, True])
def test_port_reproduces_reference_update(per):
    envm = rh.import_reference("morl_baselines.multi_policy.envelope.envelope")
    wm = rh.import_reference("morl_baselines.common.weights")
    OBS, A, D, W, B, N = 12, 5, 3, 6, 16, 512
    th.manual_seed(0)
    agent = envm.Envelope(rh.FakeEnv(obs_dim=OBS, n_actions=A, reward_dim=D), batch_size=B, num_sample_w=W, per=per, buffer_size=N,
                          net_arch=[32, 32], log=False, seed=3, device="cpu")
    store = synthetic_store(N, OBS, A, D, seed=1)
    rb = agent.replay_buffer
    rb.obs[:], rb.next_obs[:], rb.actions[:], rb.rewards[:], rb.dones[:] = (store[k] for k in ("obs", "next_obs", "actions", "rewards", "dones"))
    rb.size, rb.ptr = N, 0
    if per:
        rb.tree.batch_set(np.arange(N), np.full(N, rb.min_priority))
    port = EnvelopeUpdatePort(OBS, A, D, [32, 32], seed=0, state_dict=agent.q_net.state_dict())
    rng = np.random.default_rng(3)  # mirrors agent.np_random
    agent.global_step = 1
    for step in range(3):
        np.random.seed(50 + step)
        # replicate the reference's draws: replay indices from the global RNG first, then the weights from the agent's generator
        state = np.random.get_state()
        idx = rb.tree.sample(B) if per else np.random.choice(N, B, replace=True)
        np.random.set_state(state)
        wset = th.tensor(wm.random_weights(D, W, dist="gaussian", rng=rng)).float()
        agent.update()
        loss, prio = port.update(th.from_numpy(store["obs"][idx]), th.from_numpy(store["actions"][idx]), th.from_numpy(store["rewards"][idx]),
                                 th.from_numpy(store["next_obs"][idx]), th.from_numpy(store["dones"][idx]), wset)
    for (k, v), (k2, v2) in zip(agent.q_net.state_dict().items(), port.q_net.state_dict().items()):
        assert th.equal(v, v2), k  # same ops in the same order on the same machine: bit-identical
