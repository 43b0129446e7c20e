"""Whole-update parity of GPIPD / GPILS, CAPQL and MOSAC (CUDA engine) against golden vectors produced by the UNMODIFIED
reference on CPU (tests/golden/make_golden_updates.py -> tests/golden/updates.npz).  Same initial parameters, same replay
contents, same RNG streams (python `random`, numpy global, injected Gaussian noise); tolerance 1e-4 relative / 2e-6 absolute
on parameters after the updates (Adam amplifies fp32 GEMM-order differences of ~1e-6 in the gradients), 1e-4 on priorities."""

import os
import random

import numpy as np
import pytest
import torch as th

from oracle.ref_harness import FakeEnv

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(ROOT, "tests", "golden", "updates.npz"))


def _load_sd(module, gold, prefix, dev):
    sd = {k[len(prefix) + 1:]: th.from_numpy(gold[k]).to(dev) for k in gold.files if k.startswith(prefix + "/")}
    module.load_state_dict(sd)


def _cmp_sd(module, gold, prefix, rtol=1e-4, atol=2e-6):
    for k, v in module.state_dict().items():
        np.testing.assert_allclose(v.detach().cpu().numpy(), gold[f"{prefix}/{k}"], rtol=rtol, atol=atol, err_msg=f"{prefix}/{k}")


class _Noise:
    def __init__(self, seed, dev):
        self.rng, self.dev = np.random.default_rng(seed), dev

    def __call__(self, shape):
        return th.from_numpy(self.rng.standard_normal(tuple(shape)).astype(np.float32)).to(self.dev)


@pytest.mark.parametrize("graph", [True, False])
@pytest.mark.parametrize("gpi_pd", [True, False])
def test_gpipd_update_matches_reference(cuda, gold, gpi_pd, graph):
    from morl_baselines_b200.multi_policy.gpi_pd.gpi_pd import GPIPD

    tag = f"gpipd{int(gpi_pd)}"
    OBS, A, D, B, N = 10, 4, 3, 16, 256
    agent = GPIPD(FakeEnv(obs_dim=OBS, n_actions=A, reward_dim=D), batch_size=B, net_arch=[32, 32, 32], num_nets=2, gradient_updates=2, dyna=False,
                  per=True, gpi_pd=gpi_pd, drop_rate=0.0, layer_norm=True, buffer_size=N, log=False, seed=1, device=cuda, target_net_update_freq=3,
                  use_cuda_graph=graph)
    for i, (net, tnet) in enumerate(zip(agent.q_nets, agent.target_q_nets)):
        _load_sd(net, gold, f"{tag}/init{i}", cuda)
        tnet.load_state_dict(net.state_dict())
    rb = agent.replay_buffer
    for k in ("obs", "next_obs", "actions", "rewards", "dones"):
        getattr(rb, k)[:] = gold[f"{tag}/rb_{k}"]
    rb.size, rb.ptr = N, 0
    rb.mark_all_dirty()
    rb.tree.batch_set(np.arange(N), gold[f"{tag}/tree_leaves0"][:N])
    support = gold[f"{tag}/support"]
    agent.set_weight_support(list(support))
    w = th.tensor(support[2]).to(cuda)
    agent.global_step = 3
    random.seed(5)
    np.random.seed(6)
    for _ in range(2):
        agent.update(w)
        agent.global_step += 1
    for i, (net, tnet) in enumerate(zip(agent.q_nets, agent.target_q_nets)):
        _cmp_sd(net, gold, f"{tag}/final{i}")
        _cmp_sd(tnet, gold, f"{tag}/final_target{i}")
    np.testing.assert_allclose(rb.tree.nodes[-1][:N], gold[f"{tag}/tree_leaves1"][:N], rtol=2e-4, atol=1e-7)
    assert rb.min_priority == pytest.approx(float(gold[f"{tag}/min_priority1"]), rel=2e-4)
    acts = np.array([agent.eval(o, support[1]) for o in gold[f"{tag}/eval_obs"]], np.int32)
    assert np.array_equal(acts, gold[f"{tag}/eval_act"])
    agent._reset_priorities(w)
    np.testing.assert_allclose(rb.tree.nodes[-1][:N], gold[f"{tag}/tree_leaves_reset"][:N], rtol=2e-4, atol=1e-7)


@pytest.mark.parametrize("graph", [True, False])
def test_capql_update_matches_reference(cuda, gold, graph):
    """``graph``: the CUDA-graph replay path (default) and the eager path run the same kernels; both are held to the reference."""
    from morl_baselines_b200.multi_policy.capql.capql import CAPQL

    OBS, ACT, D, B = 9, 3, 2, 16
    agent = CAPQL(FakeEnv(obs_dim=OBS, continuous_action_dim=ACT, reward_dim=D), batch_size=B, net_arch=[32, 32], log=False, seed=2, device=cuda,
                  gradient_updates=2, use_cuda_graph=graph)
    _load_sd(agent.policy, gold, "capql/init_policy", cuda)
    for i, (q, tq) in enumerate(zip(agent.q_nets, agent.target_q_nets)):
        _load_sd(q, gold, f"capql/init_q{i}", cuda)
        tq.load_state_dict(q.state_dict())
    for row in gold["capql/transitions"]:
        o = 0
        parts = []
        for n in (OBS, ACT, D, D, OBS, 1):
            parts.append(row[o:o + n])
            o += n
        agent.replay_buffer.push(parts[0], parts[1], parts[2], parts[3], parts[4], parts[5][0])
    agent._noise_hook = _Noise(77, cuda)
    random.seed(9)
    agent.update()
    _cmp_sd(agent.policy, gold, "capql/final_policy")
    for i, (q, tq) in enumerate(zip(agent.q_nets, agent.target_q_nets)):
        _cmp_sd(q, gold, f"capql/final_q{i}")
        _cmp_sd(tq, gold, f"capql/final_tq{i}")
    a = agent.eval(np.zeros(OBS, np.float32), np.array([0.5, 0.5], np.float32))
    assert a.shape == (ACT,) and np.all(np.abs(a) <= 1.0)


@pytest.mark.parametrize("graph", [True, False])
def test_mosac_update_matches_reference(cuda, gold, graph):
    from morl_baselines_b200.single_policy.ser.mosac_continuous_action import MOSAC

    OBS, ACT, D, B, N = 9, 3, 3, 16, 128
    w = np.array([0.2, 0.5, 0.3], dtype=np.float32)
    agent = MOSAC(FakeEnv(obs_dim=OBS, continuous_action_dim=ACT, reward_dim=D), weights=w, batch_size=B, net_arch=[32, 32], log=False, seed=4,
                  device=cuda, buffer_size=N, use_cuda_graph=graph)
    for name in ("actor", "qf1", "qf2"):
        _load_sd(getattr(agent, name), gold, f"mosac/init_{name}", cuda)
    agent.qf1_target.load_state_dict(agent.qf1.state_dict())
    agent.qf2_target.load_state_dict(agent.qf2.state_dict())
    buf = agent.buffer
    for k in ("obs", "next_obs", "actions", "rewards", "dones"):
        getattr(buf, k)[:] = gold[f"mosac/rb_{k}"]
    buf.size, buf.ptr = N, 0
    buf.mark_all_dirty()
    agent._noise_hook = _Noise(88, cuda)
    np.random.seed(12)
    for step in range(2):
        agent.global_step = 2 * step
        agent.update()
    for name in ("actor", "qf1", "qf2", "qf1_target", "qf2_target"):
        _cmp_sd(getattr(agent, name), gold, f"mosac/final_{name}")
    np.testing.assert_allclose(agent.log_alpha.detach().cpu().numpy(), gold["mosac/final_log_alpha"], rtol=1e-4, atol=1e-7)


def test_morld_population_smoke(cuda):
    """MORL/D with MOSAC learners on the stand-in MOMDP: one outer iteration (train, update the others, evaluate, archive)."""
    from morl_baselines_b200.multi_policy.morld.morld import MORLD

    env = FakeEnv(obs_dim=6, continuous_action_dim=2, reward_dim=2, horizon=20)
    eval_env = FakeEnv(obs_dim=6, continuous_action_dim=2, reward_dim=2, horizon=20, seed=1)
    algo = MORLD(env, pop_size=3, exchange_every=60, update_passes=2, log=False, device=cuda, seed=0, weight_init_method="random",
                 policy_args={"learning_starts": 20, "batch_size": 16, "net_arch": [32, 32], "buffer_size": 512}, neighborhood_size=1)
    algo.train(total_timesteps=60, eval_env=eval_env, ref_point=np.array([-100.0, -100.0]), num_eval_episodes_for_front=1, checkpoints=False)
    assert len(algo.archive.evaluations) >= 1 and algo.global_front.shape[1] == 2
    from morl_baselines_b200.common.performance_indicators import hypervolume

    assert hypervolume(np.array([-100.0, -100.0]), list(algo.global_front)) > 0


def test_morld_population_graph_equals_serial_updates(cuda):
    """MORL/D ``_update_others`` (reference morld.py:423-433: a strictly serial python loop over the policies): replaying a rank's learners
    as ONE multi-branch CUDA graph must leave every learner bit-identical to the serial per-policy replays -- same replay-index stream
    (global numpy RNG consumed in policy order), same injected noise, no cross-learner data flow."""
    from morl_baselines_b200.multi_policy.morld.morld import MORLD

    def build():
        th.manual_seed(0)  # (network initialisation draws from the global torch generator)
        env = FakeEnv(obs_dim=6, continuous_action_dim=2, reward_dim=2, horizon=20)
        algo = MORLD(env, pop_size=5, exchange_every=60, update_passes=3, log=False, device=cuda, seed=0, weight_init_method="random",
                     policy_args={"learning_starts": 0, "batch_size": 16, "net_arch": [32, 32], "buffer_size": 256}, neighborhood_size=1)
        rng = np.random.default_rng(5)
        for p in algo.population:
            buf = p.wrapped.get_buffer()
            for _ in range(64):
                buf.add(rng.standard_normal(6).astype(np.float32), rng.uniform(-1, 1, 2).astype(np.float32), rng.standard_normal(2).astype(np.float32),
                        rng.standard_normal(6).astype(np.float32), False)
            p.wrapped._noise_hook = _Noise(100 + p.id, cuda)
            p.wrapped.global_step = 4
        return algo

    a, b = build(), build()
    b.population_graph = False
    for algo in (a, b):
        np.random.seed(3)
        algo._update_others(algo.population[1])
    assert len(a._pop_graphs) == 1 and len(b._pop_graphs) == 0
    for pa, pb in zip(a.population, b.population):
        for name in ("actor", "qf1", "qf2", "qf1_target", "qf2_target"):
            for (k, va), (_, vb) in zip(getattr(pa.wrapped, name).state_dict().items(), getattr(pb.wrapped, name).state_dict().items()):
                assert th.equal(va, vb), (pa.id, name, k)
        assert th.equal(pa.wrapped.log_alpha, pb.wrapped.log_alpha)
    # the candidate itself was left alone
    ref = build().population[1].wrapped.actor.state_dict()
    for k, v in a.population[1].wrapped.actor.state_dict().items():
        assert th.equal(v, ref[k])


@pytest.mark.parametrize("graph", [True, False])
@pytest.mark.parametrize("tag,n_support", [("m5", 5), ("m1", 1)])
def test_gpipd_continuous_update_matches_reference(cuda, tag, n_support, graph):
    """GPILSContinuousAction: three critic steps + two delayed actor steps, PER write-back, target syncs and the GPI evaluation against
    the unmodified reference (tests/golden/make_golden_gpipd_continuous.py), hopper dimensions (BASELINE.json configs[2])."""
    from morl_baselines_b200.multi_policy.gpi_pd.gpi_pd_continuous_action import GPILSContinuousAction

    g = np.load(os.path.join(ROOT, "tests", "golden", "gpipd_continuous.npz"))
    OBS, ACT, D, B, N = 11, 3, 3, 16, 128
    agent = GPILSContinuousAction(FakeEnv(obs_dim=OBS, continuous_action_dim=ACT, reward_dim=D), batch_size=B, net_arch=[32, 32], num_q_nets=2,
                                  gradient_updates=3, per=True, buffer_size=N, log=False, seed=3, device=cuda, use_cuda_graph=graph)
    for net in agent.q_nets + agent.target_q_nets:
        for m in net.modules():
            if isinstance(m, th.nn.Dropout):
                m.p = 0.0
    _load_sd(agent.policy, g, f"{tag}/init_policy", cuda)
    agent.target_policy.load_state_dict(agent.policy.state_dict())
    for i, (q, tq) in enumerate(zip(agent.q_nets, agent.target_q_nets)):
        _load_sd(q, g, f"{tag}/init_q{i}", cuda)
        tq.load_state_dict(q.state_dict())
    rb = agent.replay_buffer
    for k in ("obs", "next_obs", "actions", "rewards", "dones"):
        getattr(rb, k)[:] = g[f"{tag}/rb_{k}"]
    rb.size, rb.ptr = N, 0
    rb.mark_all_dirty()
    rb.tree.batch_set(np.arange(N), g[f"{tag}/tree_leaves0"][:N])
    support = g[f"{tag}/support"]
    assert support.shape[0] == n_support
    agent.set_weight_support(list(support))
    w = th.tensor(support[min(2, n_support - 1)]).to(cuda)
    agent._noise_hook = _Noise(99, cuda)
    random.seed(15)
    np.random.seed(16)
    agent.global_step = 5
    agent.update(w)
    _cmp_sd(agent.policy, g, f"{tag}/final_policy")
    _cmp_sd(agent.target_policy, g, f"{tag}/final_target_policy")
    for i, (q, tq) in enumerate(zip(agent.q_nets, agent.target_q_nets)):
        _cmp_sd(q, g, f"{tag}/final_q{i}")
        _cmp_sd(tq, g, f"{tag}/final_tq{i}")
    np.testing.assert_allclose(rb.tree.nodes[-1][:N], g[f"{tag}/tree_leaves1"][:N], rtol=2e-4, atol=1e-7)
    assert rb.min_priority == pytest.approx(float(g[f"{tag}/min_priority1"]), rel=2e-4)
    wq = support[0]
    agent.use_gpi = False
    plain = np.stack([agent.eval(o, wq) for o in g[f"{tag}/eval_obs"]])
    np.testing.assert_allclose(plain, g[f"{tag}/eval_plain"], rtol=1e-4, atol=1e-5)
    agent.use_gpi = True
    gpi = np.stack([agent.eval(o, wq) for o in g[f"{tag}/eval_obs"]])
    np.testing.assert_allclose(gpi, g[f"{tag}/eval_gpi"], rtol=1e-4, atol=1e-5)


def test_gpipd_continuous_train_iteration_and_checkpoint(cuda, tmp_path):
    """train_iteration on the stand-in MOMDP (rollout -> replay -> update), save / load round trip with the reference's keys."""
    from morl_baselines_b200.multi_policy.gpi_pd.gpi_pd_continuous_action import GPIPDContinuousAction, GPILSContinuousAction

    env = FakeEnv(obs_dim=11, continuous_action_dim=3, reward_dim=3, horizon=20)
    with pytest.raises(NotImplementedError):
        GPIPDContinuousAction(env, log=False, device=cuda)  # dyna=True is the reference default; the Dyna path is out of scope
    agent = GPILSContinuousAction(env, batch_size=16, net_arch=[32, 32], gradient_updates=2, learning_starts=20, buffer_size=512, log=False,
                                  seed=0, device=cuda)
    M = [np.array([1.0, 0.0, 0.0], np.float32), np.array([0.0, 1.0, 0.0], np.float32), np.array([0.3, 0.3, 0.4], np.float32)]
    agent.train_iteration(total_timesteps=60, weight=M[2], weight_support=M, change_weight_every_episode=True)
    assert agent.global_step == 60 and agent._n_updates == 2 * 41 and len(agent.replay_buffer) == 60
    assert all(bool(th.isfinite(p).all()) for p in agent.policy.parameters())
    agent.save(save_dir=str(tmp_path), filename="ckpt", save_replay_buffer=False)
    params = th.load(str(tmp_path / "ckpt.tar"), map_location="cpu", weights_only=False)
    assert {"policy_state_dict", "policy_optimizer_state_dict", "q_net_0_state_dict", "target_q_net_0_state_dict", "q_net_1_state_dict",
            "target_q_net_1_state_dict", "q_nets_optimizer_state_dict", "M"} <= set(params)
    other = GPILSContinuousAction(env, batch_size=16, net_arch=[32, 32], buffer_size=512, log=False, seed=1, device=cuda)
    other.load(str(tmp_path / "ckpt.tar"))
    a1 = agent.eval(np.ones(11, np.float32), M[2])
    a2 = other.eval(np.ones(11, np.float32), M[2])
    np.testing.assert_array_equal(a1, a2)


def test_mosac_graph_replay_matches_eager(cuda):
    """The CUDA-graph path of MOSAC.update (gather + critic step + actor / temperature steps + target syncs in one replay) and the
    eager path apply the same update: after 6 updates from the same state with the same injected noise the parameters agree to the
    tolerance of the golden-vector tests (1e-4 relative / 2e-6 absolute) (the library GEMMs may pick a different algorithm under capture, where no workspace can be allocated, and Adam amplifies
    the last-bit differences), and the number of applied updates is exact (warm-up / capture passes leave no trace)."""
    from morl_baselines_b200.single_policy.ser.mosac_continuous_action import MOSAC

    OBS, ACT, D, B, N = 11, 3, 3, 32, 256
    w = np.array([0.2, 0.5, 0.3], dtype=np.float32)
    agents = []
    for graph in (True, False):
        th.manual_seed(7)
        a = MOSAC(FakeEnv(obs_dim=OBS, continuous_action_dim=ACT, reward_dim=D), weights=w, batch_size=B, net_arch=[64, 64], log=False, seed=4,
                  device=cuda, buffer_size=N, use_cuda_graph=graph)
        rng = np.random.default_rng(5)
        buf = a.buffer
        buf.obs[:] = rng.standard_normal((N, OBS)).astype(np.float32)
        buf.next_obs[:] = rng.standard_normal((N, OBS)).astype(np.float32)
        buf.actions[:] = rng.uniform(-1, 1, (N, ACT)).astype(np.float32)
        buf.rewards[:] = rng.standard_normal((N, D)).astype(np.float32)
        buf.dones[:] = (rng.random((N, 1)) < 0.1).astype(np.float32)
        buf.size, buf.ptr = N, 0
        buf.mark_all_dirty()
        a._noise_hook = _Noise(123, cuda)
        np.random.seed(3)
        for step in range(6):
            a.global_step = step  # policy_freq = 2: actor + temperature steps on even steps (two graph variants)
            a.update()
        agents.append(a)
    g, e = agents
    assert len(g._graphs) == 2 and len(e._graphs) == 0
    for name in ("actor", "qf1", "qf2", "qf1_target", "qf2_target"):
        for (k, pg), (_, pe) in zip(getattr(g, name).state_dict().items(), getattr(e, name).state_dict().items()):
            np.testing.assert_allclose(pg.cpu().numpy(), pe.cpu().numpy(), rtol=1e-4, atol=2e-6, err_msg=f"{name}.{k}")
    np.testing.assert_allclose(g.log_alpha.detach().cpu().numpy(), e.log_alpha.detach().cpu().numpy(), rtol=1e-4, atol=1e-6)
    assert g.alpha == pytest.approx(e.alpha, rel=1e-4)
    for og, oe in ((g.q_optimizer, e.q_optimizer), (g.actor_optimizer, e.actor_optimizer), (g.a_optimizer, e.a_optimizer)):
        sg, se = [s["step"].item() for s in og.state.values()], [s["step"].item() for s in oe.state.values()]
        assert sg == se and len(sg) > 0
    assert [s["step"].item() for s in g.q_optimizer.state.values()][0] == 6.0
