"""GPU parity of GPI-PD's Dyna path against the UNMODIFIED reference (tests/golden/dyna.npz, frozen on CPU by
tests/golden/make_golden_dyna.py from common/model_based/probabilistic_ensemble.py, common/model_based/utils.py and
multi_policy/gpi_pd/gpi_pd.py:343-414) and against the numpy oracle (oracle/dyna_oracle.py).

Tolerances (floating point; the reference runs MKL fp32 GEMMs + numpy / torch-CPU exp, the engine cuBLAS fp32 GEMMs + CUDA expf):
ensemble means / logvars / samples / variances / uncertainties 1e-5 relative (+1e-6 absolute); which rows pass the uncertainty threshold,
the GPI actions, the termination flags and the buffer positions are discrete and must be IDENTICAL (the fixture's threshold sits in a
gap 800x wider than the arithmetic noise); parameters after training 1e-3 / 1e-5 (three epochs of Adam amplify GEMM-order noise)."""

import os
import random

import numpy as np
import pytest
import torch as th

from morl_baselines_b200.testing import FakeEnv, _Spec

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ENS = dict(OBS=6, A=4, D=3, E=5, ARCH=[64, 64], N=40)
DYN = dict(OBS=6, A=4, D=3, B=16, N=256, ENV_ID="mo-mountaincar-standin-v0", ROLLOUT_B=64, ROLLOUT_LEN=2, DYN_BUF=40, SEED_ROLLOUT=7, NOISE_SEED=43)


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(ROOT, "tests", "golden", "dyna.npz"))


def _load_sd(module, gold, prefix, dev):
    sd = {k[len(prefix) + 1:]: th.from_numpy(gold[k]).to(dev) for k in gold.files if k.startswith(prefix + "/")}
    module.load_state_dict(sd)


class _Noise:
    """The generator's TorchProxy.randn: standard normals from a seeded numpy stream, in call order."""

    def __init__(self, seed):
        self.rng = np.random.default_rng(seed)

    def __call__(self, shape, dev):
        return th.from_numpy(self.rng.standard_normal(tuple(shape)).astype(np.float32)).to(dev)


def test_ensemble_sample_kernel_matches_oracle(cuda):
    """morl_ensemble_sample_f32 against the numpy restatement on shapes that leave one warp partially idle, span several column strides and
    exercise both soft clamps, with and without noise / obs."""
    from morl_baselines_b200 import ops
    from oracle import dyna_oracle as do

    rng = np.random.default_rng(0)
    for E, N, O, rew in ((5, 1000, 35, 3), (3, 17, 70, 2), (7, 64, 9, 0), (1, 5, 4, 4)):
        out = (rng.standard_normal((E, N, 2 * O)) * 2).astype(np.float32)
        out[..., O:] = (rng.standard_normal((E, N, O)) * 6).astype(np.float32)  # raw logvars far beyond both bounds
        hi, lo = rng.uniform(-1, 0.5, O).astype(np.float32), rng.uniform(-7, -3, O).astype(np.float32)
        idx = rng.integers(0, E, N).astype(np.int32)
        noise = rng.standard_normal((E, N, O)).astype(np.float32)
        obs = rng.standard_normal((N, O - rew)).astype(np.float32)
        lv = do.clamp_logvar(out[..., O:], hi, lo)
        for use_noise in (False, True):
            for use_obs in (False, True):
                s, v, u = ops.ensemble_sample(th.from_numpy(out).to(cuda), th.from_numpy(hi).to(cuda), th.from_numpy(lo).to(cuda), th.from_numpy(idx).to(cuda),
                                              th.from_numpy(noise).to(cuda) if use_noise else None, th.from_numpy(obs).to(cuda) if use_obs else None, rew)
                so, vo, uo = do.ensemble_sample(out[..., :O], lv, idx, noise if use_noise else None, obs if use_obs else None, rew)
                np.testing.assert_allclose(s.cpu().numpy(), so, rtol=1e-5, atol=1e-6)
                np.testing.assert_allclose(v.cpu().numpy(), vo, rtol=1e-5, atol=1e-12)
                np.testing.assert_allclose(u.cpu().numpy(), uo, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("norm", [False, True])
def test_ensemble_forward_and_sample_match_reference(cuda, gold, norm):
    from morl_baselines_b200.common.model_based.probabilistic_ensemble import ProbabilisticEnsemble

    c, tag = ENS, f"ens{int(norm)}"
    m = ProbabilisticEnsemble(c["OBS"] + c["A"], c["OBS"] + c["D"], ensemble_size=c["E"], arch=c["ARCH"], num_elites=2, normalize_inputs=norm, device=cuda)
    _load_sd(m, gold, f"{tag}/init", cuda)
    x = th.from_numpy(gold[f"{tag}/x"]).to(cuda)
    with th.no_grad():
        mean, logvar = m.forward(x, deterministic=True, return_dist=True)
    np.testing.assert_allclose(mean.cpu().numpy(), gold[f"{tag}/mean"], rtol=1e-5, atol=2e-6)
    np.testing.assert_allclose(logvar.cpu().numpy(), gold[f"{tag}/logvar"], rtol=1e-5, atol=2e-6)
    m.elites = [3, 1]
    np.random.seed(3)
    s, v, u = m.sample(x, deterministic=True)
    np.testing.assert_allclose(s, gold[f"{tag}/det_sample"], rtol=1e-5, atol=2e-6)
    np.testing.assert_allclose(v, gold[f"{tag}/det_var"], rtol=1e-5, atol=1e-9)
    np.testing.assert_allclose(u, gold[f"{tag}/det_unc"], rtol=1e-5)
    m.noise_fn = _Noise(41)
    np.random.seed(4)
    s, v, u = m.sample(x, deterministic=False)
    np.testing.assert_allclose(s, gold[f"{tag}/sto_sample"], rtol=1e-5, atol=2e-6)
    np.testing.assert_allclose(v, gold[f"{tag}/sto_var"], rtol=1e-5, atol=1e-9)
    np.testing.assert_allclose(u, gold[f"{tag}/sto_unc"], rtol=1e-5)


def test_ensemble_fit_matches_reference(cuda, gold):
    """Three epochs of maximum-likelihood training: same bootstrap batches (numpy RNG consumed in the reference's order), same hold-out
    split, same elites; hold-out losses 1e-3, parameters 1e-3 / 1e-5."""
    from morl_baselines_b200.common.model_based.probabilistic_ensemble import ProbabilisticEnsemble

    c = ENS
    m = ProbabilisticEnsemble(c["OBS"] + c["A"], c["OBS"] + c["D"], ensemble_size=c["E"], arch=[32, 32], num_elites=2, normalize_inputs=True, device=cuda)
    _load_sd(m, gold, "fit/init", cuda)
    np.random.seed(5)
    mean_holdout = m.fit(gold["fit/X"], gold["fit/Y"], batch_size=64, max_epochs=3)
    assert mean_holdout == pytest.approx(float(gold["fit/mean_holdout"]), rel=1e-3)
    assert list(m.elites) == list(gold["fit/elites"])
    for k, v in m.state_dict().items():
        np.testing.assert_allclose(v.cpu().numpy(), gold[f"fit/final/{k}"], rtol=1e-3, atol=1e-5, err_msg=k)
    with th.no_grad():
        probe = m._compute_mse_losses(th.from_numpy(gold["fit/X"][:100]).to(cuda), th.from_numpy(gold["fit/Y"][:100]).to(cuda)).cpu().numpy()
    np.testing.assert_allclose(probe, gold["fit/holdout_probe"], rtol=1e-3)


def _build_agent(cuda, gold):
    from morl_baselines_b200.multi_policy.gpi_pd.gpi_pd import GPIPD

    c = DYN
    env = FakeEnv(obs_dim=c["OBS"], n_actions=c["A"], reward_dim=c["D"])
    env.spec = _Spec(c["ENV_ID"])
    agent = GPIPD(env, batch_size=c["B"], net_arch=[32, 32, 32], num_nets=2, gradient_updates=2, dyna=True, per=True, gpi_pd=True, drop_rate=0.0,
                  layer_norm=True, buffer_size=c["N"], log=False, seed=1, device=cuda, target_net_update_freq=3, dynamics_net_arch=[32, 32],
                  dynamics_rollout_batch_size=c["ROLLOUT_B"], dynamics_rollout_len=c["ROLLOUT_LEN"], dynamics_buffer_size=c["DYN_BUF"],
                  dynamics_uncertainty_threshold=float(gold["dyn/threshold"]), dynamics_rollout_starts=0, real_ratio=0.5)
    for i, (net, tnet) in enumerate(zip(agent.q_nets, agent.target_q_nets)):
        _load_sd(net, gold, f"dyn/init{i}", cuda)
        tnet.load_state_dict(net.state_dict())
    _load_sd(agent.dynamics, gold, "dyn/init_dynamics", cuda)
    agent.dynamics.elites = [4, 2]
    rb = agent.replay_buffer
    for k in ("obs", "next_obs", "actions", "rewards", "dones"):
        getattr(rb, k)[:] = gold[f"dyn/rb_{k}"]
    rb.size, rb.ptr = c["N"], 0
    rb.mark_all_dirty()
    rb.tree.batch_set(np.arange(c["N"]), gold["dyn/tree_leaves0"][: c["N"]])
    agent.set_weight_support(list(gold["dyn/support"]))
    return agent, th.tensor(gold["dyn/support"][2]).to(cuda)


def test_gpipd_dyna_rollout_sampling_and_update_match_reference(cuda, gold):
    """_rollout_dynamics (two model steps under the GPI policy, uncertainty filter, termination mask, wrapping bulk insert), the mixed
    real / imagined minibatch, and two updates on mixed minibatches, against the reference's buffers and parameters."""
    c = DYN
    agent, w = _build_agent(cuda, gold)
    assert agent.get_config()["dyna"] is True
    agent.dynamics.noise_fn = _Noise(c["NOISE_SEED"])
    np.random.seed(c["SEED_ROLLOUT"])
    added = agent._rollout_dynamics(w)
    db = agent.dynamics_buffer
    assert [db.ptr, db.size] == list(gold["dyn/db_ptr_size"]) and added > db.size  # (wrapped: more rows accepted than the buffer holds)
    assert np.array_equal(db.actions, gold["dyn/db_actions"]), "GPI actions / accepted rows differ"
    assert np.array_equal(db.dones, gold["dyn/db_dones"]), "termination flags differ"
    np.testing.assert_allclose(db.obs, gold["dyn/db_obs"], rtol=1e-5, atol=2e-6)
    np.testing.assert_allclose(db.next_obs, gold["dyn/db_next_obs"], rtol=1e-5, atol=5e-6)
    np.testing.assert_allclose(db.rewards, gold["dyn/db_rewards"], rtol=1e-5, atol=5e-6)
    # the HBM mirror of the dynamics buffer holds the same rows (the bulk insert writes both)
    dev_obs = db.device_stores()[0].cpu().numpy()
    assert np.array_equal(dev_obs, db.obs)
    # mixed minibatch: same real indices (PER tree walk), same imagined rows (np.random.choice on the dynamics buffer)
    agent.global_step = 3
    np.random.seed(8)
    obs, act, rew, nobs, done, idxes = agent._sample_batch_experiences()
    assert np.array_equal(np.asarray(idxes), gold["dyn/batch_idxes"])
    assert np.array_equal(act.cpu().numpy().reshape(-1), gold["dyn/batch_actions"].reshape(-1))
    np.testing.assert_allclose(obs.cpu().numpy(), gold["dyn/batch_obs"], rtol=1e-5, atol=2e-6)
    np.testing.assert_allclose(nobs.cpu().numpy(), gold["dyn/batch_next_obs"], rtol=1e-5, atol=5e-6)
    np.testing.assert_allclose(rew.cpu().numpy(), gold["dyn/batch_rewards"], rtol=1e-5, atol=5e-6)
    assert np.array_equal(done.cpu().numpy().reshape(-1), gold["dyn/batch_dones"].reshape(-1))
    # two updates on mixed minibatches
    random.seed(5)
    np.random.seed(6)
    for _ in range(2):
        agent.update(w)
        agent.global_step += 1
    for i, net in enumerate(agent.q_nets):
        for k, v in net.state_dict().items():
            np.testing.assert_allclose(v.detach().cpu().numpy(), gold[f"dyn/final{i}/{k}"], rtol=1e-4, atol=2e-6, err_msg=f"final{i}/{k}")
    np.testing.assert_allclose(agent.replay_buffer.tree.nodes[-1][: c["N"]], gold["dyn/tree_leaves1"][: c["N"]], rtol=2e-4, atol=1e-7)


def test_gpipd_dyna_train_iteration_smoke(cuda):
    """train_iteration with the reference's schedule hooks: the model is fitted, rolled out, and updates draw mixed minibatches."""
    from morl_baselines_b200.multi_policy.gpi_pd.gpi_pd import GPIPD

    env = FakeEnv(obs_dim=6, n_actions=4, reward_dim=3, horizon=20)
    env.spec = _Spec("mo-mountaincar-standin-v0")
    th.manual_seed(0)
    np.random.seed(0)
    random.seed(0)
    agent = GPIPD(env, batch_size=16, net_arch=[32, 32], gradient_updates=1, per=True, buffer_size=512, log=False, seed=0, device=cuda, learning_starts=30,
                  dynamics_net_arch=[32, 32], dynamics_train_freq=lambda t: 40, dynamics_rollout_starts=40, dynamics_rollout_freq=20,
                  dynamics_rollout_batch_size=32, dynamics_buffer_size=256, dynamics_uncertainty_threshold=1e9)
    support = [np.array([1.0, 0.0, 0.0], np.float32), np.array([0.0, 1.0, 0.0], np.float32), np.array([0.3, 0.3, 0.4], np.float32)]
    agent.train_iteration(90, support[2], support)
    assert agent.dyna and len(agent.dynamics_buffer) > 0 and agent._uses_model_samples()
    assert np.isfinite(float(agent._last_loss))


def test_ensemble_fit_graph_replay_matches_eager_steps(cuda, gold):
    """The captured minibatch step of ``fit`` (gather -> likelihood -> backward -> Adam as one CUDA-graph replay, capturable Adam) against
    the eager step on the same batches: same elites, parameters within 2e-5 / 1e-6 after three epochs (the capturable Adam forms
    lr / bias_correction on the device in float32; the update count is exact -- warm-up and capture passes are rolled back)."""
    from morl_baselines_b200.common.model_based import probabilistic_ensemble as pe

    c = ENS
    res = []
    saved = pe._FIT_GRAPH
    try:
        for graph in (False, True):
            pe._FIT_GRAPH = graph
            m = pe.ProbabilisticEnsemble(c["OBS"] + c["A"], c["OBS"] + c["D"], ensemble_size=c["E"], arch=[32, 32], num_elites=2, normalize_inputs=True, device=cuda)
            _load_sd(m, gold, "fit/init", cuda)
            np.random.seed(5)
            m.fit(gold["fit/X"], gold["fit/Y"], batch_size=64, max_epochs=3)
            res.append((list(m.elites), {k: v.detach().cpu().numpy().copy() for k, v in m.state_dict().items()},
                        [float(m.optim.state[p]["step"]) for g in m.optim.param_groups for p in g["params"]]))
    finally:
        pe._FIT_GRAPH = saved
    (e0, p0, s0), (e1, p1, s1) = res
    assert e0 == e1 and s0 == s1, "elites / number of optimiser steps differ"
    for k in p0:
        np.testing.assert_allclose(p1[k], p0[k], rtol=2e-5, atol=1e-6, err_msg=k)
