"""Golden vectors for GPI-PD's Dyna path, produced by the UNMODIFIED reference on CPU (run in the build container only):
      python tests/golden/make_golden_dyna.py   ->   tests/golden/dyna.npz

  * ProbabilisticEnsemble (common/model_based/probabilistic_ensemble.py): deterministic forward (mean, logvar), ``sample`` deterministic and
    stochastic (th.randn replaced by a seeded numpy stream: CPU and CUDA generators differ), ``fit`` for 3 epochs (hold-out losses, elites,
    parameters);
  * GPIPD(dyna=True) (multi_policy/gpi_pd/gpi_pd.py): ``_rollout_dynamics`` (:367-414; two model steps, termination rule "mountaincar" so that
    the non-terminal mask filters rows, an uncertainty threshold placed in the widest gap of the observed uncertainties so that CPU / GPU
    rounding cannot flip a row, a dynamics buffer small enough to wrap), ``_sample_batch_experiences`` (:343-365; real + imagined rows) and
    two ``update`` steps on mixed minibatches.
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


class TorchProxy:
    """Stands in for the name ``th`` inside a reference module: torch, except that randn comes from a seeded numpy stream."""

    def __init__(self, seed):
        self._rng = np.random.default_rng(seed)

    def randn(self, shape, device=None, **kw):
        return th.from_numpy(self._rng.standard_normal(tuple(shape)).astype(np.float32))

    def __getattr__(self, name):
        return getattr(th, name)


ENS = dict(OBS=6, A=4, D=3, E=5, ARCH=[64, 64], N=40)


def gen_ensemble(out):
    pm = rh.import_reference("morl_baselines.common.model_based.probabilistic_ensemble")
    c = ENS
    IN, OUT = c["OBS"] + c["A"], c["OBS"] + c["D"]
    rng = np.random.default_rng(0)
    for norm in (False, True):
        tag = f"ens{int(norm)}"
        th.manual_seed(0)
        m = pm.ProbabilisticEnsemble(IN, OUT, ensemble_size=c["E"], arch=c["ARCH"], num_elites=2, normalize_inputs=norm, device="cpu")
        if norm:
            m._fit_input_stats(rng.standard_normal((100, IN)).astype(np.float32) * 2 + 0.5)
        # non-trivial logvar bounds and output scale so that both soft clamps are active somewhere
        with th.no_grad():
            m.max_logvar.copy_(th.from_numpy(rng.uniform(-1.0, 0.5, (1, OUT)).astype(np.float32)))
            m.min_logvar.copy_(th.from_numpy(rng.uniform(-6.0, -3.0, (1, OUT)).astype(np.float32)))
            m.layers[-1].b.add_(th.from_numpy(rng.standard_normal((c["E"], 1, 2 * OUT)).astype(np.float32)))
        sd_to_npz(out, f"{tag}/init", m.state_dict())
        x = rng.standard_normal((c["N"], IN)).astype(np.float32)
        out[f"{tag}/x"] = x
        xt = th.from_numpy(x)
        with th.no_grad():
            mean, logvar = m.forward(xt, deterministic=True, return_dist=True)
        out[f"{tag}/mean"], out[f"{tag}/logvar"] = mean.numpy().copy(), logvar.numpy().copy()
        m.elites = [3, 1]
        np.random.seed(3)
        with th.no_grad():
            s, v, u = m.sample(xt, deterministic=True)
        out[f"{tag}/det_sample"], out[f"{tag}/det_var"], out[f"{tag}/det_unc"] = s.copy(), v.copy(), u.copy()
        pm.th = TorchProxy(41)
        try:
            np.random.seed(4)
            with th.no_grad():
                s, v, u = m.sample(xt, deterministic=False)
        finally:
            pm.th = th
        out[f"{tag}/sto_sample"], out[f"{tag}/sto_var"], out[f"{tag}/sto_unc"] = s.copy(), v.copy(), u.copy()
    # fit (normalised inputs): a noisy linear system
    th.manual_seed(1)
    m = pm.ProbabilisticEnsemble(IN, OUT, ensemble_size=c["E"], arch=[32, 32], num_elites=2, normalize_inputs=True, device="cpu")
    sd_to_npz(out, "fit/init", m.state_dict())
    Amat = rng.standard_normal((IN, OUT)).astype(np.float32) * 0.5
    X = rng.standard_normal((600, IN)).astype(np.float32)
    Y = (X @ Amat + 0.05 * rng.standard_normal((600, OUT))).astype(np.float32)
    out["fit/X"], out["fit/Y"] = X, Y
    np.random.seed(5)
    mean_holdout = m.fit(X, Y, batch_size=64, max_epochs=3)
    out["fit/mean_holdout"] = np.float64(mean_holdout)
    out["fit/elites"] = np.asarray(m.elites, np.int64)
    sd_to_npz(out, "fit/final", m.state_dict())
    with th.no_grad():
        out["fit/holdout_probe"] = m._compute_mse_losses(th.from_numpy(X[:100]), th.from_numpy(Y[:100])).numpy().copy()
    print("ensemble done; fit mean holdout", mean_holdout, "elites", m.elites)


DYN = dict(OBS=6, A=4, D=3, B=16, N=256, ENV_ID="mo-mountaincar-standin-v0", ROLLOUT_B=64, ROLLOUT_LEN=2, DYN_BUF=40, SEED_ROLLOUT=7, NOISE_SEED=43)


def build_ref_agent(gm, threshold):
    c = DYN
    env = rh.FakeEnv(obs_dim=c["OBS"], n_actions=c["A"], reward_dim=c["D"])
    env.spec = rh._Spec(c["ENV_ID"])
    th.manual_seed(0)
    agent = gm.GPIPD(env, batch_size=c["B"], net_arch=[32, 32, 32], num_nets=2, gradient_updates=2, dyna=True, per=True, gpi_pd=True, drop_rate=0.0,
                     layer_norm=True, buffer_size=c["N"], log=False, seed=1, device="cpu", target_net_update_freq=3, dynamics_net_arch=[32, 32],
                     dynamics_rollout_batch_size=c["ROLLOUT_B"], dynamics_rollout_len=c["ROLLOUT_LEN"], dynamics_buffer_size=c["DYN_BUF"],
                     dynamics_uncertainty_threshold=threshold, dynamics_rollout_starts=0, real_ratio=0.5)
    rng = np.random.default_rng(11)
    rb = agent.replay_buffer
    rb.obs[:] = rng.standard_normal((c["N"], c["OBS"])).astype(np.float32)
    rb.next_obs[:] = rng.standard_normal((c["N"], c["OBS"])).astype(np.float32)
    rb.actions[:] = rng.integers(0, c["A"], size=(c["N"], 1)).astype(np.uint8)
    rb.rewards[:] = rng.standard_normal((c["N"], c["D"])).astype(np.float32)
    rb.dones[:] = (rng.random((c["N"], 1)) < 0.1).astype(np.float32)
    rb.size, rb.ptr = c["N"], 0
    rb.tree.batch_set(np.arange(c["N"]), rng.random(c["N"]) + 0.01)
    support = rng.dirichlet(np.ones(c["D"]), 7).astype(np.float32)
    agent.set_weight_support(list(support))
    # a dynamics model with output spread: perturb the last layer so that the members disagree (uncertainty varies across rows)
    with th.no_grad():
        agent.dynamics.layers[-1].b.add_(th.from_numpy(rng.standard_normal(tuple(agent.dynamics.layers[-1].b.shape)).astype(np.float32)) * 0.3)
        agent.dynamics.layers[-1].W.mul_(3.0)          # members disagree more where the features are large: input-dependent uncertainty
        agent.dynamics.max_logvar.fill_(-3.0)          # small aleatoric part, so the ensemble disagreement decides the ranking
        agent.dynamics.min_logvar.fill_(-8.0)
    agent.dynamics.elites = [4, 2]
    return agent, support


def gen_gpipd_dyna(out):
    gm = rh.import_reference("morl_baselines.multi_policy.gpi_pd.gpi_pd")
    um = rh.import_reference("morl_baselines.common.model_based.utils")
    pm = rh.import_reference("morl_baselines.common.model_based.probabilistic_ensemble")
    c = DYN

    def rollout(agent, w):
        pm.th = TorchProxy(c["NOISE_SEED"])
        try:
            np.random.seed(c["SEED_ROLLOUT"])
            agent._rollout_dynamics(w)
        finally:
            pm.th = th

    # dry run with an infinite threshold: record the uncertainties of every model step, then place the threshold in their widest central gap
    seen = []
    orig_step = um.ModelEnv.step

    def recording_step(self, obs, act, deterministic=False):
        r = orig_step(self, obs, act, deterministic)
        seen.append(np.asarray(r[3]["uncertainty"]).copy())
        return r

    agent, support = build_ref_agent(gm, 1e30)
    w = th.tensor(support[2])
    gm.ModelEnv.step = recording_step
    try:
        rollout(agent, w)
    finally:
        gm.ModelEnv.step = orig_step
    allu = np.sort(np.concatenate(seen))
    lo, hi = int(0.35 * len(allu)), int(0.65 * len(allu))
    gaps = allu[lo + 1:hi] - allu[lo:hi - 1]
    k = lo + int(np.argmax(gaps))
    threshold = float(0.5 * (allu[k] + allu[k + 1]))
    print(f"uncertainties {allu[0]:.4f} .. {allu[-1]:.4f} ({len(allu)} rows over {len(seen)} steps); threshold {threshold:.6f} in a gap of {gaps.max():.2e}")
    assert gaps.max() > 1e-4 * threshold  # (CPU vs GPU arithmetic differs by ~1e-6 relative)

    agent, support = build_ref_agent(gm, threshold)
    for i, net in enumerate(agent.q_nets):
        sd_to_npz(out, f"dyn/init{i}", net.state_dict())
    sd_to_npz(out, "dyn/init_dynamics", agent.dynamics.state_dict())
    rb = agent.replay_buffer
    for k_ in ("obs", "next_obs", "actions", "rewards", "dones"):
        out[f"dyn/rb_{k_}"] = getattr(rb, k_).copy()
    out["dyn/tree_leaves0"] = rb.tree.nodes[-1].copy()
    out["dyn/support"] = support
    out["dyn/threshold"] = np.float64(threshold)
    w = th.tensor(support[2])
    rollout(agent, w)
    db = agent.dynamics_buffer
    for k_ in ("obs", "next_obs", "actions", "rewards", "dones"):
        out[f"dyn/db_{k_}"] = getattr(db, k_).copy()
    out["dyn/db_ptr_size"] = np.array([db.ptr, db.size], np.int64)
    print("dynamics buffer after the rollout: ptr", db.ptr, "size", db.size)
    # mixed minibatch
    agent.global_step = 3
    np.random.seed(8)
    batch = agent._sample_batch_experiences()
    for name, t in zip(("obs", "actions", "rewards", "next_obs", "dones", "idxes"), batch):
        out[f"dyn/batch_{name}"] = t.numpy().copy()
    # two updates on mixed minibatches
    random.seed(5)
    np.random.seed(6)
    for _ in range(2):
        agent.update(w)
        agent.global_step += 1
    for i, net in enumerate(agent.q_nets):
        sd_to_npz(out, f"dyn/final{i}", net.state_dict())
    out["dyn/tree_leaves1"] = rb.tree.nodes[-1].copy()
    print("gpipd dyna done")


def main():
    out = {}
    gen_ensemble(out)
    gen_gpipd_dyna(out)
    path = os.path.join(HERE, "dyna.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, len(out), "arrays", os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
