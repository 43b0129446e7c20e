"""CPU pinning of the oracle operators that the operator goldens do not cover directly: each test restates the cited reference lines with
the reference's own tensor library (PyTorch on CPU -- the arithmetic the reference executes) on seeded inputs and compares with
oracle/morl_oracle.c.  Targets and indices: bit-exact; reduced loss scalars: 1e-6 relative (summation order); gradients: bit-exact or
1 ulp where the reference's expression order is ambiguous (noted at the assert)."""

import numpy as np
import pytest
import torch as th
import torch.nn.functional as F

from oracle import oracle as orc


def _rng(seed):
    return np.random.default_rng(seed)


@pytest.mark.parametrize("n_nets,N,A,D", [(2, 64, 6, 3), (3, 33, 4, 2), (1, 17, 8, 3)])
def test_critic_min_target_restates_gpi_pd_445_463(n_nets, N, A, D):
    """gpi_pd.py:447-463: stack the target nets, scalarise with the per-row weight, argmin over nets, gather, scalarise, argmax over
    actions, gather, Bellman."""
    r = _rng(N + A)
    q = r.standard_normal((n_nets, N, A, D)).astype(np.float32)
    w = np.abs(r.standard_normal((N, D))).astype(np.float32)
    w /= w.sum(1, keepdims=True)
    rew, done = r.standard_normal((N, D)).astype(np.float32), (r.random(N) < 0.2).astype(np.float32)
    tq, tw = th.from_numpy(q), th.from_numpy(w)
    scal = th.einsum("nbar,br->nba", tq, tw)
    min_inds = th.argmin(scal, dim=0).reshape(1, N, A, 1).expand(1, N, A, D)
    next_q = tq.gather(0, min_inds).squeeze(0)
    max_q = th.einsum("bar,br->ba", next_q, tw)
    max_acts = th.argmax(max_q, dim=1)
    target = next_q.gather(1, max_acts.long().reshape(-1, 1, 1).expand(N, 1, D)).reshape(-1, D)
    target = th.from_numpy(rew) + (1 - th.from_numpy(done).reshape(-1, 1)) * 0.99 * target
    out, act = orc.critic_min_td(q, w, rew, done, 0.99)
    assert np.array_equal(act, max_acts.numpy().astype(np.int32))
    assert np.array_equal(out, target.numpy())


@pytest.mark.parametrize("D", [2, 3])
def test_actor_critic_targets_restate_capql_mosac_gpipd_continuous(D):
    r = _rng(D)
    n, N = 2, 96
    q = r.standard_normal((n, N, D)).astype(np.float32)
    w = r.dirichlet(np.ones(D), N).astype(np.float32)
    rew, done = r.standard_normal((N, D)).astype(np.float32), (r.random(N) < 0.2).astype(np.float32)
    logp = r.standard_normal(N).astype(np.float32)
    tq, tw, tr, td, tl = map(th.from_numpy, (q, w, rew, done, logp))
    # CAPQL, capql.py:329-331: per-objective min over the critics, minus alpha * logp, vector Bellman
    soft = th.min(tq, dim=0)[0] - (0.2 * tl).reshape(-1, 1)
    ref = (tr + (1 - td).reshape(-1, 1) * 0.99 * soft).numpy()
    assert np.array_equal(orc.actor_critic_td(q, None, rew, done, logp, 0.2, 0.99, orc.AC_ELEMENTWISE_MIN), ref)
    # GPI-PD continuous, gpi_pd_continuous_action.py:396-403: argmin_n w . Q_n, gather the winning critic's vector, vector Bellman
    inds = th.argmin(th.einsum("nbr,br->nb", tq, tw), dim=0, keepdim=True).reshape(1, -1, 1).expand(1, N, D)
    ref = (tr + (1 - td).reshape(-1, 1) * 0.99 * tq.gather(0, inds).squeeze(0)).numpy()
    assert np.array_equal(orc.actor_critic_td(q, w, rew, done, None, 0.0, 0.99, orc.AC_ARGMIN_GATHER), ref)
    # MOSAC, mosac_continuous_action.py:438-442 with a FIXED weight: min of the scalarised critics - alpha * logp, scalarised reward
    w1 = th.from_numpy(r.dirichlet(np.ones(D)).astype(np.float32))
    mn = th.min(th.matmul(tq[0], w1), th.matmul(tq[1], w1)) - (0.2 * tl)
    ref = (th.matmul(tr, w1) + (1 - td) * 0.99 * mn).numpy()
    got = orc.actor_critic_td(q, w1.numpy(), rew, done, logp, 0.2, 0.99, orc.AC_SCALAR_MIN)
    # th.matmul on [N, D] x [D] may use a different summation order than the oracle's left-to-right dot: 2 ulp of the magnitude
    np.testing.assert_allclose(got, ref, rtol=0, atol=4 * np.finfo(np.float32).eps * float(np.abs(ref).max() + 1))


@pytest.mark.parametrize("lam", [0.0, 0.3])
@pytest.mark.parametrize("order", [orc.ROWS_REFERENCE, orc.ROWS_BMAJOR])
def test_td_mse_loss_restates_envelope_301_313_and_330_333(lam, order):
    """envelope.py:301-313: gather the taken action, MSE (+ homotopy term on the scalarised values); :330-331 priorities of weight 0."""
    r = _rng(int(lam * 10) + order)
    B, W, A, D = 24, 5, 4, 3
    q = r.standard_normal((W * B, A, D)).astype(np.float32)
    t = r.standard_normal((W * B, D)).astype(np.float32)
    act = r.integers(0, A, B)
    wset = r.dirichlet(np.ones(D), W).astype(np.float32)
    # effective-batch row k: reference order k = i * B + b, b-major k = b * W + i
    rows = [(i, b) for i in range(W) for b in range(B)] if order == orc.ROWS_REFERENCE else [(i, b) for b in range(B) for i in range(W)]
    w_rows = th.from_numpy(np.stack([wset[i] for i, _ in rows]))
    a_rows = th.from_numpy(np.array([act[b] for _, b in rows]))
    tq = th.from_numpy(q).requires_grad_(True)
    q_value = tq.gather(1, a_rows.long().reshape(-1, 1, 1).expand(W * B, 1, D)).reshape(-1, D)
    loss = F.mse_loss(q_value, th.from_numpy(t))
    if lam > 0:
        loss = (1 - lam) * loss + lam * F.mse_loss(th.einsum("br,br->b", q_value, w_rows), th.einsum("br,br->b", th.from_numpy(t), w_rows))
    loss.backward()
    got_loss, got_grad, got_q, got_prio = orc.td_mse(q, act, t, wset, lam, B, W, order)
    assert abs(got_loss - float(loss.detach())) <= 1e-6 * abs(float(loss.detach()))
    assert np.array_equal(got_q, q_value.detach().numpy())
    np.testing.assert_allclose(got_grad, tq.grad.numpy(), rtol=2e-6, atol=1e-9)
    # priorities: |w_0 . (q - target)| of the B rows of weight index 0
    first = [k for k, (i, _) in enumerate(rows) if i == 0]
    per = th.einsum("br,br->b", (q_value.detach() - th.from_numpy(t))[first], w_rows[first]).abs().numpy()
    np.testing.assert_allclose(got_prio, per, rtol=0, atol=4 * np.finfo(np.float32).eps * float(np.abs(per).max() + 1))


@pytest.mark.parametrize("gpi", [False, True])
def test_td_huber_loss_restates_gpi_pd_469_520(gpi):
    """gpi_pd.py:469-487 per net: gather the taken action, huber(|td|, min_priority), mean over nets; :507-520 priorities from the
    element-wise max over nets of |td| (or |gtd| with gpi_pd), scalarised with the row weight."""
    r = _rng(7 + gpi)
    n, N, A, D, mp = 2, 40, 5, 3, 0.01
    q = r.standard_normal((n, N, A, D)).astype(np.float32) * 0.05
    t = r.standard_normal((N, D)).astype(np.float32) * 0.05
    tg = (t + 0.02 * r.standard_normal((N, D))).astype(np.float32) if gpi else None
    act = r.integers(0, A, N)
    w = r.dirichlet(np.ones(D), N).astype(np.float32)

    def huber(x, min_priority):  # common/networks.py:90-100
        return th.where(x < min_priority, 0.5 * x.pow(2), min_priority * x).mean()

    tq = th.from_numpy(q).requires_grad_(True)
    losses, errs = [], []
    for k in range(n):
        psi = tq[k].gather(1, th.from_numpy(act).long().reshape(-1, 1, 1).expand(N, 1, D)).squeeze(1)
        td = th.from_numpy(t) - psi
        losses.append(huber(td.abs(), mp))
        errs.append((th.from_numpy(tg) - psi).abs() if gpi else td.abs())
    loss = (1 / n) * sum(losses)
    loss.backward()
    per = th.einsum("br,br->b", th.max(th.stack(errs), dim=0)[0].detach(), th.from_numpy(w)).abs().numpy()
    got_loss, got_grad, got_prio = orc.td_huber(q, act, t, tg, w, mp, N)
    assert abs(got_loss - float(loss.detach())) <= 1e-6 * abs(float(loss.detach()))
    np.testing.assert_allclose(got_grad, tq.grad.numpy(), rtol=2e-6, atol=1e-10)
    np.testing.assert_allclose(got_prio, per, rtol=0, atol=4 * np.finfo(np.float32).eps * float(np.abs(per).max() + 1))
