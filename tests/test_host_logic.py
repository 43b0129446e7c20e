"""CPU-only tests of the host-side mirror of the reference interface: replay buffers, the PER sum-tree, weight
generation, indicators, schedules, and the N>1 front exchange over gloo (world_size 2)."""

import os
import pickle
import socket
import sys

import numpy as np
import pytest
import torch as th

from oracle import oracle as orc
from tests.golden import cases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_random_weights_consume_the_generator_like_the_reference(golden):
    from morl_baselines_b200.common.weights import extrema_weights, random_weights

    assert np.array_equal(random_weights(3, 64, "gaussian", rng=np.random.default_rng(5)), golden["random_weights_gauss"])
    assert np.array_equal(random_weights(4, 10, "dirichlet", rng=np.random.default_rng(5)), golden["random_weights_dir"])
    assert np.array_equal(random_weights(3, 1, "gaussian", rng=np.random.default_rng(9)), golden["random_weights_single"])
    assert np.array_equal(np.array(extrema_weights(3)), np.eye(3, dtype=np.float32))
    with pytest.raises(ValueError):
        random_weights(3, 2, "uniform")


@pytest.mark.parametrize("max_size,n0", [(1000, 700), (4096, 4096), (65536, 50000)])
def test_host_sumtree_matches_reference(golden, max_size, n0):
    from morl_baselines_b200.common.prioritized_buffer import SumTree

    tree = SumTree(max_size)
    rng = np.random.default_rng(max_size)
    tree.batch_set(np.arange(n0), rng.random(n0) + 1e-5)
    for rnd in range(4):
        np.random.seed(100 + rnd)
        idx = tree.sample(256)
        assert np.array_equal(idx, golden[f"sumtree_{max_size}_samples"][rnd])
        upd_idx = np.concatenate([idx, idx[:64]])
        tree.batch_set(upd_idx, rng.random(len(upd_idx)) * 3.0)
    assert tree.nodes[0][0] == float(golden[f"sumtree_{max_size}_root"])
    assert cases.digest(np.concatenate(tree.nodes)) == str(golden[f"sumtree_{max_size}_levels_sha"])


def test_replay_buffer_api_and_pickle_layout():
    from morl_baselines_b200.common.buffer import ReplayBuffer
    from morl_baselines_b200.common.prioritized_buffer import PrioritizedReplayBuffer

    rb = ReplayBuffer((5,), 1, rew_dim=3, max_size=16, action_dtype=np.uint8)
    for t in range(20):  # wraps around
        rb.add(np.full(5, t, np.float32), t % 4, np.full(3, -t, np.float32), np.full(5, t + 1, np.float32), t % 7 == 0)
    assert len(rb) == 16 and rb.ptr == 4
    assert rb.obs[3, 0] == 19 and rb.obs[4, 0] == 4  # newest / oldest
    np.random.seed(0)
    obs, act, rew, nobs, done, idx = rb.sample(8)
    np.random.seed(0)
    assert np.array_equal(idx, np.random.choice(16, 8, replace=True))
    assert np.array_equal(obs, rb.obs[idx]) and act.dtype == np.uint8 and rew.shape == (8, 3) and done.shape == (8, 1)
    _, _, _, _, _, idx = rb.sample(4, use_cer=True)
    assert idx[0] == rb.ptr - 1
    t_obs = rb.sample(4, to_tensor=True, device="cpu")[0]
    assert isinstance(t_obs, th.Tensor)
    assert len(rb.get_all_data()) == 5 and rb.get_all_data(max_samples=5)[0].shape == (5, 5)
    # pickling keeps the reference's numpy attribute layout
    rb2 = pickle.loads(pickle.dumps(rb))
    for name in ("obs", "next_obs", "actions", "rewards", "dones"):
        assert np.array_equal(getattr(rb2, name), getattr(rb, name))
    assert (rb2.ptr, rb2.size, rb2.max_size) == (rb.ptr, rb.size, rb.max_size)

    prb = PrioritizedReplayBuffer((5,), 1, rew_dim=3, max_size=16, action_dtype=np.uint8)
    for t in range(10):
        prb.add(np.zeros(5), 1, np.zeros(3), np.zeros(5), False)
    assert prb.tree.nodes[0][0] == pytest.approx(10 * 1e-5)
    prb.update_priorities(np.array([1, 1, 3]), np.array([0.5, 0.9, 0.2]))
    assert prb.tree.nodes[-1][1] == 0.5  # first occurrence wins
    assert prb.min_priority == 0.9  # ratchet (reference prioritized_buffer.py:194)
    prb.add(np.zeros(5), 1, np.zeros(3), np.zeros(5), False)
    assert prb.tree.nodes[-1][10] == 0.9
    np.random.seed(3)
    smp = prb.sample(32)
    assert smp[5].max() <= 10 and smp[0].shape == (32, 5)


def test_hypervolume_and_indicators():
    from morl_baselines_b200.common import performance_indicators as pi
    from oracle.hv_oracle import hypervolume_min

    assert pi.hypervolume(np.array([0.0, 0.0]), [np.array([1.0, 2.0]), np.array([2.0, 1.0])]) == pytest.approx(3.0)
    rng = np.random.default_rng(0)
    for d in (2, 3, 4):
        pts = rng.random((30, d)) * 5
        ref = -np.ones(d)
        assert pi.hypervolume(ref, list(pts)) == pytest.approx(hypervolume_min(-pts, -ref), rel=1e-12)
    front = [np.array([1.0, 0.0]), np.array([0.0, 1.0])]
    assert pi.cardinality(front) == 2
    assert pi.sparsity(front) == pytest.approx(2.0)
    assert pi.expected_utility(front, [np.array([1.0, 0.0]), np.array([0.5, 0.5])]) == pytest.approx(0.75)
    assert pi.igd(front, front) == 0.0
    assert pi.maximum_utility_loss(front[:1], front, np.array([[0.0, 1.0]])) == pytest.approx(1.0)


def test_schedules_and_unique_tol():
    from morl_baselines_b200.common.utils import linearly_decaying_value, nearest_neighbors, unique_tol

    assert linearly_decaying_value(1.0, 100, 0, 10, 0.1) == 1.0
    assert linearly_decaying_value(1.0, 100, 60, 10, 0.1) == pytest.approx(0.55)
    assert linearly_decaying_value(1.0, 100, 500, 10, 0.1) == pytest.approx(0.1)
    u = unique_tol([np.array([1.0, 0.5]), np.array([1.00001, 0.5]), np.array([0.0, 1.0])])  # rtol semantics of np.allclose
    assert len(u) == 2
    ws = [np.array([1.0, 0.0]), np.array([0.9, 0.1]), np.array([0.0, 1.0]), np.array([0.5, 0.5])]
    nn_ = nearest_neighbors(2, ws[0], ws, lambda a, b: float(np.abs(a - b).sum()))
    assert nn_ == [1, 3]


def test_envelope_requires_cuda():
    """The product path fails loudly without a CUDA device -- no CPU fallback."""
    from morl_baselines_b200 import _lib
    from morl_baselines_b200.multi_policy.envelope.envelope import Envelope
    from oracle.ref_harness import FakeEnv

    if th.cuda.is_available():
        pytest.skip("CPU-only check")
    with pytest.raises(_lib.MorlB200Error):
        Envelope(FakeEnv(), log=False, device="cpu")
    from morl_baselines_b200.common.pareto import filter_pareto_dominated

    with pytest.raises(_lib.MorlB200Error):
        filter_pareto_dominated(np.random.rand(5, 2))
    assert filter_pareto_dominated(np.random.rand(1, 2)).shape == (1, 2)  # < 2 candidates: returned as is (pareto.py:71-72)


def test_shard_range_and_front_records():
    from morl_baselines_b200.parallel import pack_front, shard_range, unpack_fronts

    spans = [shard_range(64, r, 8) for r in range(8)]
    assert spans[0] == (0, 8) and spans[-1] == (56, 64)
    spans = [shard_range(10, r, 4) for r in range(4)]
    assert [b - a for a, b in spans] == [3, 3, 2, 2] and spans[-1][1] == 10
    pts = th.arange(12, dtype=th.float64).view(4, 3)
    rec = pack_front(pts, cap=8, extras=th.tensor([7.0, 9.0]))
    allp, counts, ex = unpack_fronts(th.stack([rec, pack_front(pts[:1], 8, extras=th.tensor([1.0, 2.0]))]), 2, 8, 3, n_extra=2)
    assert counts.tolist() == [4, 1] and allp.shape == (16, 3) and int(th.isfinite(allp).all(dim=1).sum()) == 5
    assert ex.tolist() == [[7.0, 9.0], [1.0, 2.0]]
    rec = pack_front(pts, cap=2)  # overflow: the count is NOT clipped, only the rows are
    assert float(rec[0]) == 4 and rec.numel() == 1 + 2 * 3


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist

    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from morl_baselines_b200.parallel import allgather_fronts
    from oracle import oracle as orc_

    def prune(p):  # the CUDA kernel is not available on the CPU test host: inject the oracle as the dominance test
        return th.from_numpy(orc_.pareto_mask(p.numpy(), True))

    rng = np.random.default_rng(rank)
    pts = np.abs(rng.standard_normal((300, 3)))
    pts = th.from_numpy(pts / np.linalg.norm(pts, axis=1, keepdims=True))  # mutually non-dominated: 300 > cap forces grow-and-retry
    stats = {}
    front = allgather_fronts(pts, cap=64, prune=prune, stats=stats)
    q.put((rank, front.numpy(), stats))
    dist.destroy_process_group()


def test_allgather_fronts_gloo_world2():
    import torch.multiprocessing as mp

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, f0, s0), (_, f1, s1) = res
    assert np.array_equal(f0, f1)  # identical archive on every rank
    assert s0["rounds"] == 2 and s0["cap"] == 512 and s0["counts"] == [300, 300]  # overflow was detected, never truncated
    allpts = np.concatenate([np.abs(np.random.default_rng(r).standard_normal((300, 3))) for r in range(2)])
    allpts = allpts / np.linalg.norm(allpts, axis=1, keepdims=True)
    expect = allpts[orc.pareto_mask(allpts, True)]
    assert {tuple(r) for r in f0} == {tuple(r) for r in expect}


def _morld_eval_worker(rank, world, port, q):
    """Rank `rank` of a 2-rank MORL/D evaluation round on CPU / gloo: stub learners with fixed evaluations, the oracle as dominance test
    (the CUDA prune is not available on this host), the REAL MORLD._eval_all_policies / owner / local_policies / ParetoArchive logic."""
    import torch.distributed as dist

    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import morl_baselines_b200.common.pareto as pareto_mod
    from morl_baselines_b200.multi_policy.morld.morld import MORLD
    from oracle import oracle as orc_

    pareto_mod.get_non_pareto_dominated_inds = lambda c, remove_duplicates=True: orc_.pareto_mask(np.asarray(c, dtype=np.float64), remove_duplicates)
    calls = {"n": 0}
    real_gather, real_gather_list = dist.all_gather_into_tensor, dist.all_gather

    def counting(fn):
        def wrapped(*a, **k):
            calls["n"] += 1
            return fn(*a, **k)
        return wrapped

    dist.all_gather_into_tensor, dist.all_gather = counting(real_gather), counting(real_gather_list)

    class _Learner:
        def __init__(self, pid):
            self.pid = pid

        def policy_eval(self, eval_env, weights=None, scalarization=None, log=False):
            return None, None, None, _MORLD_EVALS[self.pid].copy()

    class _Pol:
        def __init__(self, pid):
            self.id, self.weights, self.wrapped = pid, np.array([0.5, 0.5]), _Learner(pid)

    algo = object.__new__(MORLD)
    algo.pop_size, algo.reward_dim, algo.rank, algo.world = len(_MORLD_EVALS), 2, rank, world
    algo.device, algo.log, algo.evaluation_mode, algo.scalarization = th.device("cpu"), False, "ser", None
    algo.population = [_Pol(i) for i in range(algo.pop_size)]
    algo.archive = pareto_mod.ParetoArchive()
    algo.global_front = None
    algo._front_prune = lambda p: th.from_numpy(orc_.pareto_mask(p.numpy(), True))
    evals = algo._eval_all_policies(None, 1, 5, np.zeros(2))
    q.put((rank, np.array(evals), algo.global_front, calls["n"], [p.id for p in algo.local_policies()], [int(i.id) for i in algo.archive.individuals]))
    dist.destroy_process_group()


_MORLD_EVALS = np.array([[1.0, 9.0], [2.0, 8.0], [2.0, 7.0], [5.0, 5.0], [4.0, 4.0], [9.0, 1.0], [8.0, 0.5]])


def test_morld_eval_round_gloo_world2():
    """MORLD._eval_all_policies on 2 ranks (reference morld.py:306-335 is the single-process form): policy p is evaluated by rank p % 2,
    every rank ends up with ALL evaluations and the SAME global front, and the round issues exactly ONE collective."""
    import torch.multiprocessing as mp

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_morld_eval_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    expect_front = _MORLD_EVALS[orc.pareto_mask(_MORLD_EVALS, True)]
    for rank, evals, front, n_coll, local_ids, arch_ids in res:
        assert np.array_equal(evals, _MORLD_EVALS)                       # everyone holds every policy's evaluation
        assert {tuple(r) for r in front} == {tuple(r) for r in expect_front}
        assert n_coll == 1                                               # one collective per evaluation round
        assert local_ids == [i for i in range(len(_MORLD_EVALS)) if i % 2 == rank]
        assert set(arch_ids) <= set(local_ids)                           # the local archive only holds the rank's own policies
    assert np.array_equal(res[0][2], res[1][2])                          # identical front on both ranks


class _NumpySumTree:
    """Restatement of the reference's SumTree (prioritized_buffer.py:12-82) with its own numpy calls, level arrays root first."""

    def __init__(self, max_size):
        self.nodes, level = [], 1
        for _ in range(int(np.ceil(np.log2(max_size))) + 1):
            self.nodes.append(np.zeros(level))
            level *= 2

    def walk(self, query):
        query = np.array(query, dtype=np.float64)
        node = np.zeros(len(query), dtype=int)
        for nodes in self.nodes[1:]:
            node *= 2
            left = nodes[node]
            greater = np.greater(query, left)
            node += greater
            query -= left * greater
        return node

    def batch_set(self, node_index, new_priority):
        node_index, unique_index = np.unique(node_index, return_index=True)
        diff = new_priority[unique_index] - self.nodes[-1][node_index]
        for nodes in self.nodes[::-1]:
            np.add.at(nodes, node_index, diff)
            node_index //= 2


@pytest.mark.parametrize("size", [5, 100, 4096, 100000])
def test_native_sum_tree_is_bit_identical_to_numpy_semantics(size):
    """The C sum tree (csrc/host_replay.cu) performs the reference's float64 operations in the reference's order: every level
    array and every sampled index is bit-identical, duplicates included (first occurrence wins)."""
    from morl_baselines_b200.common.prioritized_buffer import SumTree

    rng = np.random.default_rng(size)
    a, b = SumTree(size), _NumpySumTree(size)
    for _ in range(25):
        n = int(rng.integers(1, 1500))
        idx = rng.integers(0, size, n)
        pr = rng.random(n) * 10
        a.batch_set(idx, pr)
        b.batch_set(idx.copy(), pr)
        for la, lb in zip(a.nodes, b.nodes):
            assert np.array_equal(la, lb)
        q = rng.uniform(0, a.nodes[0][0], 777)
        assert np.array_equal(a.walk(q), b.walk(q))
    a.set(3 % size, 0.25)
    b.batch_set(np.array([3 % size]), np.array([0.25]))
    assert np.array_equal(a.nodes[0], b.nodes[0])
    import pickle

    c = pickle.loads(pickle.dumps(a))
    assert all(np.array_equal(x, y) for x, y in zip(a.nodes, c.nodes))
    c.batch_set(np.array([0]), np.array([1.5]))  # views still alias the flat array after unpickling
    assert c.nodes[-1][0] == 1.5 and c.nodes[0][0] != a.nodes[0][0]


def test_native_row_gather_matches_fancy_indexing():
    from morl_baselines_b200 import _lib

    lib = _lib.load()
    rng = np.random.default_rng(0)
    src = rng.standard_normal((500, 7)).astype(np.float32)
    idx = rng.integers(0, 500, 64).astype(np.int64)
    dst = np.empty((64, 7), np.float32)
    _lib.check(lib.morl_host_gather_rows(src.ctypes.data, src.strides[0], idx.ctypes.data, 64, dst.ctypes.data), "gather")
    assert np.array_equal(dst, src[idx])
    act = rng.integers(0, 255, (500, 1)).astype(np.uint8)
    d32 = np.empty((64, 1), np.int32)
    _lib.check(lib.morl_host_gather_u8_to_i32(act.ctypes.data, 1, idx.ctypes.data, 64, d32.ctypes.data), "gather u8")
    assert np.array_equal(d32, act[idx].astype(np.int32))
    assert lib.morl_host_gather_rows(None, 4, idx.ctypes.data, 1, dst.ctypes.data) == -1
