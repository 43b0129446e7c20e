"""Deterministic synthetic inputs shared by the golden-vector generator (make_golden.py, run in the build container
against the real reference) and by the parity tests (run anywhere).  Everything derives from numpy's PCG64 streams,
which are bit-reproducible across platforms, so the fixtures only need to store seeds + outputs (or output digests).
"""

from __future__ import annotations

import hashlib

import numpy as np


def digest(a: np.ndarray) -> str:
    a = np.ascontiguousarray(a)
    return hashlib.sha256(a.tobytes()).hexdigest()


def gaussian_weights(rng: np.random.Generator, n: int, d: int) -> np.ndarray:
    """Same arithmetic as reference common/weights.py:27-28 ("gaussian"), float64 -> float32."""
    w = rng.standard_normal((n, d))
    w = np.abs(w) / np.linalg.norm(w, ord=1, axis=1, keepdims=True)
    return w.astype(np.float32)


# (name, B, W, A, D, ties)
ENVELOPE_CASES = [
    ("default_w4", 256, 4, 6, 3, False),  # reference default num_sample_w=4 (envelope.py:105), minecart |A|=6
    ("small", 64, 8, 8, 3, False),
    ("ragged", 37, 5, 7, 3, False),  # nothing a multiple of 4/32
    ("d2", 96, 16, 4, 2, False),
    ("d4", 48, 12, 5, 4, False),
    ("ties", 128, 16, 8, 3, True),  # quantised Q and weights: exact ties exercise first-occurrence
    ("config2", 256, 32, 6, 3, False),  # BASELINE.json configs[1]: minecart, |W|=32, batch 256
    ("north_star", 1024, 64, 8, 3, False),  # metric shape
]


def envelope_inputs(name: str):
    spec = {c[0]: c for c in ENVELOPE_CASES}[name]
    _, B, W, A, D, ties = spec
    seed = int.from_bytes(hashlib.sha256(("envelope:" + name).encode()).digest()[:4], "little")
    rng = np.random.default_rng(seed)
    if ties:
        q_on = rng.integers(-2, 3, size=(B, W, A, D)).astype(np.float32)
        q_tg = rng.integers(-2, 3, size=(B, W, A, D)).astype(np.float32)
        wset = rng.integers(1, 4, size=(W, D)).astype(np.float32)
        wset = (wset / 8.0).astype(np.float32)  # exact binary fractions: products and sums are exact -> true ties
    else:
        q_on = (rng.standard_normal((B, W, A, D)) * 3.0).astype(np.float32)
        q_tg = (q_on + 0.05 * rng.standard_normal((B, W, A, D))).astype(np.float32)
        wset = gaussian_weights(rng, W, D)
    reward = rng.standard_normal((B, D)).astype(np.float32)
    done = (rng.random(B) < 0.1).astype(np.float32)
    gamma = 0.99
    return dict(B=B, W=W, A=A, D=D, q_on=q_on, q_tg=q_tg, wset=wset, reward=reward, done=done, gamma=gamma)


# (name, n_nets, B, P, A, D, ties)
GPI_CASES = [
    ("gpi_small", 2, 64, 5, 6, 3, False),  # sampled_w = [weight] + 4 support weights (gpi_pd.py:440-441)
    ("gpi_support64", 2, 128, 64, 8, 3, False),  # |M| = 64
    ("gpi_one_net", 1, 33, 7, 5, 3, False),
    ("gpi_three_nets", 3, 40, 9, 4, 2, False),
    ("gpi_ties", 2, 64, 8, 8, 3, True),
]


def gpi_inputs(name: str):
    spec = {c[0]: c for c in GPI_CASES}[name]
    _, n_nets, B, P, A, D, ties = spec
    seed = int.from_bytes(hashlib.sha256(("gpi:" + name).encode()).digest()[:4], "little")
    rng = np.random.default_rng(seed)
    if ties:
        q = rng.integers(-2, 3, size=(n_nets, B, P, A, D)).astype(np.float32)
        w = (rng.integers(1, 4, size=(B, D)) / 8.0).astype(np.float32)
    else:
        base = rng.standard_normal((1, B, P, A, D)) * 2.0
        q = (base + 0.3 * rng.standard_normal((n_nets, B, P, A, D))).astype(np.float32)
        w = gaussian_weights(rng, B, D)
    reward = rng.standard_normal((B, D)).astype(np.float32)
    done = (rng.random(B) < 0.1).astype(np.float32)
    return dict(n_nets=n_nets, B=B, P=P, A=A, D=D, q=q, w=w, reward=reward, done=done, gamma=0.99)


def pareto_points(name: str):
    """Point sets for the Pareto mask; 'unit_ball_*' follow the constructions of the reference's known-answer tests
    (reference tests/test_pruning.py:25-65: non-dominated points on the positive unit sphere x10, dominated points
    obtained by shrinking a non-dominated point) -- re-derived here, not copied."""
    seed = int.from_bytes(hashlib.sha256(("pareto:" + name).encode()).digest()[:4], "little")
    rng = np.random.default_rng(seed)
    if name.startswith("unit_ball"):
        _, _, d, n_nd, n_dom = name.split("_")
        d, n_nd, n_dom = int(d[1:]), int(n_nd), int(n_dom)
        x = np.abs(rng.standard_normal((n_nd, d)))
        nd = 10.0 * x / np.linalg.norm(x, axis=1, keepdims=True)
        picks = rng.integers(0, n_nd, size=n_dom)
        shrink = rng.uniform(0.1, 0.95, size=(n_dom, 1))
        dom = nd[picks] * shrink
        pts = np.concatenate([nd, dom], axis=0)
        perm = rng.permutation(len(pts))
        return pts[perm]
    if name == "dups_int":
        return rng.integers(0, 6, size=(400, 3)).astype(np.float64)
    if name == "dups_int_f32":
        return rng.integers(0, 5, size=(300, 2)).astype(np.float32)
    if name == "random_f32_d4":
        return rng.standard_normal((700, 4)).astype(np.float32)
    if name == "random_f64_d3":
        return rng.standard_normal((513, 3))
    if name == "single_dim":
        return rng.integers(0, 9, size=(50, 1)).astype(np.float64)
    if name == "appendix_a5":
        return np.array([[1, 2], [2, 1], [1, 2], [0, 0], [2, 1], [1, 1], [2, 2], [2, 2]], dtype=np.float64)
    raise KeyError(name)


PARETO_CASES = [
    "appendix_a5",
    "unit_ball_d2_100_500",  # the shape of reference tests/test_pruning.py:74-86
    "unit_ball_d4_1000_5000",  # the shape of reference tests/test_pruning.py:102-114
    "dups_int",
    "dups_int_f32",
    "random_f32_d4",
    "random_f64_d3",
    "single_dim",
]


def archive_sequence():
    """Evaluations fed one by one to ParetoArchive.add (reference common/pareto.py:149-175): a coarse integer grid, so later points
    duplicate or dominate earlier ones and the archive shrinks as well as grows."""
    rng = np.random.default_rng(77)
    seq = []
    while len(seq) < 120:
        p = rng.integers(0, 10, size=3)
        if p.sum() <= 14:  # a budget: the non-dominated set is the sum ~ 14 band, not a single corner
            seq.append(p.astype(np.float64))
    return seq
