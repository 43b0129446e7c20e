"""Pareto utilities backed by the warp-ballot dominance kernel (mirrors reference morl_baselines/common/pareto.py).

Signatures and return conventions are the reference's: numpy in, numpy out, output order = input order, comparisons exact
in the input dtype (float64 from rollouts, float32 kept as float32).  The O(N^2 d) all-pairs test runs on the GPU
(morl_pareto_mask_f32/_f64); there is no CPU fallback -- without a CUDA device these functions raise.
"""

from __future__ import annotations

from copy import deepcopy
from typing import List, Union

import numpy as np
import torch as th

from .. import ops


def pareto_dominates(a: np.ndarray, b: np.ndarray) -> np.bool_:
    """a weakly dominates b and is better somewhere (reference pareto.py:10-14)."""
    a, b = np.array(a), np.array(b)
    return np.all(a >= b) and np.any(a > b)


def strict_pareto_dominates(a: np.ndarray, b: np.ndarray) -> np.bool_:
    """a is better than b everywhere (reference pareto.py:17-21)."""
    return np.all(np.array(a) > np.array(b))


def batched_strict_pareto_dominates(p1: np.ndarray, p2: np.ndarray) -> np.ndarray:
    return np.all(p1 > p2, axis=-1)


def batched_pareto_dominates(p1: np.ndarray, p2: np.ndarray) -> np.ndarray:
    return np.logical_and(np.all(p1 >= p2, axis=-1), np.any(p1 > p2, axis=-1))


def _device():
    if not th.cuda.is_available():
        raise ops._lib.MorlB200Error("Pareto pruning needs a CUDA device (morl_baselines_b200 has no CPU fallback)")
    return th.device("cuda", th.cuda.current_device())


def get_non_pareto_dominated_inds(candidates: Union[np.ndarray, List], remove_duplicates: bool = True) -> np.ndarray:
    """Boolean mask of the points to keep (reference pareto.py:34-57), computed by the CUDA dominance kernel."""
    cand = np.array(candidates)
    if cand.ndim != 2:
        cand = cand.reshape(len(cand), -1)
    if cand.dtype not in (np.float32, np.float64):
        cand = cand.astype(np.float64)
    if len(cand) == 0:
        return np.zeros(0, dtype=bool)
    pts = th.from_numpy(np.ascontiguousarray(cand)).to(_device())
    return ops.pareto_mask(pts, remove_duplicates).cpu().numpy()


def filter_pareto_dominated(candidates: Union[np.ndarray, List], remove_duplicates: bool = True) -> np.ndarray:
    """Pareto coverage set in input order (reference pareto.py:60-73; fewer than two candidates are returned as is)."""
    cand = np.array(candidates)
    if len(cand) < 2:
        return cand
    return cand[get_non_pareto_dominated_inds(cand, remove_duplicates=remove_duplicates)]


def filter_convex_dominated(candidates: Union[np.ndarray, List]) -> np.ndarray:
    """Convex coverage set: QuickHull vertices (scipy / Qhull, as in the reference pareto.py:76-93), then the device prune."""
    from scipy.spatial import ConvexHull

    cand = np.array(candidates)
    ccs = cand[ConvexHull(cand).vertices] if len(cand) > 2 else cand
    return filter_pareto_dominated(ccs)


def get_non_dominated(candidates: set) -> set:
    """Non-dominated subset of a set of tuples (reference pareto.py:96-125)."""
    cand = np.array(list(candidates))
    if len(cand) == 0:
        return set()
    keep = get_non_pareto_dominated_inds(cand, remove_duplicates=True) if len(cand) > 1 else np.ones(1, dtype=bool)
    return {tuple(c) for c in cand[keep]}


def get_non_dominated_inds(solutions: np.ndarray) -> np.ndarray:
    """Boolean mask of non-dominated rows, duplicates kept (reference pareto.py:128-137)."""
    sol = np.asarray(solutions)
    if len(sol) < 2:
        return np.ones(len(sol), dtype=bool)
    return get_non_pareto_dominated_inds(sol, remove_duplicates=False)


class ParetoArchive:
    """Archive of non-dominated evaluations and the individuals that produced them (reference pareto.py:140-175)."""

    def __init__(self, convex_hull: bool = False):
        self.convex_hull = convex_hull
        self.individuals: list = []
        self.evaluations: List[np.ndarray] = []

    def add(self, candidate, evaluation: np.ndarray):
        """Append, re-filter the whole archive, rebuild both lists in insertion order with tuple de-duplication (reference pareto.py:149-175).
        The reference deep-copies the candidate BEFORE the dominance test and throws the copy away when the candidate is dominated or a
        duplicate; here the copy (a whole learner with its networks in MORL/D) is made only for a candidate that stays -- same archive."""
        evals_all = self.evaluations + [evaluation]
        if self.convex_hull:
            nd = {tuple(x) for x in filter_convex_dominated(evals_all)}
        else:
            nd = {tuple(x) for x in filter_pareto_dominated(evals_all)}
        evals, seen, inds = [], [], []
        for k, e in enumerate(evals_all):
            te = tuple(e)
            if te in nd and te not in seen:
                evals.append(e)
                seen.append(te)
                inds.append(self.individuals[k] if k < len(self.individuals) else deepcopy(candidate))
        self.evaluations = evals
        self.individuals = inds
