"""Performance indicators (mirrors reference morl_baselines/common/performance_indicators.py).

The reference delegates hypervolume / IGD to pymoo (not installed here, "parity unpinned" -- SURVEY.md 8(c)).  Hypervolume
is an exact quantity, so it is computed here by an exact dimension-sweep (HSO-style) algorithm in float64; it agrees with
any other exact algorithm to rounding, which is what the "hypervolume within 1 %" criterion needs.
"""

from copy import deepcopy
from typing import Callable, List

import numpy as np


def _hv_max(pts: np.ndarray, ref: np.ndarray) -> float:
    """Exact hypervolume dominated by `pts` above `ref` (maximisation), recursive slicing over the last objective."""
    if len(pts) == 0:
        return 0.0
    d = pts.shape[1]
    if d == 1:
        return float(pts[:, 0].max() - ref[0])
    order = np.argsort(-pts[:, -1])
    pts = pts[order]
    total = 0.0
    for i in range(len(pts)):
        lower = pts[i + 1, -1] if i + 1 < len(pts) else ref[-1]
        depth = pts[i, -1] - lower
        if depth > 0:
            total += depth * _hv_max(pts[: i + 1, :-1], ref[:-1])
    return total


def hypervolume(ref_point: np.ndarray, points: List) -> float:
    """Hypervolume of value vectors w.r.t. a reference point (reference performance_indicators.py:15-25)."""
    ref = np.asarray(ref_point, dtype=np.float64)
    pts = np.asarray(points, dtype=np.float64).reshape(-1, len(ref))
    pts = pts[np.all(pts > ref, axis=1)]
    return _hv_max(pts, ref)


def igd(known_front: List[np.ndarray], current_estimate: List[np.ndarray]) -> float:
    """Inverted generational distance (reference performance_indicators.py:28-39)."""
    ref = np.asarray(known_front, dtype=np.float64)
    cur = np.asarray(current_estimate, dtype=np.float64)
    d = np.linalg.norm(ref[:, None, :] - cur[None, :, :], axis=-1)
    return float(d.min(axis=1).mean())


def sparsity(front: List[np.ndarray]) -> float:
    """PGMORL sparsity (reference performance_indicators.py:42-68)."""
    if len(front) < 2:
        return 0.0
    arr = np.array(front)
    val = 0.0
    for dim in range(arr.shape[1]):
        objs = np.sort(deepcopy(arr.T[dim]))
        val += float(np.square(objs[1:] - objs[:-1]).sum())
    return val / (len(arr) - 1)


def expected_utility(front: List[np.ndarray], weights_set: List[np.ndarray], utility: Callable = np.dot) -> float:
    """Expected utility metric (reference performance_indicators.py:71-91)."""
    maxs = [np.max(np.array([utility(w, p) for p in front])) for w in weights_set]
    return np.mean(np.array(maxs), axis=0)


def cardinality(front: List[np.ndarray]) -> float:
    """Number of points of the front (reference performance_indicators.py:94-105)."""
    return len(front)


def maximum_utility_loss(front: List[np.ndarray], reference_set: List[np.ndarray], weights_set: np.ndarray,
                         utility: Callable = np.dot) -> float:
    """Maximum utility loss (reference performance_indicators.py:108-128)."""
    ref = [np.max([utility(w, p) for p in reference_set]) for w in weights_set]
    cur = [np.max([utility(w, p) for p in front]) for w in weights_set]
    return np.max([r - c for r, c in zip(ref, cur)])
