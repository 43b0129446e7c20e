"""General host-side helpers (mirrors reference morl_baselines/common/utils.py)."""

import math
from typing import Callable, List

import numpy as np


def linearly_decaying_value(initial_value, decay_period, step, warmup_steps, final_value):
    """Nature-DQN style linear schedule (reference utils.py:10-32)."""
    steps_left = decay_period + warmup_steps - step
    bonus = (initial_value - final_value) * steps_left / decay_period
    value = final_value + bonus
    return np.clip(value, min(initial_value, final_value), max(initial_value, final_value))


def unique_tol(a: List[np.ndarray], tol=1e-4) -> List[np.ndarray]:
    """Unique elements of a list of arrays within a tolerance, first occurrence kept (reference utils.py:35-47)."""
    if len(a) == 0:
        return a
    arr = np.array(a)
    delete = np.zeros(len(arr), dtype=bool)
    for i in range(len(arr)):
        if delete[i]:
            continue
        for j in range(i + 1, len(arr)):
            if np.allclose(arr[i], arr[j], tol):
                delete[j] = True
    return list(arr[~delete])


def nearest_neighbors(n: int, current_weight: np.ndarray, all_weights: List[np.ndarray],
                      dist_metric: Callable[[np.ndarray, np.ndarray], float]) -> List[int]:
    """Indices of the n nearest distinct weight vectors (reference utils.py:71-107)."""
    assert n < len(all_weights)
    cur = tuple(current_weight)
    chosen, chosen_ids = [], []
    while len(chosen_ids) < n:
        best_id, best, best_d = -1, None, math.inf
        for i, w in enumerate(all_weights):
            wt = tuple(w)
            if wt not in chosen and cur != wt:
                d = dist_metric(current_weight, w)
                if best_d > d:
                    best_id, best, best_d = i, wt, d
        chosen.append(best if best is not None else tuple(np.zeros_like(current_weight)))
        chosen_ids.append(best_id)
    return chosen_ids
