"""Scalarisation functions on numpy vectors (mirrors reference morl_baselines/common/scalarization.py)."""

import numpy as np


def weighted_sum(reward: np.ndarray, weights: np.ndarray) -> float:
    """Linear scalarisation (reference scalarization.py:7-17)."""
    return np.dot(reward, weights)


def tchebicheff(tau: float, reward_dim: int):
    """Adaptive Tchebycheff scalarisation (reference scalarization.py:20-41; the pymoo decomposition is restated:
    max_r w_r * |f_r - z_r| against the running utopian point)."""
    best_so_far = [float("-inf") for _ in range(reward_dim)]

    def thunk(reward: np.ndarray, weights: np.ndarray):
        for i, r in enumerate(reward):
            if best_so_far[i] < r + tau:
                best_so_far[i] = r + tau
        v = np.abs(np.asarray(reward, dtype=np.float64) - np.asarray(best_so_far)) * np.asarray(weights, dtype=np.float64)
        return -float(v.max())

    return thunk
