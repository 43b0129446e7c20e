"""Scalarisation functions on numpy vectors (mirrors reference morl_baselines/common/scalarization.py)."""

import numpy as np


def weighted_sum(reward: np.ndarray, weights: np.ndarray) -> float:
    """Linear scalarisation w . r (reference scalarization.py:7-17); fused into every device kernel of this package."""
    return np.dot(reward, weights)


def tchebicheff(tau: float, reward_dim: int):
    """Adaptive Tchebycheff scalarisation (reference scalarization.py:20-41).  The reference delegates to pymoo's decomposition; here it
    is restated directly: the utopian point z tracks max_t (r_t + tau) per objective and the utility is -max_r w_r |r_r - z_r|."""
    utopia = np.full(reward_dim, -np.inf)

    def scalarize(reward: np.ndarray, weights: np.ndarray) -> float:
        r = np.asarray(reward, dtype=np.float64)
        np.maximum(utopia, r + tau, out=utopia)
        return -float(np.max(np.abs(r - utopia) * np.asarray(weights, dtype=np.float64)))

    return scalarize
