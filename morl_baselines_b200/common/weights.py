"""Weight-vector utilities (host side; mirrors reference morl_baselines/common/weights.py).

``random_weights`` consumes the numpy Generator exactly like the reference (weights.py:10-35), so an agent seeded like the
reference draws bit-identical weight sets.  ``equally_spaced_weights`` needs pymoo's Riesz s-energy reference directions
(weights.py:38-49; third-party, "parity unpinned" -- SURVEY.md 8(c)); when pymoo is absent an in-repo Riesz-energy
descent on the simplex is used instead, and both engines of a parity run must be fed the same list.
"""

from __future__ import annotations

from functools import lru_cache
from typing import List, Optional

import numpy as np


def random_weights(dim: int, n: int = 1, dist: str = "dirichlet", seed: Optional[int] = None,
                   rng: Optional[np.random.Generator] = None) -> np.ndarray:
    """Random normalised weight vectors: |N(0,1)| / L1 ("gaussian") or Dirichlet(1) (reference weights.py:10-35)."""
    if rng is None:
        rng = np.random.default_rng(seed)
    if dist == "gaussian":
        w = rng.standard_normal((n, dim))
        w = np.abs(w) / np.linalg.norm(w, ord=1, axis=1, keepdims=True)
    elif dist == "dirichlet":
        w = rng.dirichlet(np.ones(dim), n)
    else:
        raise ValueError(f"Unknown distribution {dist}")
    return w[0] if n == 1 else w


def _riesz_energy_directions(dim: int, n: int, seed: int, iters: int = 400) -> np.ndarray:
    """Minimise the Riesz s-energy (s = dim^2) of n points on the unit simplex by projected gradient descent."""
    rng = np.random.default_rng(seed)
    x = rng.dirichlet(np.ones(dim), n)
    ext = np.eye(dim)[: min(dim, n)]
    x[: len(ext)] = ext  # keep the extrema, as pymoo's "energy" method does
    free = np.ones(n, dtype=bool)
    free[: len(ext)] = False
    s = float(dim * dim)
    lr = 1e-3
    for _ in range(iters):
        diff = x[:, None, :] - x[None, :, :]
        dist = np.linalg.norm(diff, axis=-1) + np.eye(n)
        g = (-s * diff / dist[..., None] ** (s + 2)).sum(axis=1)
        g = g - g.mean(axis=1, keepdims=True)  # tangent to sum(x) = 1
        gn = np.linalg.norm(g, axis=1, keepdims=True) + 1e-12
        step = lr * g / gn
        x[free] = x[free] - step[free]
        x = np.clip(x, 0.0, None)
        x = x / x.sum(axis=1, keepdims=True)
        lr *= 0.995
    return x


@lru_cache
def equally_spaced_weights(dim: int, n: int, seed: int = 42) -> List[np.ndarray]:
    """Approximately equally spaced weights on the simplex (reference weights.py:38-49)."""
    try:
        from pymoo.util.ref_dirs import get_reference_directions

        return list(get_reference_directions("energy", dim, n, seed=seed))
    except Exception:
        return list(_riesz_energy_directions(dim, n, seed))


def extrema_weights(dim: int) -> List[np.ndarray]:
    """One-hot weight vectors (reference weights.py:52-58)."""
    return list(np.eye(dim, dtype=np.float32))
