"""TEST INFRASTRUCTURE ONLY -- exact hypervolume (minimisation form) by recursive dimension sweep; stands in for
pymoo.indicators.hv.HV (not installed) when the reference's performance_indicators.hypervolume is exercised
(reference common/performance_indicators.py:15-25).  pymoo's HV is an exact algorithm, so any exact HV agrees with it to
fp64 rounding ("parity unpinned")."""

import numpy as np


def hypervolume_min(points: np.ndarray, ref: np.ndarray) -> float:
    pts = np.asarray(points, dtype=np.float64).reshape(-1, len(ref))
    pts = pts[np.all(pts < ref, axis=1)]
    return _hv(pts, np.asarray(ref, dtype=np.float64))


def _hv(pts, ref):
    if len(pts) == 0:
        return 0.0
    d = pts.shape[1]
    if d == 1:
        return float(ref[0] - pts[:, 0].min())
    order = np.argsort(pts[:, -1])
    pts = pts[order]
    total = 0.0
    for i in range(len(pts)):
        upper = pts[i + 1, -1] if i + 1 < len(pts) else ref[-1]
        depth = upper - pts[i, -1]
        if depth > 0:
            total += depth * _hv(pts[: i + 1, :-1], ref[:-1])
    return total
