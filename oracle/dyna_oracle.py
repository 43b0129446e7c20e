"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

numpy restatement of the sampling half of the reference's probabilistic ensemble (GPI-PD's Dyna path):
  * ``clamp_logvar``  : common/model_based/probabilistic_ensemble.py:118-119 (two soft clamps; softplus as torch's F.softplus, threshold 20)
  * ``ensemble_sample``: :127-128 (reparameterised sample) and :136-154 (``sample``: variances, the elite drawn per row, ensemble moments,
    uncertainty), plus ``samples[:, rew_dim:] += obs`` of ModelEnv.step (common/model_based/utils.py:165).
Pinned against outputs of the unmodified reference (tests/golden/dyna.npz, tests/test_dyna_cpu.py): bit-exact given the reference's own
mean / logvar tensors.  Checker of ``morl_ensemble_sample_f32`` (csrc/dyna.cu) in tests/test_dyna_gpu.py."""

import numpy as np


def softplus(x):
    x = np.asarray(x, np.float32)
    with np.errstate(over="ignore"):
        return np.where(x > 20, x, np.log1p(np.exp(x, dtype=np.float32), dtype=np.float32)).astype(np.float32)


def clamp_logvar(logvar, max_logvar, min_logvar):
    logvar = max_logvar - softplus(max_logvar - logvar)
    return (min_logvar + softplus(logvar - min_logvar)).astype(np.float32)


def ensemble_sample(means, logvar, model_inds, noise=None, obs=None, rew_dim=0):
    """means, logvar [E, N, O] (logvar already clamped); model_inds [N]; noise [E, N, O] or None (deterministic).
    Returns (sample [N, O], var [N, O], uncertainty [N]) in float32, operation for operation as the reference's numpy code."""
    means = np.asarray(means, np.float32)
    logvar = np.asarray(logvar, np.float32)
    if noise is not None:
        std = np.exp(np.float32(0.5) * logvar)
        samples = means + std * np.asarray(noise, np.float32)
    vars_ = np.exp(logvar)
    batch_inds = np.arange(0, means.shape[1])
    mean_ensemble = means.mean(axis=0)
    var_ensemble = (means**2 + vars_).mean(axis=0) - mean_ensemble**2
    std_ensemble = np.sqrt(var_ensemble + 1e-12)
    uncertainties = std_ensemble.sum(-1)
    picked = (means if noise is None else samples)[model_inds, batch_inds].copy()
    if obs is not None:
        picked[:, rew_dim:] += np.asarray(obs, np.float32)
    return picked, vars_[model_inds, batch_inds], uncertainties
