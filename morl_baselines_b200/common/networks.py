"""Neural-network helpers (mirrors reference morl_baselines/common/networks.py).

``mlp`` / ``layer_init`` build the same torch modules with the same parameter names, so reference ``state_dict``s load
unchanged (and orthogonal init consumes the torch RNG identically).  ``polyak_update`` runs as ONE multi-tensor CUDA
launch (morl_polyak_f32) instead of 1-2 tiny kernels per parameter tensor (reference networks.py:121-139).
"""

from __future__ import annotations

from typing import Iterable, List, Type

import numpy as np
import torch as th
from torch import nn

from .. import ops


def mlp(input_dim: int, output_dim: int, net_arch: List[int], activation_fn: Type[nn.Module] = nn.ReLU, drop_rate: float = 0.0,
        layer_norm: bool = False) -> nn.Sequential:
    """Fully connected stack: Linear [-> Dropout] [-> LayerNorm] -> activation per hidden layer, optional output Linear
    (layout of reference networks.py:10-48, so module indices / state_dict keys coincide)."""
    assert len(net_arch) > 0
    dims = [input_dim] + list(net_arch)
    modules: List[nn.Module] = []
    for i in range(len(net_arch)):
        modules.append(nn.Linear(dims[i], dims[i + 1]))
        if drop_rate > 0.0:
            modules.append(nn.Dropout(p=drop_rate))
        if layer_norm:
            modules.append(nn.LayerNorm(dims[i + 1]))
        modules.append(activation_fn())
    if output_dim > 0:
        modules.append(nn.Linear(dims[-1], output_dim))
    return nn.Sequential(*modules)


class NatureCNN(nn.Module):
    """DQN-Nature convolutional feature extractor (reference networks.py:51-87)."""

    def __init__(self, observation_shape: np.ndarray, features_dim: int = 512):
        super().__init__()
        self.features_dim = features_dim
        c_in = 1 if len(observation_shape) == 2 else observation_shape[0]
        self.cnn = nn.Sequential(
            nn.Conv2d(c_in, 32, kernel_size=8, stride=4, padding=0), nn.ReLU(),
            nn.Conv2d(32, 64, kernel_size=4, stride=2, padding=0), nn.ReLU(),
            nn.Conv2d(64, 64, kernel_size=3, stride=1, padding=0), nn.ReLU(),
            nn.Flatten(),
        )
        with th.no_grad():
            n_flatten = self.cnn(th.as_tensor(np.zeros(observation_shape)[np.newaxis]).float()).shape[1]
        self.linear = nn.Sequential(nn.Linear(n_flatten, features_dim), nn.ReLU())

    def forward(self, observations: th.Tensor) -> th.Tensor:
        if observations.dim() == 3:
            observations = observations.unsqueeze(0)
        return self.linear(self.cnn(observations / 255.0))


def huber(x, min_priority=0.01):
    """where(x < min_priority, x^2 / 2, min_priority * x).mean() (reference networks.py:90-100)."""
    return th.where(x < min_priority, 0.5 * x.pow(2), min_priority * x).mean()


def get_grad_norm(params: Iterable[th.nn.Parameter]) -> th.Tensor:
    """Global L2 norm of the gradients (reference networks.py:103-117)."""
    grads = [p.grad.detach() for p in params if p.grad is not None]
    if len(grads) == 0:
        return th.tensor(0.0)
    return th.norm(th.stack([th.norm(g, 2.0) for g in grads]), 2.0)


_POLYAK_PLANS = {}


@th.no_grad()
def polyak_update(params: Iterable[th.nn.Parameter], target_params: Iterable[th.nn.Parameter], tau: float) -> None:
    """target <- param if tau == 1 else fma(tau, param, (1 - tau) * target), all tensors in one CUDA launch."""
    params, targets = [p.data for p in params], [t.data for t in target_params]
    if len(params) == 0:
        return
    key = (tuple(p.data_ptr() for p in params), tuple(t.data_ptr() for t in targets))
    plan = _POLYAK_PLANS.get(key)
    if plan is None:
        if len(_POLYAK_PLANS) > 256:
            _POLYAK_PLANS.clear()
        plan = _POLYAK_PLANS[key] = ops.PolyakPlan(params, targets)
    plan.run(tau)


@th.no_grad()
def layer_init(layer, method="orthogonal", weight_gain: float = 1, bias_const: float = 0) -> None:
    """Orthogonal (default) or Xavier init of Linear / Conv2d layers, constant bias (reference networks.py:143-157)."""
    if isinstance(layer, (nn.Linear, nn.Conv2d)):
        if method == "xavier":
            th.nn.init.xavier_uniform_(layer.weight, gain=weight_gain)
        elif method == "orthogonal":
            th.nn.init.orthogonal_(layer.weight, gain=weight_gain)
        th.nn.init.constant_(layer.bias, bias_const)
