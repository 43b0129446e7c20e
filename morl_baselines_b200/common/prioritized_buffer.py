"""Prioritised replay (mirrors reference morl_baselines/common/prioritized_buffer.py).

The sum-tree keeps the reference's layout -- one float64 array per level, root first (prioritized_buffer.py:19-28; here views
of one flat array so that the walk and the update run in C, csrc/host_replay.cu) -- and its exact semantics: proportional sampling by a batched level walk with the GLOBAL numpy RNG (:40-54, strict '>' goes
right), duplicate-safe ``batch_set`` where the first occurrence of an index wins (:76-82), new transitions entering with
``min_priority`` which ratchets up to the largest priority ever written (:194).  Transition storage and the minibatch
gather are inherited from the device-mirrored ReplayBuffer.
"""

from __future__ import annotations

import numpy as np
import torch as th

from .. import _lib
from .buffer import ReplayBuffer, ReplayBufferSamplesNp


class SumTree:
    """Fixed-size sum tree over float64 level arrays (reference prioritized_buffer.py:12-82).

    ``nodes`` is the reference's list of per-level arrays (root first); here they are views of ONE flat float64 array so that
    the walk and the batched update run in C (csrc/host_replay.cu: morl_host_sumtree_walk / _batch_set) with the same float64
    operations in the same order -- the tree sits on the critical path between two GPU steps."""

    def __init__(self, max_size):
        self.n_levels = int(np.ceil(np.log2(max_size))) + 1
        self._alloc(np.zeros((1 << self.n_levels) - 1))

    def _alloc(self, flat):
        self._flat = np.ascontiguousarray(flat, dtype=np.float64)
        self.nodes = [self._flat[(1 << l) - 1 : (1 << (l + 1)) - 1] for l in range(self.n_levels)]
        self._lib = _lib.load()

    def __getstate__(self):
        return {"n_levels": self.n_levels, "flat": self._flat}

    def __setstate__(self, state):
        self.n_levels = state["n_levels"]
        self._alloc(state["flat"])

    def sample(self, batch_size):
        query = np.random.uniform(0, self.nodes[0][0], size=batch_size)  # global numpy RNG, as the reference (:40)
        return self.walk(query)

    def walk(self, query):
        """Descend the tree for given query values (the deterministic part of ``sample``; strict '>' goes right)."""
        query = np.ascontiguousarray(query, dtype=np.float64)
        out = np.empty(query.shape[0], dtype=np.int64)
        _lib.check(self._lib.morl_host_sumtree_walk(self._flat.ctypes.data, self.n_levels, query.ctypes.data, query.shape[0], out.ctypes.data),
                   "morl_host_sumtree_walk")
        return out

    def set(self, node_index, new_priority):
        self.batch_set(np.array([node_index], dtype=np.int64), np.array([new_priority], dtype=np.float64))

    def batch_set(self, node_index, new_priority):
        """np.unique(node_index, return_index=True) + np.add.at on every level (:73-82), in C."""
        idx = np.ascontiguousarray(node_index, dtype=np.int64).reshape(-1)
        pr = np.ascontiguousarray(new_priority, dtype=np.float64).reshape(-1)
        if pr.shape[0] != idx.shape[0]:
            raise ValueError("batch_set: node_index and new_priority must have the same length")
        _lib.check(self._lib.morl_host_sumtree_batch_set(self._flat.ctypes.data, self.n_levels, idx.ctypes.data, pr.ctypes.data, idx.shape[0]),
                   "morl_host_sumtree_batch_set")


class PrioritizedReplayBuffer(ReplayBuffer):
    """Prioritised replay buffer (same constructor and methods as reference prioritized_buffer.py:85-226)."""

    def __init__(self, obs_shape, action_dim, rew_dim=1, max_size=100000, obs_dtype=np.float32, action_dtype=np.float32,
                 min_priority=1e-5, device=None):
        super().__init__(obs_shape, action_dim, rew_dim=rew_dim, max_size=max_size, obs_dtype=obs_dtype, action_dtype=action_dtype,
                         device=device)
        self.tree = SumTree(max_size)
        self.min_priority = min_priority

    def add(self, obs, action, reward, next_obs, done, priority=None):
        p = self.ptr
        super().add(obs, action, reward, next_obs, done)
        self.tree.set(p, self.min_priority if priority is None else priority)

    def sample(self, batch_size, to_tensor=False, device=None):
        idxes = self.tree.sample(batch_size)
        if to_tensor and self._dev is not None and (device is None or th.device(device).type == "cuda"):
            obs, act, rew, nobs, done = self.gather_device(idxes)
            return obs, act, rew, nobs, done, th.from_numpy(idxes)
        tup = ReplayBufferSamplesNp(self.obs[idxes], self.actions[idxes], self.rewards[idxes], self.next_obs[idxes], self.dones[idxes], idxes)
        if to_tensor:
            return tuple(map(lambda x: th.tensor(x).to(device), tup))
        return tup

    def sample_obs(self, batch_size, to_tensor=False, device=None):
        idxes = self.tree.sample(batch_size)
        return th.tensor(self.obs[idxes]).to(device) if to_tensor else self.obs[idxes]

    def update_priorities(self, idxes, priorities):
        self.min_priority = max(self.min_priority, priorities.max())
        self.tree.batch_set(np.asarray(idxes), priorities)

    def get_all_data(self, max_samples=None, to_tensor=False, device=None):
        if max_samples is not None and max_samples < self.size:
            inds = np.random.choice(self.size, max_samples, replace=False)
        else:
            inds = np.arange(self.size)
        tup = (self.obs[inds], self.actions[inds], self.rewards[inds], self.next_obs[inds], self.dones[inds])
        if to_tensor:
            return tuple(map(lambda x: th.tensor(x).to(device), tup))
        return tup
