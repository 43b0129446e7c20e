"""Prioritised replay (mirrors reference morl_baselines/common/prioritized_buffer.py).

The sum-tree keeps the reference's layout -- one float64 array per level, root first (prioritized_buffer.py:19-28) -- and
its exact semantics: proportional sampling by a batched level walk with the GLOBAL numpy RNG (:40-54, strict '>' goes
right), duplicate-safe ``batch_set`` where the first occurrence of an index wins (:76-82), new transitions entering with
``min_priority`` which ratchets up to the largest priority ever written (:194).  Transition storage and the minibatch
gather are inherited from the device-mirrored ReplayBuffer.
"""

from __future__ import annotations

import numpy as np
import torch as th

from .buffer import ReplayBuffer, ReplayBufferSamplesNp


class SumTree:
    """Fixed-size sum tree over float64 level arrays (reference prioritized_buffer.py:12-82)."""

    def __init__(self, max_size):
        self.nodes = []
        level_size = 1
        for _ in range(int(np.ceil(np.log2(max_size))) + 1):
            self.nodes.append(np.zeros(level_size))
            level_size *= 2

    def sample(self, batch_size):
        query = np.random.uniform(0, self.nodes[0][0], size=batch_size)
        return self.walk(query)

    def walk(self, query):
        """Descend the tree for given query values (the deterministic part of ``sample``)."""
        query = np.array(query, dtype=np.float64)
        node = np.zeros(len(query), dtype=int)
        for nodes in self.nodes[1:]:
            node *= 2
            left = nodes[node]
            greater = np.greater(query, left)
            node += greater
            query -= left * greater
        return node

    def set(self, node_index, new_priority):
        diff = new_priority - self.nodes[-1][node_index]
        for nodes in self.nodes[::-1]:
            np.add.at(nodes, node_index, diff)
            node_index //= 2

    def batch_set(self, node_index, new_priority):
        node_index, unique_index = np.unique(node_index, return_index=True)
        diff = new_priority[unique_index] - self.nodes[-1][node_index]
        for nodes in self.nodes[::-1]:
            np.add.at(nodes, node_index, diff)
            node_index //= 2


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
