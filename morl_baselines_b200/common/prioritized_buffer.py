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


class DeviceSumTree:
    """The same tree resident in HBM (csrc/sumtree.cu): same layout, same float64 operations in the same order as the
    reference's numpy tree, so walks and updates are bit-identical to ``SumTree`` -- but stream-ordered, so that sampling, the minibatch
    gather, the update and the priority write-back can be captured in ONE CUDA graph with no device->host->device round trip.

    Host-facing methods (``sample``, ``walk``, ``set``, ``batch_set``, ``nodes``) keep the reference's signatures and synchronise;
    ``walk_into`` / ``batch_set_dev`` are the in-graph forms on device tensors."""

    MAX_BATCH = 2048

    def __init__(self, max_size, device):
        self.n_levels = int(np.ceil(np.log2(max_size))) + 1
        self.device = th.device(device)
        self._lib = _lib.load()
        self.flat = th.zeros((1 << self.n_levels) - 1, dtype=th.float64, device=self.device)
        self.err = th.zeros(1, dtype=th.int32, device=self.device)

    # ---- host-facing (reference signatures)
    @property
    def nodes(self):
        """The reference's list of per-level arrays (root first), as numpy views of a host copy (synchronises)."""
        flat = self.flat.cpu().numpy()
        return [flat[(1 << l) - 1 : (1 << (l + 1)) - 1] for l in range(self.n_levels)]

    def __getstate__(self):
        return {"n_levels": self.n_levels, "flat": self.flat.cpu().numpy(), "device": str(self.device)}

    def __setstate__(self, state):
        self.n_levels, self.device = state["n_levels"], th.device(state["device"])
        self._lib = _lib.load()
        self.flat = th.from_numpy(np.ascontiguousarray(state["flat"], dtype=np.float64)).to(self.device)
        self.err = th.zeros(1, dtype=th.int32, device=self.device)

    def _stream(self):
        return th.cuda.current_stream().cuda_stream

    def sample(self, batch_size):
        u = np.random.random_sample(batch_size)  # the stream np.random.uniform(0, root, n) consumes (:40); query = root * u on the device
        return self._walk_host(u, scale_by_root=True)

    def walk(self, query):
        return self._walk_host(np.asarray(query, dtype=np.float64), scale_by_root=False)

    def _walk_host(self, u, scale_by_root):
        u_dev = th.from_numpy(np.ascontiguousarray(u, dtype=np.float64)).to(self.device)
        out = th.empty(u_dev.shape[0], dtype=th.int64, device=self.device)
        self.walk_into(u_dev, out, scale_by_root)
        return out.cpu().numpy()

    def set(self, node_index, new_priority=None, min_priority_dev=None):
        """One leaf (``replay_buffer.add``); ``new_priority=None`` takes the buffer's current (device-resident) min_priority."""
        use_min = new_priority is None
        _lib.check(self._lib.morl_sumtree_set_f64(self.flat.data_ptr(), self.n_levels, int(node_index), 0.0 if use_min else float(new_priority),
                                                  int(use_min), None if min_priority_dev is None else min_priority_dev.data_ptr(),
                                                  self.err.data_ptr(), self._stream()), "morl_sumtree_set_f64")

    def batch_set(self, node_index, new_priority):
        """np.unique first-occurrence semantics + per-level np.add.at order (:66-82).  Batches beyond MAX_BATCH are made unique and sorted
        on the host and applied in consecutive chunks -- the same additions in the same order."""
        idx = np.ascontiguousarray(node_index, dtype=np.int64).reshape(-1)
        pr = np.ascontiguousarray(new_priority, dtype=np.float64).reshape(-1)
        if pr.shape[0] != idx.shape[0]:
            raise ValueError("batch_set: node_index and new_priority must have the same length")
        if idx.shape[0] > self.MAX_BATCH:
            idx, first = np.unique(idx, return_index=True)
            pr = pr[first]
        for i in range(0, idx.shape[0], self.MAX_BATCH):
            self.batch_set_dev(th.from_numpy(idx[i:i + self.MAX_BATCH]).to(self.device), th.from_numpy(pr[i:i + self.MAX_BATCH]).to(self.device))
        self.check()

    def check(self):
        if int(self.err.item()) != 0:
            self.err.zero_()
            raise _lib.MorlB200Error("DeviceSumTree: leaf index out of range")

    # ---- in-graph forms (device tensors, no synchronisation)
    def walk_into(self, u_dev, out_idx_dev, scale_by_root=True):
        _lib.check(self._lib.morl_sumtree_walk_f64(self.flat.data_ptr(), self.n_levels, u_dev.data_ptr(), u_dev.shape[0], int(scale_by_root),
                                                   out_idx_dev.data_ptr(), self._stream()), "morl_sumtree_walk_f64")

    def batch_set_dev(self, idx_dev, prio64_dev):
        _lib.check(self._lib.morl_sumtree_batch_set_f64(self.flat.data_ptr(), self.n_levels, idx_dev.data_ptr(), prio64_dev.data_ptr(),
                                                        idx_dev.shape[0], self.err.data_ptr(), self._stream()), "morl_sumtree_batch_set_f64")


class PrioritizedReplayBuffer(ReplayBuffer):
    """Prioritised replay buffer (same constructor and methods as reference prioritized_buffer.py:85-226).  With ``tree_on_device`` (needs
    ``device``) the sum tree and ``min_priority`` live in HBM next to the transition mirror (DeviceSumTree)."""

    def __init__(self, obs_shape, action_dim, rew_dim=1, max_size=100000, obs_dtype=np.float32, action_dtype=np.float32,
                 min_priority=1e-5, device=None, tree_on_device=False):
        super().__init__(obs_shape, action_dim, rew_dim=rew_dim, max_size=max_size, obs_dtype=obs_dtype, action_dtype=action_dtype,
                         device=device)
        self.tree_on_device = bool(tree_on_device and device is not None)
        if self.tree_on_device:
            self.tree = DeviceSumTree(max_size, device)
            self._min_p_dev = th.full((1,), float(min_priority), dtype=th.float64, device=device)  # (a python float in the reference until the first ratchet)
        else:
            self.tree = SumTree(max_size)
            self._min_p_host = min_priority

    # -- pickling / checkpoints: a device tree is saved as the (bit-identical) host tree and re-created by ``to(device)``
    def __getstate__(self):
        d = dict(super().__getstate__())
        if self.tree_on_device:
            host = SumTree.__new__(SumTree)
            host.n_levels = self.tree.n_levels
            host._alloc(self.tree.flat.cpu().numpy())
            d.pop("_min_p_dev", None)
            d.update(tree=host, _min_p_host=self.min_priority, tree_on_device=False, _want_device_tree=True)
        return d

    def to(self, device):
        super().to(device)
        if getattr(self, "_want_device_tree", False) and not self.tree_on_device and th.device(device).type == "cuda":
            t = DeviceSumTree(self.max_size, device)
            t.flat.copy_(th.from_numpy(self.tree._flat))
            self._min_p_dev = th.full((1,), float(self._min_p_host), dtype=th.float64, device=device)
            self.tree, self.tree_on_device = t, True
        return self

    @property
    def min_priority(self):
        """Largest priority ever written (starts at the constructor's value).  Device trees keep it on the device: reading synchronises."""
        return float(self._min_p_dev.item()) if self.tree_on_device else self._min_p_host

    @min_priority.setter
    def min_priority(self, value):
        if self.tree_on_device:
            self._min_p_dev.fill_(float(value))
        else:
            self._min_p_host = value

    def add(self, obs, action, reward, next_obs, done, priority=None):
        p = self.ptr
        super().add(obs, action, reward, next_obs, done)
        if self.tree_on_device:
            self.tree.set(p, priority, self._min_p_dev)  # one tiny launch, no synchronisation (min_priority is read on the device)
        else:
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

    def update_priorities_dev(self, idx_dev, raw_dev, alpha: float, prio64_dev, prio32_dev=None):
        """In-graph form of  p = (raw + min_priority) ** alpha; update_priorities(idx, p)  (envelope.py:333-334 + :186-195 here) on device
        tensors: two launches, no synchronisation.  ``raw_dev`` float32 [n] (e.g. |w . td|), ``prio64_dev`` float64 [n] scratch."""
        if not self.tree_on_device:
            raise _lib.MorlB200Error("update_priorities_dev needs tree_on_device=True")
        lib = _lib.load()
        st = th.cuda.current_stream().cuda_stream
        _lib.check(lib.morl_per_priority_f32(raw_dev.data_ptr(), raw_dev.shape[0], float(alpha), self._min_p_dev.data_ptr(), prio64_dev.data_ptr(),
                                             None if prio32_dev is None else prio32_dev.data_ptr(), st), "morl_per_priority_f32")
        self.tree.batch_set_dev(idx_dev, prio64_dev)

    def get_all_data(self, max_samples=None, to_tensor=False, device=None):
        if max_samples is not None and max_samples < self.size:
            inds = np.random.choice(self.size, max_samples, replace=False)
        else:
            inds = np.arange(self.size)
        tup = (self.obs[inds], self.actions[inds], self.rewards[inds], self.next_obs[inds], self.dones[inds])
        if to_tensor:
            return tuple(map(lambda x: th.tensor(x).to(device), tup))
        return tup
