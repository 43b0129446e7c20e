"""Replay buffer with a device-resident mirror (mirrors reference morl_baselines/common/buffer.py).

The numpy attributes ``obs, next_obs, actions, rewards, dones, ptr, size`` are the reference's (buffer.py:42-48) so that
pickled buffers / checkpoints interchange.  They are views of PINNED host tensors; when a CUDA ``device`` is given the
buffer keeps a mirror of every array in HBM, flushes newly added rows with asynchronous copies, and ``sample`` gathers the
minibatch with ONE kernel (morl_replay_gather) instead of five host gathers + six synchronous host->device copies
(buffer.py:82-94).  Index sampling stays on the host with the reference's global-RNG call (np.random.choice, :82) so
seeded runs draw identical minibatches.
"""

from __future__ import annotations

from typing import NamedTuple, Optional

import numpy as np
import torch as th

from .. import ops


class ReplayBufferSamplesNp(NamedTuple):
    observations: np.ndarray
    actions: np.ndarray
    rewards: np.ndarray
    next_observations: np.ndarray
    dones: np.ndarray
    idxes: np.ndarray


_TORCH_DTYPES = {np.dtype(np.float32): th.float32, np.dtype(np.uint8): th.uint8, np.dtype(np.float64): th.float64,
                 np.dtype(np.int64): th.int64, np.dtype(np.int32): th.int32}


def _host_array(shape, dtype, pin: bool):
    t = th.zeros(shape, dtype=_TORCH_DTYPES[np.dtype(dtype)])
    if pin:
        try:
            t = t.pin_memory()
        except Exception:
            pass
    return t


class ReplayBuffer:
    """Multi-objective replay buffer (same constructor and methods as reference buffer.py:20-139)."""

    def __init__(self, obs_shape, action_dim, rew_dim=1, max_size=100000, obs_dtype=np.float32, action_dtype=np.float32,
                 device: Optional[th.device] = None):
        self.max_size = max_size
        self.ptr, self.size = 0, 0
        self.device = th.device(device) if device is not None else None
        on_gpu = self.device is not None and self.device.type == "cuda"
        self._h_obs = _host_array((max_size,) + tuple(obs_shape), obs_dtype, on_gpu)
        self._h_next_obs = _host_array((max_size,) + tuple(obs_shape), obs_dtype, on_gpu)
        self._h_actions = _host_array((max_size, action_dim), action_dtype, on_gpu)
        self._h_rewards = _host_array((max_size, rew_dim), np.float32, on_gpu)
        self._h_dones = _host_array((max_size, 1), np.float32, on_gpu)
        self.obs, self.next_obs = self._h_obs.numpy(), self._h_next_obs.numpy()
        self.actions, self.rewards, self.dones = self._h_actions.numpy(), self._h_rewards.numpy(), self._h_dones.numpy()
        self._dev = None
        self._dirty = []  # list of (start, stop) row ranges not yet mirrored
        if on_gpu:
            self._dev = tuple(th.zeros_like(h, device=self.device) for h in self._host_tensors())

    def _host_tensors(self):
        return (self._h_obs, self._h_next_obs, self._h_actions, self._h_rewards, self._h_dones)

    # -- pickling: keep the reference's attribute layout, drop device state ------------------------------
    def __getstate__(self):
        d = {k: v for k, v in self.__dict__.items() if not k.startswith("_h_") and k not in ("_dev", "_dirty", "device")}
        return d

    def __setstate__(self, state):
        self.__dict__.update(state)
        self.device, self._dev, self._dirty = None, None, []
        for name in ("obs", "next_obs", "actions", "rewards", "dones"):
            setattr(self, "_h_" + name, th.from_numpy(getattr(self, name)))

    def to(self, device):
        """(Re)attach a device mirror, e.g. after unpickling a checkpointed buffer."""
        self.device = th.device(device)
        if self.device.type == "cuda":
            self._dev = tuple(h.to(self.device) for h in self._host_tensors())
            self._dirty = []
        return self

    def add(self, obs, action, reward, next_obs, done):
        """Append one transition (reference buffer.py:50-66)."""
        p = self.ptr
        self.obs[p] = np.array(obs).copy()
        self.next_obs[p] = np.array(next_obs).copy()
        self.actions[p] = np.array(action).copy()
        self.rewards[p] = np.array(reward).copy()
        self.dones[p] = np.array(done).copy()
        self._mark_dirty(p, p + 1)
        self.ptr = (self.ptr + 1) % self.max_size
        self.size = min(self.size + 1, self.max_size)

    def add_batch(self, obs, actions, rewards, next_obs, dones):
        """Append n transitions at once (arrays or tensors with a leading batch axis): the state afterwards equals n ``add`` calls in
        row order (ring wrap-around included).  Device tensors are written straight into the HBM mirror as well, so nothing has to be
        re-sent.  Replaces the row-by-row python loop of the reference's Dyna rollout (gpi_pd.py:394-397)."""
        n = int(obs.shape[0])
        if n == 0:
            return
        cols = (obs, next_obs, actions, rewards, dones)
        on_dev = self._dev is not None and all(isinstance(c, th.Tensor) and c.is_cuda for c in cols)
        if n > self.max_size:  # only the last max_size rows survive n sequential adds
            skip = n - self.max_size
            cols = tuple(c[skip:] for c in cols)
            self.ptr = (self.ptr + skip) % self.max_size
            n = self.max_size
        first = min(n, self.max_size - self.ptr)
        spans = [(self.ptr, 0, first)] + ([(0, first, n - first)] if n > first else [])
        for h, d, c in zip(self._host_tensors(), self._dev if on_dev else (None,) * 5, cols):
            c = c if isinstance(c, th.Tensor) else th.as_tensor(np.asarray(c))
            c = c.reshape((n,) + tuple(h.shape[1:])).to(h.dtype)
            for dst, src, cnt in spans:
                if on_dev:
                    d[dst:dst + cnt].copy_(c[src:src + cnt])
                h[dst:dst + cnt].copy_(c[src:src + cnt], non_blocking=False)
        if not on_dev:
            for dst, _, cnt in spans:
                self._mark_dirty(dst, dst + cnt)
        self.ptr = (self.ptr + n) % self.max_size
        self.size = min(self.size + n, self.max_size)

    def _mark_dirty(self, a, b):
        if self._dev is None:
            return
        if self._dirty and self._dirty[-1][1] == a:
            self._dirty[-1] = (self._dirty[-1][0], b)
        else:
            self._dirty.append((a, b))

    def mark_all_dirty(self):
        """Call after writing the numpy attributes directly (bulk fills in benchmarks / tests)."""
        if self._dev is not None:
            self._dirty = [(0, self.max_size)]

    def flush(self):
        """Mirror newly written rows into HBM (asynchronous copies from pinned memory on the current stream)."""
        if self._dev is None or not self._dirty:
            return
        for a, b in self._dirty:
            for h, d in zip(self._host_tensors(), self._dev):
                d[a:b].copy_(h[a:b], non_blocking=True)
        self._dirty = []

    def device_stores(self):
        self.flush()
        return self._dev

    def _draw(self, batch_size, replace=True, use_cer=False):
        inds = np.random.choice(self.size, batch_size, replace=replace)  # global numpy RNG, as the reference (:82)
        if use_cer:
            inds[0] = self.ptr - 1
        return inds

    def gather_device(self, inds: np.ndarray, idx_staging: Optional[th.Tensor] = None):
        """Minibatch for host-drawn indices, gathered on the GPU: (obs, actions int32|f32, rewards, next_obs, dones)."""
        obs_s, nobs_s, act_s, rew_s, done_s = self.device_stores()
        if idx_staging is None:
            idx = th.from_numpy(np.ascontiguousarray(inds, dtype=np.int64)).to(self.device, non_blocking=True)
        else:
            idx_staging.copy_(th.from_numpy(np.ascontiguousarray(inds, dtype=np.int64)), non_blocking=True)
            idx = idx_staging
        return ops.replay_gather(obs_s, nobs_s, act_s, rew_s, done_s, idx)

    def sample(self, batch_size, replace=True, use_cer=False, to_tensor=False, device=None):
        """Sample a minibatch (reference buffer.py:68-96).  With ``to_tensor`` and a device mirror the gather runs on the
        GPU; uint8 actions are returned as int32 there (the reference's callers immediately call ``.long()``)."""
        inds = self._draw(batch_size, replace, use_cer)
        if to_tensor and self._dev is not None and (device is None or th.device(device).type == "cuda"):
            obs, act, rew, nobs, done = self.gather_device(inds)
            return obs, act, rew, nobs, done, th.from_numpy(inds)
        tup = ReplayBufferSamplesNp(self.obs[inds], self.actions[inds], self.rewards[inds], self.next_obs[inds], self.dones[inds], inds)
        if to_tensor:
            return tuple(map(lambda x: th.tensor(x, device=device), tup))
        return tup

    def sample_obs(self, batch_size, replace=True, to_tensor=False, device=None):
        """Sample observations only (reference buffer.py:98-114)."""
        inds = np.random.choice(self.size, batch_size, replace=replace)
        return th.tensor(self.obs[inds], device=device) if to_tensor else self.obs[inds]

    def get_all_data(self, max_samples=None):
        """All stored transitions, optionally a random subset (reference buffer.py:116-135)."""
        if max_samples is not None:
            inds = np.random.choice(self.size, min(max_samples, self.size), replace=False)
        else:
            inds = np.arange(self.size)
        return (self.obs[inds], self.actions[inds], self.rewards[inds], self.next_obs[inds], self.dones[inds])

    def __len__(self):
        return self.size
