"""Multi-GPU plumbing: one process per GPU (torch.distributed, NCCL over NVLink/NVSwitch).

The reference has no parallelism at all.  The hot path shards at the population / replica level: every rank
runs its own update stream with no data-path collective, and once per evaluation round the ranks exchange their local
non-dominated fronts with ONE all-gather of fixed-shape records, after which every rank runs the same global prune and
therefore holds the identical archive.

Record (float64): ``[ count | cap x d rows (-inf padded) | n_extra extras ]``.  ``extras`` carries whatever else the round has to
exchange (MORL/D: the evaluation of every policy the rank owns), so a round costs exactly one collective.  On CUDA the whole round is
stream-ordered -- local prune, pack, all-gather, unpack, global prune, pack (csrc/pareto.cu) -- with no host-visible count in between;
the result crosses to the host once, through a pinned buffer.  The count in a record is never clipped: if any rank's front exceeds
``cap`` every rank sees it in the gathered headers and the exchange is repeated with a larger capacity (same decision everywhere).

CPU tensors (the ``gloo`` tests of the protocol on a host without a GPU) take the same steps with torch ops and an injected dominance
test; that path is test plumbing, not a compute fallback.
"""

from __future__ import annotations

from typing import Callable, Optional

import torch as th
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int):
    """Contiguous block partition of ``n_items`` policies / weight vectors over ``world`` ranks (first ranks get the remainder)."""
    base, rem = divmod(n_items, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def pack_front(points: th.Tensor, cap: int, keep: Optional[th.Tensor] = None, extras: Optional[th.Tensor] = None) -> th.Tensor:
    """[n, d] -> float64 record [1 + cap*d + n_extra] (torch ops; CPU tensors / tests)."""
    pts = points.to(th.float64)
    if keep is not None:
        pts = pts[keep.bool()]
    n, d = pts.shape
    n_extra = 0 if extras is None else extras.numel()
    rec = th.full((1 + cap * d + n_extra,), float("-inf"), dtype=th.float64, device=points.device)
    rec[0] = n
    m = min(n, cap)
    rec[1 : 1 + m * d] = pts[:m].reshape(-1)
    if n_extra:
        rec[1 + cap * d :] = extras.to(th.float64).reshape(-1)
    return rec


def unpack_fronts(gathered: th.Tensor, world: int, cap: int, d: int, n_extra: int = 0):
    """[world, 1 + cap*d + n_extra] -> (all packed rows [world*cap, d] incl. -inf padding, counts [world], extras [world, n_extra])."""
    g = gathered.view(world, 1 + cap * d + n_extra)
    return g[:, 1 : 1 + cap * d].reshape(world * cap, d), g[:, 0].round().long(), g[:, 1 + cap * d :]


class _FrontBuffers:
    """Device + pinned buffers of one (world, cap, d, n_extra) exchange shape, reused across rounds."""

    def __init__(self, dev, world, cap, d, n_extra):
        rl = 1 + cap * d + n_extra
        self.rec = th.empty(rl, dtype=th.float64, device=dev)
        self.gathered = th.empty(world * rl, dtype=th.float64, device=dev)
        self.allpts = th.empty((world * cap, d), dtype=th.float64, device=dev)
        self.meta = th.empty((world, 1 + n_extra), dtype=th.float64, device=dev)
        self.keep = th.empty(world * cap, dtype=th.uint8, device=dev)
        self.final = th.empty(1 + world * cap * d + world * (1 + n_extra), dtype=th.float64, device=dev)
        self.final_pin = th.empty(self.final.numel(), dtype=th.float64).pin_memory()
        self.done = th.cuda.Event()


_buffers = {}


def _exchange_cuda(pts, cap, world, group, extras):
    from . import ops

    dev, d = pts.device, pts.shape[1]
    n_extra = 0 if extras is None else extras.numel()
    key = (dev.index, world, cap, d, n_extra)
    b = _buffers.get(key)
    if b is None:
        b = _buffers[key] = _FrontBuffers(dev, world, cap, d, n_extra)
    keep = ops.pareto_mask(pts, True, raw=True) if pts.shape[0] > 1 else None
    ops.front_pack(pts, keep, cap, b.rec, extras)
    if world > 1:
        dist.all_gather_into_tensor(b.gathered, b.rec, group=group)  # THE collective of the round
        ops.front_unpack(b.gathered, world, d, cap, n_extra, b.allpts, b.meta)
        ops.pareto_mask(b.allpts, True, raw=True, out=b.keep)
        ops.front_pack(b.allpts, b.keep, world * cap, b.final, b.meta.view(-1))
        b.final_pin.copy_(b.final, non_blocking=True)
    else:  # single rank: the local record is the result; same layout as the gathered case with world = 1
        b.final_pin[: 1 + cap * d].copy_(b.rec[: 1 + cap * d], non_blocking=True)
        b.final_pin[1 + cap * d : 2 + cap * d].copy_(b.rec[:1], non_blocking=True)
        if n_extra:
            b.final_pin[2 + cap * d :].copy_(b.rec[1 + cap * d :], non_blocking=True)
    b.done.record()
    b.done.synchronize()  # the only host wait of the round
    host = b.final_pin
    m = min(int(host[0]), world * cap)  # (a single rank's own overflow shows up here; the caller retries with a larger cap)
    front = host[1 : 1 + m * d].view(m, d).clone()
    meta = host[1 + world * cap * d :].view(world, 1 + n_extra)
    counts = meta[:, 0].round().long()
    if m > 0 and not bool(th.isfinite(front).all()):  # (only -inf padding survived: every real row was NaN)
        front = front[th.isfinite(front).all(dim=1)]
    return front, counts, meta[:, 1:].clone()


def _exchange_cpu(pts, cap, world, group, extras, prune):
    keep = prune(pts) if pts.shape[0] > 1 else None
    d = pts.shape[1]
    n_extra = 0 if extras is None else extras.numel()
    rec = pack_front(pts, cap, keep, extras)
    if world > 1:
        parts = [th.empty_like(rec) for _ in range(world)]
        dist.all_gather(parts, rec, group=group)
        gathered = th.stack(parts)
    else:
        gathered = rec.view(1, -1)
    allpts, counts, ex = unpack_fronts(gathered, world, cap, d, n_extra)
    valid = th.isfinite(allpts).all(dim=1)  # (torch path: drop the padding before the injected dominance test)
    allpts = allpts[valid]
    if allpts.shape[0] > 1:
        allpts = allpts[prune(allpts)]
    return allpts, counts, ex.clone()


def allgather_fronts(local_points: th.Tensor, cap: int = 256, prune: Optional[Callable[[th.Tensor], th.Tensor]] = None, group=None,
                     stats: Optional[dict] = None, extras: Optional[th.Tensor] = None):
    """Local prune -> ONE all-gather of fixed-shape front records -> global prune.  Returns the global non-dominated front (float64 [m, d]
    on the host, identical on every rank); with ``extras`` (a float64 vector, same length on every rank) returns ``(front, gathered
    extras [world, n_extra])``.  Works without an initialised process group (world size 1)."""
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    pts = local_points.to(th.float64).contiguous()
    if extras is not None:
        extras = extras.to(device=pts.device, dtype=th.float64).contiguous().view(-1)
    rounds = 0
    while True:
        rounds += 1
        if pts.is_cuda:
            front, counts, ex = _exchange_cuda(pts, cap, world, group, extras)
        else:
            if prune is None:
                raise RuntimeError("allgather_fronts on CPU tensors needs an explicit `prune` (the dominance kernel is CUDA-only)")
            front, counts, ex = _exchange_cpu(pts, cap, world, group, extras, prune)
        need = int(counts.max())
        if need <= cap:
            break
        cap = 1 << (need - 1).bit_length()  # every rank computes the same new capacity from the same headers
    if stats is not None:
        stats.update({"rounds": rounds, "cap": cap, "counts": counts.tolist()})
    return front if extras is None else (front, ex)


class DPFlat:
    """ONE collective per gradient update of a data-parallel Envelope learner ("DP-Envelope": the scalarising weight set
    of an update is sharded over the ranks, every rank back-propagates the loss rows of its own weights, and the conditioned network stays
    consistent through a gradient all-reduce).

    Flat float32 buffer ``[ gradients (every parameter, padded to 4 floats) | priorities (B) | loss (1) ]``: the parameters' ``.grad``
    tensors are VIEWS of the first segment (the backward kernels write into it directly), so the all-reduce needs no packing of the
    851 KB of gradients.  The priorities of an update come from the loss rows of weight 0 (reference envelope.py:329-331), which only the
    owner rank holds: it contributes them, the others contribute zeros, and the SUM hands them to everyone -- each rank then applies the
    same PER write-back and samples the same minibatch next step.  After ``allreduce()``: gradients and loss are the global means
    (mean over ranks of the local means: equal shard sizes), priorities the owner's."""

    def __init__(self, params, n_prio: int, group=None):
        params = list(params)
        dev = params[0].device
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        offs, o = [], 0
        for p in params:
            offs.append(o)
            o += (p.numel() + 3) // 4 * 4
        self.n_grad = o
        self.flat = th.zeros(o + n_prio + 1, dtype=th.float32, device=dev)
        self.grads = [self.flat[a : a + p.numel()].view_as(p) for a, p in zip(offs, params)]
        self.prio = self.flat[o : o + n_prio]
        self.loss = self.flat[o + n_prio : o + n_prio + 1]

    def allreduce(self, prio: Optional[th.Tensor], loss: th.Tensor, owns_priorities: bool):
        """prio [B] (raw |w . td| of the local weight 0 rows) and loss [1] of the local shard -> in place: global mean gradients in the
        ``.grad`` views, the owner's priorities in ``self.prio``, the global mean loss in ``self.loss``.  ONE all-reduce."""
        if prio is not None and owns_priorities:
            self.prio.copy_(prio.reshape(-1))
        else:
            self.prio.zero_()
        self.loss.copy_(loss.reshape(-1))
        if self.world > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
            inv = 1.0 / self.world
            self.flat[: self.n_grad].mul_(inv)
            self.loss.mul_(inv)
        return self.prio, self.loss
