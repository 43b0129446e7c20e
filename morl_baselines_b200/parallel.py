"""Multi-GPU plumbing: one process per GPU (torch.distributed, NCCL over NVLink/NVSwitch).

The reference has no parallelism at all.  The hot path shards at the population / replica level: every rank
runs its own update stream with no data-path collective, and once per evaluation round the ranks exchange their local
non-dominated fronts with ONE all-gather of fixed-capacity buffers, after which every rank runs the same global prune and
therefore holds the identical archive.  Variable-size fronts travel in a fixed ``1 + cap*d`` float64
record whose first slot is the true count; overflow is never silent: if any rank's count exceeds ``cap`` every rank sees
it in the gathered headers and the exchange is repeated with a larger capacity.
"""

from __future__ import annotations

from typing import Callable, Optional

import torch as th
import torch.distributed as dist


def _default_prune(points: th.Tensor) -> th.Tensor:
    from . import ops

    return ops.pareto_mask(points, True)


def shard_range(n_items: int, rank: int, world: int):
    """Contiguous block partition of ``n_items`` policies / weight vectors over ``world`` ranks (first ranks get the remainder)."""
    base, rem = divmod(n_items, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def pack_front(points: th.Tensor, cap: int) -> th.Tensor:
    """[n, d] -> float64 record [1 + cap*d]: count, then the first min(n, cap) rows."""
    n, d = points.shape
    rec = th.zeros(1 + cap * d, dtype=th.float64, device=points.device)
    rec[0] = n
    m = min(n, cap)
    rec[1 : 1 + m * d] = points[:m].to(th.float64).reshape(-1)
    return rec


def unpack_fronts(gathered: th.Tensor, world: int, cap: int, d: int):
    """[world, 1 + cap*d] -> (concatenated valid rows in rank order, per-rank counts)."""
    gathered = gathered.view(world, 1 + cap * d)
    counts = gathered[:, 0].round().long()
    rows = []
    for r in range(world):
        m = int(min(int(counts[r]), cap))
        rows.append(gathered[r, 1 : 1 + m * d].view(m, d))
    return th.cat(rows, dim=0), counts


def allgather_fronts(local_points: th.Tensor, cap: int = 256, prune: Optional[Callable[[th.Tensor], th.Tensor]] = None, group=None,
                     stats: Optional[dict] = None) -> th.Tensor:
    """Local prune -> one all-gather of fixed-capacity front records -> global prune.  Returns the global non-dominated
    front (float64 [m, d], identical on every rank).  Works without an initialised process group (world size 1)."""
    prune = prune or _default_prune
    pts = local_points.to(th.float64)
    if pts.shape[0] > 1:
        pts = pts[prune(pts)]
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    if world == 1:
        return pts
    d = pts.shape[1]
    rounds = 0
    while True:
        rounds += 1
        rec = pack_front(pts, cap)
        out = th.empty(world * rec.numel(), dtype=th.float64, device=rec.device)
        if dist.get_backend(group) == "nccl":
            dist.all_gather_into_tensor(out, rec, group=group)
        else:
            parts = [th.empty_like(rec) for _ in range(world)]
            dist.all_gather(parts, rec, group=group)
            out = th.cat(parts)
        allpts, counts = unpack_fronts(out, world, cap, d)
        need = int(counts.max())
        if need <= cap:
            break
        cap = 1 << (need - 1).bit_length()  # every rank computes the same new capacity from the same headers
    if stats is not None:
        stats.update({"rounds": rounds, "cap": cap, "counts": counts.tolist()})
    if allpts.shape[0] > 1:
        allpts = allpts[prune(allpts)]
    return allpts
