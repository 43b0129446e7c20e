"""World-size-2 gloo test (CPU) of the DP-Envelope exchange record (parallel.DPFlat; SURVEY 8(e): weight-set sharding with ONE gradient
all-reduce per update): after the collective every rank holds the mean of the ranks' gradients in its parameters' .grad views, the OWNER
rank's priorities, and the mean loss -- and exactly one collective was issued."""

import os
import socket

import numpy as np
import torch as th
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from morl_baselines_b200.parallel import DPFlat

        th.manual_seed(0)
        net = th.nn.Sequential(th.nn.Linear(7, 5), th.nn.ReLU(), th.nn.Linear(5, 3))  # parameter sizes 35, 5, 15, 3: padding between the segments
        params = list(net.parameters())
        flat = DPFlat(params, n_prio=6)
        calls = {"n": 0}
        orig = dist.all_reduce

        def counting(*a, **k):
            calls["n"] += 1
            return orig(*a, **k)

        dist.all_reduce = counting
        for p, g in zip(params, flat.grads):
            p.grad = g
            g.copy_(th.full_like(g, float(rank + 1)) * th.arange(g.numel(), dtype=th.float32).view_as(g))
        prio = th.arange(6, dtype=th.float32) + 10.0 * (rank + 1)
        got_prio, got_loss = flat.allreduce(prio, th.tensor([float(rank + 1)]), owns_priorities=(rank == 0))
        dist.all_reduce = orig
        ok = calls["n"] == 1
        for p in params:
            want = 1.5 * th.arange(p.numel(), dtype=th.float32).view_as(p)  # mean of 1x and 2x
            ok = ok and bool(th.equal(p.grad, want))
        ok = ok and bool(th.equal(got_prio, th.arange(6, dtype=th.float32) + 10.0)) and float(got_loss) == 1.5
        out.put((rank, ok))
    finally:
        dist.destroy_process_group()


def test_dp_flat_gloo_world2():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(out.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(60)
    assert res == [(0, True), (1, True)]
