"""CUDA-graph capture of a device-only update step (B200-native replacement of the reference's eager op-by-op updates).

The small actor-critic updates of the reference (MOSAC: mosac_continuous_action.py:429-507; CAPQL: capql.py:321-362;
GPI-PD continuous: gpi_pd_continuous_action.py:373-452) are ~250 tiny tensor operations -- launch-bound on any GPU (7.7 ms
eager on a B200 for a 128-row minibatch, 13 ms on the reference's CPU path).  ``GraphedStep`` captures one whole update over
STATIC input buffers (replay indices, optional injected noise) once and replays it: one host call per update.

Warm-up iterations and the capture pass itself must leave no trace (the number of updates applied to the parameters has to match
the reference exactly), so the caller lists every tensor the step mutates and they are restored IN PLACE after capture.
"""

from __future__ import annotations

from typing import Callable, Iterable, List

import torch as th


class GraphedStep:
    def __init__(self, fn: Callable[[], None], mutated: Callable[[], Iterable[th.Tensor]], warmup: int = 3):
        self.fn = fn
        self.mutated = mutated
        self.warmup = warmup
        self.graph = None

    def capture(self):
        tensors: List[th.Tensor] = list(self.mutated())
        snap = [t.detach().clone() for t in tensors]
        rng = th.cuda.get_rng_state()
        side = th.cuda.Stream()
        side.wait_stream(th.cuda.current_stream())
        with th.cuda.stream(side):
            for _ in range(self.warmup):
                self.fn()
        th.cuda.current_stream().wait_stream(side)
        g = th.cuda.CUDAGraph()
        with th.cuda.graph(g):
            self.fn()
        with th.no_grad():
            for t, s in zip(tensors, snap):
                t.copy_(s)
        th.cuda.set_rng_state(rng)
        self.graph = g

    def __call__(self):
        if self.graph is None:
            self.capture()
        self.graph.replay()


def optimizer_tensors(opt) -> List[th.Tensor]:
    """Every state tensor of a FusedClipAdam (created if the optimiser has not stepped yet, so that a snapshot exists)."""
    opt._ensure_state()
    out = []
    for st in opt.state.values():
        out += [st["step"], st["exp_avg"], st["exp_avg_sq"]]
    return out
