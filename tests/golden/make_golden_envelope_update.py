"""Whole-update golden vectors of the UNMODIFIED reference ``Envelope.update()`` (multi_policy/envelope/envelope.py:266-367) at the
BASELINE.json shapes, produced on CPU in the build container:

    python tests/golden/make_golden_envelope_update.py     ->  tests/golden/envelope_update.npz

Cases (all ``per=True``, default lr / gamma / max_grad_norm, target net = online net + 0.01 N(0,1) so the double-Q distinction of
envelope.py:420 vs :429 is live from the first step):
  * north_star : obs 32, |A| 8, d 3, |W| 64, B 1024, net 4x256                    (BASELINE.json metric shape), 2 updates
  * config2    : minecart dims obs 7, |A| 6, d 3, |W| 32, B 256, net 4x256        (BASELINE.json configs[1]),   3 updates
  * homotopy   : config2 dims, net 2x256, homotopy schedule 0.2 -> 1.0 over 10 steps (lambda differs every update), 3 updates
Frozen per case: the initial parameters, per update the critic loss (``critic_loss.item()``), the B priorities handed to
``update_priorities``, the sampled indices and weight set (to check the RNG mirror), and every parameter tensor after the last update
(+ float64 sums of every tensor after each intermediate update).
Everything the consumer (tests/test_envelope_update_golden_gpu.py) needs besides this file derives from numpy PCG64 / MT19937 streams.
"""

from __future__ import annotations

import os
import sys
import time

import numpy as np
import torch as th

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_harness as rh  # noqa: E402
from tests.golden.envelope_update_cases import CASES, fill_agent, perturbed_target  # noqa: E402


class _RecordingF:
    """Stands in for ``torch.nn.functional`` inside the reference module to read the loss tensors it builds."""

    def __init__(self, real):
        self._real = real
        self.mse = []

    def __getattr__(self, k):
        return getattr(self._real, k)

    def mse_loss(self, *a, **k):
        out = self._real.mse_loss(*a, **k)
        self.mse.append(out.detach().clone())
        return out


def run_case(envm, name, out):
    c = CASES[name]
    th.manual_seed(0)
    env = rh.FakeEnv(obs_dim=c["obs"], n_actions=c["A"], reward_dim=c["D"])
    agent = envm.Envelope(env, batch_size=c["B"], num_sample_w=c["W"], per=True, buffer_size=c["N"], net_arch=c["net"], log=False,
                          seed=c["seed"], device="cpu", **c["kwargs"])
    fill_agent(agent, c)
    init = {k: v.detach().clone() for k, v in agent.q_net.state_dict().items()}
    for k, v in init.items():
        out[f"{name}/init/{k}"] = v.numpy().copy()
    agent.target_q_net.load_state_dict(perturbed_target(init))
    rec = _RecordingF(envm.F)
    envm.F = rec
    prios, inds = [], []
    orig_up = agent.replay_buffer.update_priorities
    orig_sample = agent.replay_buffer.sample

    def rec_up(idx, p):
        prios.append(np.asarray(p).copy())
        return orig_up(idx, p)

    def rec_sample(*a, **k):
        r = orig_sample(*a, **k)
        inds.append(np.asarray(r[-1]).copy())
        return r

    agent.replay_buffer.update_priorities = rec_up
    agent.replay_buffer.sample = rec_sample
    try:
        for step in range(c["steps"]):
            agent.global_step = c["global_step0"] + step
            np.random.seed(c["np_seed"] + step)
            lam = float(agent.homotopy_lambda)
            n0 = len(rec.mse)
            t0 = time.perf_counter()
            agent.update()
            dt = time.perf_counter() - t0
            mse = rec.mse[n0:]
            loss = mse[0] if lam <= 0 else (1 - lam) * mse[0] + lam * mse[1]
            out[f"{name}/step{step}/loss"] = np.float32(loss.item())
            out[f"{name}/step{step}/lambda"] = np.float64(lam)
            out[f"{name}/step{step}/priority"] = prios[-1].astype(np.float32)
            out[f"{name}/step{step}/inds"] = inds[-1].astype(np.int64)
            out[f"{name}/step{step}/param_sums"] = np.array([float(v.double().sum()) for v in agent.q_net.state_dict().values()])
            out[f"{name}/step{step}/param_abs_sums"] = np.array([float(v.double().abs().sum()) for v in agent.q_net.state_dict().values()])
            print(f"{name} step {step}: loss {loss.item():.8f} lambda {lam:.3f} mean prio {prios[-1].mean():.6f} ({dt:.1f} s)", flush=True)
    finally:
        envm.F = rec._real
    for k, v in agent.q_net.state_dict().items():
        out[f"{name}/final/{k}"] = v.detach().numpy().copy()
    out[f"{name}/tree_total"] = np.float64(agent.replay_buffer.tree.total_sum() if hasattr(agent.replay_buffer.tree, "total_sum")
                                           else agent.replay_buffer.tree.nodes[0][0])
    out[f"{name}/min_priority"] = np.float64(agent.replay_buffer.min_priority)


def main():
    assert rh.reference_available(), "needs /root/reference"
    th.set_num_threads(len(os.sched_getaffinity(0)))
    envm = rh.import_reference("morl_baselines.multi_policy.envelope.envelope")
    out = {}
    only = sys.argv[1:]
    path = os.path.join(HERE, "envelope_update.npz")
    if only and os.path.exists(path):
        out.update({k: v for k, v in np.load(path).items()})
    for name in CASES:
        if only and name not in only:
            continue
        for k in [k for k in out if k.startswith(name + "/")]:
            del out[k]
        run_case(envm, name, out)
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
