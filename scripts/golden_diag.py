"""Error statistics of Envelope.update() against the frozen outputs of the unmodified reference (tests/golden/envelope_update.npz):
prints what tests/test_envelope_update_golden_gpu.py asserts, without stopping at the first bound.  usage: golden_diag.py [case ...]"""
import os
import sys

import numpy as np
import torch as th

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from morl_baselines_b200.testing import FakeEnv  # noqa: E402
from tests.golden.envelope_update_cases import CASES, fill_agent, perturbed_target  # noqa: E402


def run(name, tc, graph, **kw):
    from morl_baselines_b200.multi_policy.envelope.envelope import Envelope

    g = np.load(os.path.join(ROOT, "tests", "golden", "envelope_update.npz"))
    c = CASES[name]
    cuda = th.device("cuda:0")
    agent = Envelope(FakeEnv(obs_dim=c["obs"], n_actions=c["A"], reward_dim=c["D"]), batch_size=c["B"], num_sample_w=c["W"], per=True,
                     buffer_size=c["N"], net_arch=c["net"], log=False, seed=c["seed"], device=cuda, use_cuda_graph=graph, use_tensor_cores=tc,
                     **c["kwargs"], **kw)
    fill_agent(agent, c)
    init = {k: th.from_numpy(g[f"{name}/init/{k}"]) for k in agent.q_net.state_dict()}
    agent.q_net.load_state_dict(init)
    agent.target_q_net.load_state_dict(perturbed_target(init))
    for step in range(c["steps"]):
        agent.global_step = c["global_step0"] + step
        np.random.seed(c["np_seed"] + step)
        agent.update()
        inds_ok = np.array_equal(agent._last_inds, g[f"{name}/step{step}/inds"])
        loss, ref = float(agent._last_loss), float(g[f"{name}/step{step}/loss"])
        pr, pref = agent._last_priority, g[f"{name}/step{step}/priority"]
        perr = np.abs(pr - pref)
        sums = np.array([float(v.double().sum()) for v in agent.q_net.state_dict().values()])
        serr = np.abs(sums - g[f"{name}/step{step}/param_sums"]).max() / g[f"{name}/step{step}/param_abs_sums"].max()
        print(f"{name} tc={tc} graph={graph} step {step}: inds_ok={inds_ok} loss {loss:.8f} ref {ref:.8f} rel {abs(loss-ref)/abs(ref):.2e} | prio max abs {perr.max():.2e} "
              f"max rel {np.max(perr/np.abs(pref)):.2e} viol(1e-5,2e-6) {int((perr > 1e-5*np.abs(pref)+2e-6).sum())} | param-sum err/abs-sum {serr:.2e}")
    lr = agent.learning_rate
    for k, v in agent.q_net.state_dict().items():
        ref = g[f"{name}/final/{k}"]
        err = np.abs(v.cpu().numpy() - ref)
        ok = err <= 1e-5 * np.abs(ref) + 2e-6
        print(f"   {k:14s} frac ok {ok.mean():.5f}  max err {err.max():.2e} ({err.max()/lr:.2f} lr)  p99.9 {np.quantile(err, 0.999):.2e}  mean {err.mean():.2e}")
    print(f"   min_priority {agent.replay_buffer.min_priority:.8g} ref {float(g[f'{name}/min_priority']):.8g}")


if __name__ == "__main__":
    names = sys.argv[1:] or list(CASES)
    for n in names:
        for tc in (True, False):
            for graph in (True, False):
                try:
                    run(n, tc, graph)
                except Exception as e:  # noqa: BLE001
                    print(f"{n} tc={tc} graph={graph}: EXCEPTION {type(e).__name__}: {e}")
