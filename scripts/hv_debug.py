"""Trace of the hypervolume-parity training run (tests/test_hv_parity_gpu.py, seed 0) for a given engine configuration:
    python scripts/hv_debug.py [f16x2|bf16x3|notc] [steps]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch as th
from tests.golden.standin_env import TreasureChain
from morl_baselines_b200 import ops
from morl_baselines_b200.multi_policy.envelope.envelope import Envelope

mode = sys.argv[1] if len(sys.argv) > 1 else "f16x2"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
CHUNK = int(sys.argv[3]) if len(sys.argv) > 3 else 250
gold = json.load(open("tests/golden/hv_parity.json"))
hp = dict(gold["hyper_parameters"])
kw = dict(use_tensor_cores=False) if mode == "notc" else dict(tensor_core_format=mode)
th.manual_seed(0); np.random.seed(0)
env = TreasureChain(seed=0)
agent = Envelope(env, log=False, seed=0, device="cuda:0", **hp, **kw)
ws = [np.asarray(w, dtype=np.float32) for w in gold["eval_weights"]][:6]
done = 0
while done < steps:
    agent.train(total_timesteps=CHUNK, reset_num_timesteps=False)
    done += CHUNK
    obs, _ = env.reset()
    with th.no_grad():
        q = agent.q_net(th.as_tensor(obs).float().to(agent.device).unsqueeze(0).repeat(len(ws), 1), th.as_tensor(np.stack(ws)).to(agent.device))
    acts = [agent.eval(obs, w) for w in ws]
    pn = float(sum(p.detach().double().pow(2).sum() for p in agent.q_net.parameters()).sqrt())
    print(f"{mode} step {done}: loss {float(agent._last_loss) if agent._last_loss is not None else float('nan'):.5f} |theta| {pn:.4f} "
          f"Q range [{float(q.min()):.3f}, {float(q.max()):.3f}] finite {bool(th.isfinite(q).all())} acts {acts} min_p {agent.replay_buffer.min_priority:.4g} "
          f"overflow {ops.plane_overflow_count()}")

# final hypervolume, as the test computes it
from morl_baselines_b200.common.pareto import filter_pareto_dominated
from morl_baselines_b200.common.performance_indicators import hypervolume
from tests.golden.standin_env import HV_REF_POINT
env2 = TreasureChain(seed=123)
rets = []
for w in [np.asarray(w, dtype=np.float32) for w in gold["eval_weights"]]:
    obs, _ = env2.reset(); d, g, disc = False, 1.0, np.zeros(3)
    while not d:
        obs, r, term, trunc, _ = env2.step(agent.eval(obs, w)); disc += g * r; g *= hp["gamma"]; d = term or trunc
    rets.append(disc)
print(f"{mode} step {done}: final hv {hypervolume(HV_REF_POINT, list(filter_pareto_dominated(rets))):.4f} (reference {gold['seeds']['0']['hv']:.4f})")
