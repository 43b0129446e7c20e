"""2-GPU parity check of DP-Envelope: the weight set of every update sharded over 2 ranks + ONE gradient all-reduce must
reproduce the single-GPU update -- same sampled indices (the priorities reach every rank through the collective), losses and priorities
to 1e-5 relative (the loss / gradient means are formed per shard and then averaged: a different fp32 summation order), parameters within
1e-5 |p| + 2e-6 after 4 updates incl. a target sync, and BIT-IDENTICAL parameters on the two ranks.

Not collected by pytest (needs 2 GPUs):   torchrun --nproc-per-node 2 --master-addr 127.0.0.1 tests/dp_envelope_check.py"""
import os
import sys

import numpy as np
import torch as th
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from morl_baselines_b200.multi_policy.envelope.envelope import Envelope  # noqa: E402
from morl_baselines_b200.testing import FakeEnv, synthetic_store  # noqa: E402


def run(dev, dp, graph, per_dev=False):
    OBS, A, D, B, W = 12, 4, 3, 64, 8
    th.manual_seed(5)
    np.random.seed(5)
    agent = Envelope(FakeEnv(obs_dim=OBS, n_actions=A, reward_dim=D), batch_size=B, num_sample_w=W, per=True, buffer_size=2048, net_arch=[64, 64, 64], log=False,
                     seed=5, device=dev, use_cuda_graph=graph, target_net_update_freq=3, dp_group=True if dp else None, per_on_device=per_dev)
    st = synthetic_store(1024, OBS, A, D, seed=2)
    rb = agent.replay_buffer
    rb.obs[:1024], rb.next_obs[:1024], rb.actions[:1024], rb.rewards[:1024], rb.dones[:1024] = st["obs"], st["next_obs"], st["actions"], st["rewards"], st["dones"]
    rb.size, rb.ptr = 1024, 0
    rb.mark_all_dirty()
    rb.tree.batch_set(np.arange(1024), np.linspace(0.1, 1.0, 1024))
    rec = []
    for step in range(4):
        np.random.seed(20 + step)
        agent.global_step = step + 1
        agent.update()
        rec.append((float(agent._last_loss), agent._last_inds.copy(), np.asarray(agent._last_priority).copy()))
    th.cuda.synchronize()
    return rec, [p.detach().clone() for p in agent.q_net.parameters()]


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dev = th.device("cuda", int(os.environ["LOCAL_RANK"]))
    th.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    ok = True
    for graph, per_dev in ((False, False), (True, False), (True, True)):  # (device-resident PER needs the captured step)
        rec1, p1 = run(dev, dp=False, graph=graph, per_dev=per_dev)
        rec2, p2 = run(dev, dp=True, graph=graph, per_dev=per_dev)
        for (l1, i1, q1), (l2, i2, q2) in zip(rec1, rec2):
            ok &= bool(np.array_equal(i1, i2)) and abs(l1 - l2) <= 1e-5 * abs(l1) and bool(np.allclose(q1, q2, rtol=1e-5, atol=1e-7))
        worst = 0.0
        for a, b in zip(p1, p2):
            ok &= bool(((a - b).abs() <= 1e-5 * a.abs() + 2e-6).all())
            worst = max(worst, float((a - b).abs().max()))
            mine = b.double().sum().reshape(1)
            lo, hi = mine.clone(), mine.clone()
            dist.all_reduce(lo, op=dist.ReduceOp.MIN)
            dist.all_reduce(hi, op=dist.ReduceOp.MAX)
            ok &= float(lo) == float(hi)
        if rank == 0:
            print(f"graph={graph} device_per={per_dev}: losses single {[round(r[0], 7) for r in rec1]} dp {[round(r[0], 7) for r in rec2]} max |dp - single| over parameters {worst:.2e}")
    flag = th.tensor([1.0 if ok else 0.0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("DP-ENVELOPE PARITY", "OK" if float(flag) == 1.0 else "FAILED")
    dist.destroy_process_group()
    sys.exit(0 if float(flag) == 1.0 else 1)


if __name__ == "__main__":
    main()
