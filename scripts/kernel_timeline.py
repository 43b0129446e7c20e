"""Per-kernel durations of the captured Envelope update in situ (warm, inside graph replays), from CUPTI through torch.profiler:
   python scripts/kernel_timeline.py [n_updates]     (north-star shape; not a bench value -- the profiler adds host overhead)"""
import os, sys
from collections import OrderedDict

import numpy as np
import torch as th

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    dev = th.device("cuda:0")
    from morl_baselines_b200.testing import synthetic_store

    agent = bench._make_agent(dev, 0, True)
    bench._fill_store(agent.replay_buffer, synthetic_store(bench.STORE, bench.OBS, bench.A, bench.D, seed=0))
    agent.global_step = 1
    for _ in range(6):
        agent.update()
    th.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile

    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(n):
            agent.update()
        th.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == th.autograd.DeviceType.CUDA]
    evs.sort(key=lambda e: e.time_range.start)
    tot = OrderedDict()
    for e in evs:
        k = e.name[:110]
        d = tot.setdefault(k, [0, 0.0])
        d[0] += 1
        d[1] += e.device_time if hasattr(e, "device_time") else e.cuda_time
    span = (evs[-1].time_range.end - evs[0].time_range.start) / n
    print(f"# {n} updates, {len(evs)} device activities, wall span per update {span:.1f} us")
    s = 0.0
    for k, (c, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        print(f"{t / n:9.1f} us/update {c / n:6.1f}x  avg {t / c:7.2f} us  {k}")
        s += t / n
    print(f"# sum of kernel time per update {s:.1f} us")
    # the sequence of one update (the last), in launch order
    per = len(evs) // n
    print("# launch order of the last update:")
    for e in evs[-per:]:
        d = e.device_time if hasattr(e, "device_time") else e.cuda_time
        print(f"   {d:7.2f} us  {e.name[:100]}")


if __name__ == "__main__":
    main()
