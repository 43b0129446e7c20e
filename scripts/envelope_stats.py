"""Per-phase cycle accounting of the tensor-core envelope kernel (MORL_ENVELOPE_STATS=1): where does a worker thread spend its time?"""
import ctypes, os, sys
os.environ["MORL_ENVELOPE_STATS"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch as th
from morl_baselines_b200 import ops, _lib
dev = th.device("cuda:0")
B, W, A, D = 1024, 64, 8, 3
g = th.Generator(device=dev).manual_seed(0)
sets = []
for _ in range(16):
    q_on = th.randn(B, W, A, D, device=dev, generator=g) * 3
    wset = th.rand(W, D, device=dev, generator=g); wset = wset / wset.sum(1, keepdim=True)
    sets.append((q_on, q_on + 0.05 * th.randn(B, W, A, D, device=dev, generator=g), wset, th.randn(B, D, device=dev, generator=g), th.zeros(B, device=dev)))
out = th.empty(W * B, D, device=dev)
lib = _lib.load()
def run(n):
    for i in range(n):
        ops.envelope_td(*sets[i % 16], 0.99, ops.DOT_UNFUSED, ops.ROWS_BMAJOR, want_indices=False, out=out)
run(32)
buf = (ctypes.c_ulonglong * 8)()
lib.morl_debug_envelope_stats(buf, 1)
n = 200
e0, e1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
e0.record(); run(n); e1.record(); th.cuda.synchronize()
lib.morl_debug_envelope_stats(buf, 1)
us = e0.elapsed_time(e1) / n * 1e3
ctas = min(ops.sm_count(), B)
c = [float(x) / ctas / n for x in buf]
print(f"launch {us:.2f} us (with stats overhead); per CTA ({ctas} CTAs, {B / ctas:.2f} transitions each), cycles:")
print(f"  converter  thread: waiting {c[0]:8.0f}  busy {c[1]:8.0f}")
print(f"  MMA        thread: waiting {c[2]:8.0f}")
print(f"  scanner    thread: waiting {c[3]:8.0f}  busy {c[4]:8.0f}")
print(f"  finisher   thread: waiting {c[5]:8.0f}  busy {c[6]:8.0f}")
print(f"  kernel (start -> last finisher iteration): {c[7]:8.0f}")
