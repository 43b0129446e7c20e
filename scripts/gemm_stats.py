"""Per-role cycle accounting of the K-major bf16x3 GEMM (MORL_GEMM_STATS=1): where does the MMA thread wait?"""
import ctypes, os, sys
os.environ["MORL_GEMM_STATS"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch as th
from morl_baselines_b200 import ops, _lib
dev = th.device("cuda:0")
g = th.Generator(device=dev).manual_seed(0)
M, N, K = 65536, 256, 256
lib = _lib.load()
bp = ops.split_bf16x3(th.randn(N, K, device=dev, generator=g) / 16)
bias = th.randn(N, device=dev, generator=g)
sets = [ops.split_bf16x3(th.randn(M, K, device=dev, generator=g).relu_()) for _ in range(3)]
outs = [th.empty_like(sets[0]) for _ in range(3)]
def run(n):
    for i in range(n):
        ops.gemm_bf16x3(sets[i % 3], bp, N, bias=bias, relu=True, out_f32=False, out_planes=True, c_planes=outs[i % 3])
run(6)
buf = (ctypes.c_ulonglong * 8)()
lib.morl_debug_gemm_stats(buf, 1)
n = 30
e0, e1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
e0.record(); run(n); e1.record(); th.cuda.synchronize()
lib.morl_debug_gemm_stats(buf, 1)
us = e0.elapsed_time(e1) / n * 1e3
sm = ops.sm_count()
pairs = sm // 2
mhz = 1965.0
c = [float(x) for x in buf]
print(f"launch {us:.1f} us = {us * mhz:.0f} cycles @ {mhz:.0f} MHz (stats add a little overhead)")
print(f"MMA thread (per leader, per launch): total {c[2] / pairs / n:.0f} cyc, waiting TMA {c[0] / pairs / n:.0f}, waiting epilogue {c[1] / pairs / n:.0f}")
print(f"TMA thread (per CTA): waiting for a free stage {c[3] / sm / n:.0f} cyc")
print(f"epilogue warp (per CTA): waiting for accumulator {c[4] / sm / n:.0f} cyc, busy {c[5] / sm / n:.0f} cyc")
