"""Per-role cycle accounting of the K-major split-operand GEMM (MORL_GEMM_STATS=1): where does the MMA thread wait?
usage: gemm_stats.py [f16x2|bf16x3]"""
import ctypes, os, sys
os.environ["MORL_GEMM_STATS"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch as th
from morl_baselines_b200 import ops, _lib
dev = th.device("cuda:0")
g = th.Generator(device=dev).manual_seed(0)
M, N, K = 65536, 256, 256
lib = _lib.load()
fmt = ops.FMT_BF16X3 if (len(sys.argv) > 1 and sys.argv[1] == "bf16x3") else ops.FMT_F16X2
sa = ops.scale_tensor(8.0, dev) if fmt == ops.FMT_F16X2 else None
sb = ops.scale_tensor(2048.0, dev) if fmt == ops.FMT_F16X2 else None
bp = ops.split_planes(th.randn(N, K, device=dev, generator=g) / 16, fmt, scale=sb)
bias = th.randn(N, device=dev, generator=g)
sets = [ops.split_planes(th.randn(M, K, device=dev, generator=g).relu_(), fmt, scale=sa) for _ in range(3)]
outs = [th.empty_like(sets[0]) for _ in range(3)]
def run(n):
    for i in range(n):
        ops.gemm_planes(sets[i % 3], bp, N, bias=bias, relu=True, out_f32=False, out_planes=True, c_planes=outs[i % 3], a_scale=sa, b_scale=sb, c_scale=sa)
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
print(f"format {'bf16x3' if fmt == ops.FMT_BF16X3 else 'f16x2'}, accumulators {'split' if os.environ.get('MORL_GEMM_SPLIT_ACC') == '1' else 'single'}")
print(f"launch {us:.1f} us = {us * mhz:.0f} cycles @ {mhz:.0f} MHz (stats add a little overhead)")
print(f"MMA thread (per leader, per launch): total {c[2] / pairs / n:.0f} cyc, waiting TMA {c[0] / pairs / n:.0f}, waiting epilogue {c[1] / pairs / n:.0f}")
print(f"TMA thread (per CTA): waiting for a free stage {c[3] / sm / n:.0f} cyc")
print(f"epilogue warp (per CTA): waiting for accumulator {c[4] / sm / n:.0f} cyc, busy {c[5] / sm / n:.0f} cyc")
