"""Quick probe of the tcgen05 GEMM on a real B200: correctness on one shape + timing vs cuBLAS fp32 (CUDA events)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch as th
from morl_baselines_b200 import ops
dev = th.device("cuda:0")
g = th.Generator(device=dev).manual_seed(0)
M, N, K = 65536, 256, 256
a = th.randn(M, K, device=dev, generator=g); b = th.randn(N, K, device=dev, generator=g) / 16; bias = th.randn(N, device=dev, generator=g)
ap, bp = ops.split_bf16x3(a), ops.split_bf16x3(b)
c, cp = ops.gemm_bf16x3(ap, bp, N, bias=bias, relu=True, out_f32=True, out_planes=True)
th.cuda.synchronize()
ref = (a[:4096].double() @ b.double().t() + bias.double()).clamp_min(0)
print("max abs err", float((c[:4096].double() - ref).abs().max()), "ref max", float(ref.abs().max()))
def timeit(fn, n=20):
    for _ in range(3): fn()
    th.cuda.synchronize(); e0, e1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); th.cuda.synchronize(); return e0.elapsed_time(e1) / n * 1e3
t1 = timeit(lambda: ops.gemm_bf16x3(ap, bp, N, bias=bias, relu=True, out_f32=False, out_planes=True, c_planes=cp))
t2 = timeit(lambda: ops.gemm_bf16x3(ap, bp, N, bias=bias, relu=True, out_f32=True, out_planes=False, c_f32=c))
t3 = timeit(lambda: th.relu(th.addmm(bias, a, b.t())))
fl = 2.0 * M * N * K
print(f"tcgen05 bf16x3 -> planes: {t1:.1f} us ({6*fl/t1/1e6:.0f} TFLOP/s bf16-issued, {fl/t1/1e6:.1f} fp32-equivalent); -> f32: {t2:.1f} us; cuBLAS fp32 addmm+relu: {t3:.1f} us")
