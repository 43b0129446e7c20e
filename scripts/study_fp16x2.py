"""Numerical feasibility of the fp16 x 2 operand split against the bf16 x 3 split the GEMMs use today.

CPU-only simulation on the tensors of a real Envelope update (reference-shaped 4 x 256 Q-network, a 4,096-row sample of the pair batch):
every term of a split is an exactly representable fp32 value, every product of two terms is exact in fp32, so an fp32 matmul of the term
matrices reproduces the tensor-core arithmetic up to accumulation order.  Reported: max |C - C_exact| / (|A| |B|) over the output, for
  fp32        : plain fp32 GEMM (what cuBLAS SGEMM / the reference's MKL path delivers)
  bf16x3 (6)  : A0B0 + A0B1 + A1B0 + A1B1 + A0B2 + A2B0           (today: 6 MMAs, 6 B / element)
  fp16x2 (3)  : A0B0 + A0B1 + A1B0 with per-tensor power-of-two scaling into fp16's normal range   (3 MMAs, 4 B / element)
  fp16x2 static: the same with FIXED scales (activations 2^6, weights 2^8, gradients 2^22) -- no amax pass, but small values lose bits to
                 fp16's denormal range
"""
import numpy as np
import torch as th

th.manual_seed(0)
rng = np.random.default_rng(0)


def split_bf16(x, n):
    terms, r = [], x.clone()
    for _ in range(n):
        t = r.to(th.bfloat16).to(th.float32)
        terms.append(t)
        r = r - t
    return terms


def split_fp16(x, n, scale):
    terms, r = [], x * scale
    for _ in range(n):
        t = r.to(th.float16).to(th.float32)
        terms.append(t)
        r = r - t
    return terms


def pow2_scale(x, target=2.0**10):
    """power of two that brings max|x| to ~target (fp16 max is 65504; the residual term then sits ~2^-11 below, still normal for the bulk)."""
    m = float(x.abs().max())
    return 2.0 ** np.floor(np.log2(target / m)) if m > 0 else 1.0


def report(name, A, B, static=None):
    exact = A.double() @ B.double().t()
    bound = A.abs().double() @ B.abs().double().t() + 1e-300
    out = {}
    out["fp32"] = (A @ B.t()).double()
    a, b = split_bf16(A, 3), split_bf16(B, 3)
    out["bf16x3 (6)"] = sum((a[i] @ b[j].t()) for i, j in ((2, 0), (0, 2), (1, 1), (1, 0), (0, 1), (0, 0))).double()
    sa, sb = pow2_scale(A), pow2_scale(B)
    a, b = split_fp16(A, 2, sa), split_fp16(B, 2, sb)
    out["fp16x2 (3)"] = (sum((a[i] @ b[j].t()) for i, j in ((1, 0), (0, 1), (0, 0))).double()) / (sa * sb)
    if static is not None:  # fixed scales chosen without looking at the data (no amax pass): how much accuracy do they cost?
        a, b = split_fp16(A, 2, static[0]), split_fp16(B, 2, static[1])
        out["fp16x2 static"] = (sum((a[i] @ b[j].t()) for i, j in ((1, 0), (0, 1), (0, 0))).double()) / (static[0] * static[1])
    row = "  ".join(f"{k}: {float(((v - exact).abs() / bound).max()):.2e}" for k, v in out.items())
    print(f"{name:34s} {row}   (scales 2^{int(np.log2(sa))}, 2^{int(np.log2(sb))})")


M, H = 4096, 256
lin = [th.nn.Linear(35, H)] + [th.nn.Linear(H, H) for _ in range(3)] + [th.nn.Linear(H, 24)]
for l in lin:
    th.nn.init.orthogonal_(l.weight)
    th.nn.init.zeros_(l.bias)
x = th.cat([th.randn(M, 32), th.from_numpy(rng.dirichlet(np.ones(3), M).astype(np.float32))], 1)
hs, h = [], x
for l in lin[:-1]:
    h = th.relu(l(h)).detach()
    hs.append(h)
q = lin[-1](h).detach()
# gradient seed of the TD loss: 2 (q - target) / N on the taken action's sector, zero elsewhere
g = th.zeros_like(q)
act = th.from_numpy(rng.integers(0, 8, M))
for d in range(3):
    g[th.arange(M), act * 3 + d] = 2.0 * th.randn(M) / (65536 * 3)
print("forward layers  C = H_{k-1} W_k^T")
for k in (1, 2, 3):
    report(f"  layer {k + 1} forward", hs[k - 1], lin[k].weight.detach(), static=(2.0**6, 2.0**8))
report("  output layer forward", hs[3], lin[4].weight.detach(), static=(2.0**6, 2.0**8))
print("backward  dX = G W,  dW = G^T H")
G = g
for k in (4, 3, 2):
    W = lin[k].weight.detach()
    # static gradient scale 2^22: the seed is 2 delta / (B W d) with |delta| <~ 10, i.e. |g| <~ 1e-4
    report(f"  dX through layer {k + 1}", G, W.t().contiguous(), static=(2.0**22, 2.0**8))
    report(f"  dW of layer {k + 1}", G.t().contiguous(), hs[k - 1].t().contiguous(), static=(2.0**22, 2.0**6))
    G = (G @ W) * (hs[k - 1] > 0)
