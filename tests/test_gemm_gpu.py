"""GPU tests of the tcgen05 bf16x3 GEMM (csrc/gemm_bf16x3.cu) against a float64 reference.

Tolerance: the six-term bf16x3 expansion is exact to ~2^-24 per product and accumulates in fp32, so the result must agree
with the float64 product to fp32-GEMM accuracy: |err| <= 2e-6 * (|A| . |B|^T) elementwise (2e-6 ~ 32 ulp of headroom for the
K = 256 accumulation; a plain BF16 or TF32 GEMM misses this bound by two to three orders of magnitude)."""

import numpy as np
import pytest
import torch as th

pytestmark = pytest.mark.gpu


def _ref(a, b, bias):
    r = a.double() @ b.double().t()
    if bias is not None:
        r = r + bias.double()
    return r


def _bound(a, b):
    return 2e-6 * (a.abs().double() @ b.abs().double().t()) + 1e-30


def test_split_planes_are_exact_to_2pow24(cuda):
    from morl_baselines_b200 import ops

    g = th.Generator(device=cuda).manual_seed(0)
    x = th.randn(300, 70, device=cuda, generator=g) * th.exp(3 * th.randn(300, 70, device=cuda, generator=g))
    p = ops.split_bf16x3(x, rows_pad=320, ldp=96)
    s = p[0].double() + p[1].double() + p[2].double()
    assert th.all(s[300:] == 0) and th.all(s[:, 70:] == 0)
    assert float(((s[:300, :70] - x.double()).abs() / x.abs().double()).max()) <= 2.0**-23
    pt = ops.split_bf16x3(x, transpose=True)
    assert th.equal(pt[0][:70, :300], p[0][:300, :70].t())


@pytest.mark.parametrize("M,N,K", [(128, 256, 256), (1000, 256, 256), (65536, 256, 256), (4096, 24, 256), (777, 256, 32), (513, 64, 64)])
@pytest.mark.parametrize("relu", [False, True])
def test_gemm_bf16x3_matches_float64(cuda, M, N, K, relu):
    from morl_baselines_b200 import ops

    g = th.Generator(device=cuda).manual_seed(M + N + K)
    a = th.randn(M, K, device=cuda, generator=g)
    b = th.randn(N, K, device=cuda, generator=g) / np.sqrt(K)
    bias = th.randn(N, device=cuda, generator=g)
    ap = ops.split_bf16x3(a)
    bp = ops.split_bf16x3(b, rows_pad=(N + 31) // 32 * 32)
    c, cp = ops.gemm_bf16x3(ap, bp, N, bias=bias, relu=relu, out_f32=True, out_planes=(N % 32 == 0))
    ref = _ref(a, b, bias)
    if relu:
        ref = ref.clamp_min(0)
    err = (c.double() - ref).abs()
    bound = _bound(a, b) + 2e-6 * bias.abs().double()
    assert bool((err <= bound).all()), float((err / bound).max())
    # a plain fp32 cuBLAS product sits inside the same bound (sanity of the bound itself)
    c32 = th.addmm(bias, a, b.t())
    c32 = c32.clamp_min(0) if relu else c32
    assert bool(((c32.double() - ref).abs() <= bound).all())
    if cp is not None:
        s = cp[0].double() + cp[1].double() + cp[2].double()
        assert float((s - c.double()).abs().max()) <= 2.0**-22 * float(c.abs().max())


def test_gemm_relu_mask_and_chaining(cuda):
    """Two chained layers through the plane format (no fp32 round trip) and the ReLU-backward mask."""
    from morl_baselines_b200 import ops

    g = th.Generator(device=cuda).manual_seed(5)
    M, H = 2048, 256
    x = th.randn(M, H, device=cuda, generator=g)
    w1 = th.randn(H, H, device=cuda, generator=g) / 16
    w2 = th.randn(H, H, device=cuda, generator=g) / 16
    b1 = th.randn(H, device=cuda, generator=g) * 0.1
    xp = ops.split_bf16x3(x)
    _, h1p = ops.gemm_bf16x3(xp, ops.split_bf16x3(w1), H, bias=b1, relu=True, out_f32=False, out_planes=True)
    y, _ = ops.gemm_bf16x3(h1p, ops.split_bf16x3(w2), H)
    h1 = (x.double() @ w1.double().t() + b1.double()).clamp_min(0)
    ref = h1 @ w2.double().t()
    assert float((y.double() - ref).abs().max()) <= 1e-5 * float(ref.abs().max())
    # backward of layer 2 w.r.t. h1, masked by relu'(h1):  dH = (dY . W2) * [h1 > 0]
    dy = th.randn(M, H, device=cuda, generator=g)
    dh, _ = ops.gemm_bf16x3(ops.split_bf16x3(dy), ops.split_bf16x3(w2, transpose=True), H, relu_mask=h1p)
    ref_dh = (dy.double() @ w2.double()) * (h1 > 0)
    assert float((dh.double() - ref_dh).abs().max()) <= 1e-5 * float(ref_dh.abs().max())


def test_pairs_relu_split(cuda):
    from morl_baselines_b200 import ops

    g = th.Generator(device=cuda).manual_seed(1)
    u, v = th.randn(37, 256, device=cuda, generator=g), th.randn(5, 256, device=cuda, generator=g)
    p = ops.pairs_relu_split(u, v)
    ref = (u.unsqueeze(1) + v.unsqueeze(0)).clamp_min(0).view(-1, 256)
    s = p[0].double() + p[1].double() + p[2].double()
    assert float((s - ref.double()).abs().max()) <= 2.0**-22 * float(ref.abs().max())


@pytest.mark.parametrize("M,gc,hc", [(65536, 256, 256), (5000, 256, 256), (4096, 24, 256), (1000, 128, 64), (333, 256, 192)])
def test_gemm_mn_weight_gradient(cuda, M, gc, hc):
    """dW[n, k] = sum_m G[m, n] H[m, k] (MN-major operands, split-K) against float64."""
    from morl_baselines_b200 import ops

    g_ = th.Generator(device=cuda).manual_seed(M + gc)
    G = th.randn(M, gc, device=cuda, generator=g_)
    H = th.randn(M, hc, device=cuda, generator=g_).clamp_min(0)
    Gp = ops.split_bf16x3(G, ldp=(gc + 63) // 64 * 64)
    Hp = ops.split_bf16x3(H, ldp=(hc + 63) // 64 * 64)
    dW = ops.gemm_bf16x3_mn(Gp, gc, Hp, hc)
    ref = G.double().t() @ H.double()
    bound = 2e-6 * (G.abs().double().t() @ H.abs().double()) * max(1.0, np.sqrt(M / 4096)) + 1e-30
    err = (dW.double() - ref).abs()
    assert bool((err <= bound).all()), float((err / bound).max())
    dWt = ops.gemm_bf16x3_mn(Gp, gc, Hp, hc, transpose_out=True)
    assert th.equal(dWt, dW.t().contiguous())
    cs = ops.colsum_bf16x3(Gp, gc)
    np.testing.assert_allclose(cs.cpu().numpy(), G.double().sum(0).cpu().numpy(), rtol=1e-5, atol=1e-5 * float(G.abs().sum(0).max()))
    # bias gradient fused into the same pass (G^T . ones on the tensor cores); the weight gradient must be unchanged by it
    cs2 = th.full((gc,), float("nan"), device=cuda)
    dW2 = ops.gemm_bf16x3_mn(Gp, gc, Hp, hc, colsum=cs2)
    assert th.equal(dW2, dW)
    np.testing.assert_allclose(cs2.cpu().numpy(), G.double().sum(0).cpu().numpy(), rtol=1e-5, atol=1e-5 * float(G.abs().sum(0).max()))


@pytest.mark.parametrize("B,W,F,D,H", [(1024, 64, 32, 3, 256), (37, 5, 11, 2, 64), (3, 70, 7, 4, 300), (8, 8, 59, 3, 32)])
def test_pair_layer1_kernels(cuda, B, W, F, D, H):
    """Separable first layer on the pair batch: u = feats W1_s^T, v = wset W1_w^T + b1 in one launch, against float64; fp32 FMA
    chains -> 1e-6 of the magnitude sums."""
    from morl_baselines_b200 import ops

    g_ = th.Generator(device=cuda).manual_seed(B + 3 * W + F)
    feats, wset = th.randn(B, F, device=cuda, generator=g_), th.rand(W, D, device=cuda, generator=g_)
    W1, b1 = th.randn(H, F + D, device=cuda, generator=g_) / 4, th.randn(H, device=cuda, generator=g_)
    u, v = ops.pair_layer1_uv(feats, wset, W1, b1)
    ru = feats.double() @ W1[:, :F].double().t()
    rv = wset.double() @ W1[:, F:].double().t() + b1.double()
    assert float((u.double() - ru).abs().max()) <= 1e-6 * float((feats.abs().double() @ W1[:, :F].abs().double().t()).max())
    assert float((v.double() - rv).abs().max()) <= 1e-6 * float((wset.abs().double() @ W1[:, F:].abs().double().t() + b1.abs().double()).max())


@pytest.mark.parametrize("B,W,F,D,H", [(1024, 64, 32, 3, 256), (256, 32, 7, 3, 256), (37, 5, 11, 2, 64), (3, 70, 7, 4, 300), (8, 8, 59, 3, 32)])
def test_pair_layer1_grad(cuda, B, W, F, D, H):
    """dW1 = [dU^T feats | dV^T wset], db1 = colsum(dV) in one launch (split reduction, fixed summation order), against float64;
    run twice on the same workspace: the arrival counters must reset themselves and the result must be bit-identical."""
    from morl_baselines_b200 import ops

    g_ = th.Generator(device=cuda).manual_seed(B + 3 * W + F)
    feats, wset = th.randn(B, F, device=cuda, generator=g_), th.rand(W, D, device=cuda, generator=g_)
    dU, dV = th.randn(B, H, device=cuda, generator=g_), th.randn(W, H, device=cuda, generator=g_)
    ws = ops.pair_layer1_grad_workspace(F, D, H, cuda)
    dW1, db1 = ops.pair_layer1_grad(dU, dV, feats, wset, workspace=ws)
    ref_w = th.cat([dU.double().t() @ feats.double(), dV.double().t() @ wset.double()], dim=1)
    mag_w = th.cat([dU.abs().double().t() @ feats.abs().double(), dV.abs().double().t() @ wset.abs().double()], dim=1)
    assert float((dW1.double() - ref_w).abs().max()) <= 2e-6 * float(mag_w.max())
    assert float((db1.double() - dV.double().sum(0)).abs().max()) <= 2e-6 * float(dV.abs().double().sum(0).max())
    dW1b, db1b = ops.pair_layer1_grad(dU, dV, feats, wset, workspace=ws)
    assert th.equal(dW1, dW1b) and th.equal(db1, db1b)


def test_split_vectorised_path_matches_scalar_path(cuda):
    """ldp % 8 == 0 takes the 8-columns-per-thread kernel; an unaligned source (ld_src % 4 != 0) and ragged columns must give the same
    planes as the transposed-input scalar kernel."""
    from morl_baselines_b200 import ops

    g_ = th.Generator(device=cuda).manual_seed(9)
    for rows, cols in ((65536, 24), (1000, 250), (77, 13)):
        x = th.randn(rows, cols, device=cuda, generator=g_)
        ldp = (cols + 31) // 32 * 32
        a = ops.split_bf16x3(x, ldp=ldp)  # vectorised
        b = ops.split_bf16x3(x.t().contiguous(), ldp=ldp, transpose=True)  # scalar kernel on the transposed source
        assert th.equal(a, b)
        assert float((a[0].double() + a[1].double() + a[2].double())[:, :cols].sub(x.double()).abs().max()) <= 2.0**-23 * float(x.abs().max())
        assert float(a[:, :, cols:].abs().max()) == 0.0


def test_pairs_grad_reduce(cuda):
    from morl_baselines_b200 import ops

    g_ = th.Generator(device=cuda).manual_seed(2)
    for B, W in ((37, 5), (64, 64), (300, 33), (5, 70)):
        G = th.randn(B * W, 256, device=cuda, generator=g_)
        Gp = ops.split_bf16x3(G)
        dU, dV = ops.pairs_grad_reduce(Gp, B, W)
        ref = G.double().view(B, W, 256)
        np.testing.assert_allclose(dU.cpu().numpy(), ref.sum(1).cpu().numpy(), rtol=1e-5, atol=1e-5)
        np.testing.assert_allclose(dV.cpu().numpy(), ref.sum(0).cpu().numpy(), rtol=1e-5, atol=1e-5)


def test_gemm_single_cta_kernel_still_correct(cuda):
    """Large shapes use the CTA-pair (cta_group::2) kernel; MORL_GEMM_FORCE_1CTA=1 keeps the one-CTA kernel alive as a cross-check.
    The switch is read once per process, hence the subprocess."""
    import os
    import subprocess
    import sys

    code = (
        "import torch as th, numpy as np\n"
        "from morl_baselines_b200 import ops\n"
        "g = th.Generator(device='cuda').manual_seed(3)\n"
        "a = th.randn(5000, 256, device='cuda', generator=g); b = th.randn(256, 256, device='cuda', generator=g) / 16\n"
        "c, _ = ops.gemm_bf16x3(ops.split_bf16x3(a), ops.split_bf16x3(b), 256)\n"
        "ref = a.double() @ b.double().t()\n"
        "err = float((c.double() - ref).abs().max()); print('ERR', err); assert err < 2e-5\n"
    )
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    outs = {}
    for flag in ("1", "0"):
        env = dict(os.environ, MORL_GEMM_FORCE_1CTA=flag, PYTHONPATH=root)
        r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stdout + r.stderr
        outs[flag] = r.stdout
    assert "ERR" in outs["1"] and "ERR" in outs["0"]


def test_reverse_tile_order_and_multi_split_are_bit_identical(cuda):
    from morl_baselines_b200 import ops

    g = th.Generator(device=cuda).manual_seed(11)
    a = th.randn(40000, 256, device=cuda, generator=g)
    ws = [th.randn(256, 256, device=cuda, generator=g) / 16, th.randn(24, 256, device=cuda, generator=g), th.randn(256, 64, device=cuda, generator=g)]
    ap = ops.split_bf16x3(a)
    singles = [ops.split_bf16x3(ws[0]), ops.split_bf16x3(ws[1], rows_pad=32), ops.split_bf16x3(ws[2], rows_pad=64, ldp=256, transpose=True)]
    multi = [th.empty_like(s) for s in singles]
    ops.split_bf16x3_multi([(ws[0], multi[0], False), (ws[1], multi[1], False), (ws[2], multi[2], True)])
    for s, m in zip(singles, multi):
        assert th.equal(s, m)
    c0, p0 = ops.gemm_bf16x3(ap, singles[0], 256, relu=True, out_f32=True, out_planes=True)
    c1, p1 = ops.gemm_bf16x3(ap, singles[0], 256, relu=True, out_f32=True, out_planes=True, reverse_tiles=True)
    assert th.equal(c0, c1) and th.equal(p0, p1)
