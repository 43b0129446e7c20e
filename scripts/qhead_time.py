"""In-graph launch time of the fused output-layer + envelope + Bellman kernel (morl_qhead_envelope_td_f32) against the three-launch chain
it replaces, at the north-star shape, on rotating activation sets (> L2), CUDA events around graph replays:
   python scripts/qhead_time.py [B W A D K]"""
import os, sys

import torch as th

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from morl_baselines_b200 import ops  # noqa: E402


def timed(fn, nsets, replays=10):
    side = th.cuda.Stream()
    side.wait_stream(th.cuda.current_stream())
    with th.cuda.stream(side):
        for i in range(nsets):
            fn(i)
    th.cuda.current_stream().wait_stream(side)
    th.cuda.synchronize()
    g = th.cuda.CUDAGraph()
    with th.cuda.graph(g):
        for i in range(nsets):
            fn(i)
    for _ in range(3):
        g.replay()
    th.cuda.synchronize()
    e0, e1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(replays):
        g.replay()
    e1.record()
    th.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (replays * nsets)


def main():
    B, W, A, D, K = (int(x) for x in sys.argv[1:6]) if len(sys.argv) >= 6 else (1024, 64, 8, 3, 256)
    dev = th.device("cuda:0")
    fmt = ops.FMT_F16X2
    g = th.Generator(device=dev).manual_seed(0)
    M, N = B * W, A * D
    s_act, s_w = ops.scale_tensor(2.0, dev), ops.scale_tensor(4096.0, dev)
    nsets = 4
    a_on = [ops.split_planes(th.randn(M, K, device=dev, generator=g).relu_(), fmt, rows_pad=M, ldp=K, scale=s_act) for _ in range(nsets)]
    a_tg = [ops.split_planes(th.randn(M, K, device=dev, generator=g).relu_(), fmt, rows_pad=M, ldp=K, scale=s_act) for _ in range(nsets)]
    p_on = ops.split_planes(th.randn(N, K, device=dev, generator=g) / 16, fmt, rows_pad=32, ldp=K, scale=s_w)
    p_tg = ops.split_planes(th.randn(N, K, device=dev, generator=g) / 16, fmt, rows_pad=32, ldp=K, scale=s_w)
    b_on, b_tg = th.randn(N, device=dev, generator=g), th.randn(N, device=dev, generator=g)
    wset = th.rand(W, D, device=dev, generator=g)
    wset = wset / wset.sum(1, keepdim=True)
    rew, done = th.randn(B, D, device=dev, generator=g), (th.rand(B, device=dev, generator=g) < 0.02).float()
    out = th.empty(W * B, D, device=dev)
    q1, q2 = th.empty(M, N, device=dev), th.empty(M, N, device=dev)

    def fused(i):
        ops.qhead_envelope_td(a_on[i], a_tg[i], p_on, p_tg, b_on, b_tg, wset, rew, done, 0.99, B, W, A, D, ops.DOT_UNFUSED, ops.ROWS_BMAJOR, a_scale_on=s_act,
                              a_scale_tg=s_act, w_scale_on=s_w, w_scale_tg=s_w, out=out)

    def chain(i):
        ops.gemm_planes(a_on[i], p_on, N, bias=b_on, c_f32=q1, a_scale=s_act, b_scale=s_w)
        ops.gemm_planes(a_tg[i], p_tg, N, bias=b_tg, c_f32=q2, a_scale=s_act, b_scale=s_w)
        ops.envelope_td(q1.view(B, W, A, D), q2.view(B, W, A, D), wset, rew, done, 0.99, ops.DOT_UNFUSED, ops.ROWS_BMAJOR, want_indices=False, out=out)

    def gemm_only(i):
        ops.gemm_planes(a_on[i], p_on, N, bias=b_on, c_f32=q1, a_scale=s_act, b_scale=s_w)

    t_f = [timed(fused, nsets) for _ in range(3)]
    t_c = [timed(chain, nsets) for _ in range(3)]
    t_g = [timed(gemm_only, nsets) for _ in range(3)]
    plane_bytes = 2 * 4 * M * K
    print(f"shape B={B} W={W} A={A} D={D} K={K}: activation planes of both nets {plane_bytes / 1e6:.1f} MB")
    print("fused  us/launch:", " ".join(f"{x:.2f}" for x in t_f), f" -> {plane_bytes / min(t_f) / 1e3:.0f} GB/s of plane bytes")
    print("chain  us (2 GEMM + envelope):", " ".join(f"{x:.2f}" for x in t_c))
    print("one output-layer GEMM us:", " ".join(f"{x:.2f}" for x in t_g))


if __name__ == "__main__":
    main()
