"""Weight-conditioned MLP evaluated on the tcgen05 tensor cores (bf16x3 split operands, fp32-accurate).

``TCPairMlp`` runs the reference's ``mlp`` stack (common/networks.py:10-48: Linear -> ReLU ... -> Linear) for every
(observation b, weight vector j) pair of a minibatch without ever materialising fp32 activations in HBM:

    layer 1   : u = feats @ W1[:, :F]^T (B rows), v = wset @ W1[:, F:]^T + b1 (W rows)  -- two small library GEMMs --
                h1[b*W + j] = relu(u[b] + v[j]) written straight into bf16x3 planes (morl_pairs_relu_split_bf16x3);
    layers 2..: morl_gemm_bf16x3_f32 (TMA -> tcgen05.mma -> TMEM -> epilogue) with the activation re-split fused in the
                epilogue; the last layer writes fp32 Q-values.

Forward-only (no autograd) for now: it serves the two no-grad passes of the envelope target (online + target net on s').
"""

from __future__ import annotations

from typing import List, Optional

import torch as th
import torch.nn as nn

from . import ops


def _pad32(n: int) -> int:
    return (n + 31) // 32 * 32


class TCPairMlp:
    """Static plan (buffers + weight planes) for one nn.Sequential MLP and a fixed number of pair rows."""

    @staticmethod
    def supported(net: nn.Sequential) -> bool:
        mods = list(net)
        lin = [m for m in mods if isinstance(m, nn.Linear)]
        if len(lin) < 2 or any(not isinstance(m, (nn.Linear, nn.ReLU)) for m in mods):
            return False
        if not all(isinstance(mods[2 * i], nn.Linear) for i in range(len(lin))):
            return False
        hidden = [l.out_features for l in lin[:-1]]
        return all(h % 32 == 0 and h <= 256 for h in hidden) and lin[-1].out_features <= 256

    def __init__(self, net: nn.Sequential, feat_dim: int, n_obs: int, n_w: int):
        self.net = net
        self.lin: List[nn.Linear] = [m for m in net if isinstance(m, nn.Linear)]
        self.feat_dim = feat_dim
        self.B, self.W = n_obs, n_w
        dev = self.lin[0].weight.device
        M = n_obs * n_w
        self.h = [th.empty((3, M, l.out_features), device=dev, dtype=th.bfloat16) for l in self.lin[:-1]]
        self.wp = [th.empty((3, _pad32(l.out_features), l.in_features), device=dev, dtype=th.bfloat16) for l in self.lin[1:]]
        self.q = th.empty((M, self.lin[-1].out_features), device=dev, dtype=th.float32)

    def refresh_weights(self):
        """Re-split the (fp32) weights of layers 2.. into bf16x3 planes; call after every optimiser step / target sync."""
        for l, wp in zip(self.lin[1:], self.wp):
            ops.split_bf16x3(l.weight.detach(), rows_pad=wp.shape[1], ldp=wp.shape[2], out=wp)

    @th.no_grad()
    def forward_pairs(self, feats: th.Tensor, wset: th.Tensor) -> th.Tensor:
        """feats [B, F], wset [W, D] -> Q [B*W, out] (fp32, row b*W + j).  Uses the planes of the last refresh_weights()."""
        first = self.lin[0]
        u = feats @ first.weight[:, : self.feat_dim].t()
        v = th.addmm(first.bias, wset, first.weight[:, self.feat_dim :].t())
        a = ops.pairs_relu_split(u, v, out=self.h[0])
        n = len(self.lin)
        for k in range(1, n - 1):
            l = self.lin[k]
            _, a = ops.gemm_bf16x3(a, self.wp[k - 1], l.out_features, bias=l.bias, relu=True, out_f32=False, out_planes=True, c_planes=self.h[k])
        last = self.lin[-1]
        q, _ = ops.gemm_bf16x3(a, self.wp[n - 2], last.out_features, bias=last.bias, relu=False, out_f32=True, c_f32=self.q)
        return q
