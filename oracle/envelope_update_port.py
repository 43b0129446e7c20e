"""TEST / BASELINE INFRASTRUCTURE ONLY -- PyTorch-CPU restatement ("port") of one gradient update of the reference's
Envelope Q-learning, used (a) as the CPU baseline timed by bench.py beside the CUDA engine and (b) by the differential
tests that compare a whole update of morl_baselines_b200.Envelope with the reference's arithmetic.

It follows the reference's own structure, including its |W|-fold redundancy (SURVEY.md headline 2):
  * Envelope.update           multi_policy/envelope/envelope.py:266-334   (tile the minibatch |W| times, :284-291)
  * Envelope.envelope_target  envelope.py:404-440                          (tile |W| times AGAIN, :416-418; online net
                                                                            selects, target net evaluates)
  * QNet                      envelope.py:33-77, common/networks.py:10-48, 143-157 (orthogonal init, zero bias)
When /root/reference is mounted, ``make_reference_agent`` returns the UNMODIFIED reference class instead (kind="reference").
"""

from __future__ import annotations

import numpy as np
import torch as th
import torch.nn as nn
import torch.nn.functional as F


def build_qnet(obs_dim, n_actions, rew_dim, net_arch):
    layers, d = [], obs_dim + rew_dim
    for h in net_arch:
        layers += [nn.Linear(d, h), nn.ReLU()]
        d = h
    layers.append(nn.Linear(d, n_actions * rew_dim))
    net = nn.Sequential(*layers)
    for m in net:
        if isinstance(m, nn.Linear):
            nn.init.orthogonal_(m.weight, gain=1)
            nn.init.constant_(m.bias, 0.0)
    return net


class EnvelopeUpdatePort:
    def __init__(self, obs_dim, n_actions, rew_dim, net_arch=(256, 256, 256, 256), lr=3e-4, gamma=0.99, max_grad_norm=1.0, seed=0,
                 state_dict=None):
        th.manual_seed(seed)
        self.A, self.D, self.gamma, self.max_grad_norm = n_actions, rew_dim, gamma, max_grad_norm
        self.q_net = build_qnet(obs_dim, n_actions, rew_dim, list(net_arch))
        self.target_q_net = build_qnet(obs_dim, n_actions, rew_dim, list(net_arch))
        if state_dict is not None:
            self.q_net.load_state_dict({k.replace("net.", "", 1): v for k, v in state_dict.items()})
        self.target_q_net.load_state_dict(self.q_net.state_dict())
        for p in self.target_q_net.parameters():
            p.requires_grad = False
        self.optim = th.optim.Adam(self.q_net.parameters(), lr=lr)

    def _q(self, net, obs, w):
        return net(th.cat((obs, w), dim=1)).view(-1, self.A, self.D)

    @th.no_grad()
    def envelope_target(self, obs, w, sampled_w):
        n_w = sampled_w.size(0)
        w_rep = sampled_w.repeat(obs.size(0), 1)  # second tiling of the reference (envelope.py:416)
        next_obs = obs.repeat_interleave(n_w, 0)  # (envelope.py:418)
        nq = self._q(self.q_net, next_obs, w_rep).view(obs.size(0), n_w, self.A, self.D)
        scal = th.einsum("br,bwar->bwa", w, nq)
        max_q, ac = th.max(scal, dim=2)
        pref = th.argmax(max_q, dim=1)
        nqt = self._q(self.target_q_net, next_obs, w_rep).view(obs.size(0), n_w, self.A, self.D)
        picked = nqt.gather(2, ac.unsqueeze(2).unsqueeze(3).expand(-1, -1, 1, self.D)).squeeze(2)
        return picked.gather(1, pref.reshape(-1, 1, 1).expand(-1, 1, self.D)).squeeze(1)

    @th.no_grad()
    def envelope_target_dedup(self, next_obs, sampled_w):
        """NOT the reference's code path: the same target (envelope.py:404-440) with Q evaluated once per DISTINCT (s'_b, w_j) pair
        (B*|W| rows instead of B*|W|^2) and a joint first-occurrence argmax over (j, a) -- the "de-duplicated CPU restatement" BASELINE.md
        section 2 quotes for context.  Returns rows in the reference's order k = i*B + b."""
        B, n_w = next_obs.size(0), sampled_w.size(0)
        w_rep = sampled_w.repeat(B, 1)
        nobs = next_obs.repeat_interleave(n_w, 0)
        nq = self._q(self.q_net, nobs, w_rep).view(B, n_w * self.A, self.D)
        nqt = self._q(self.target_q_net, nobs, w_rep).view(B, n_w * self.A, self.D)
        scal = th.einsum("id,bkd->ibk", sampled_w, nq)          # [W, B, W*A]
        best = th.argmax(scal, dim=2)                            # first occurrence over (j, a) row-major == two-stage max / argmax
        return nqt.gather(1, best.t().unsqueeze(2).expand(-1, -1, self.D)).transpose(0, 1).reshape(n_w * B, self.D)

    def update(self, obs, actions, rewards, next_obs, dones, sampled_w, homotopy_lambda=0.0, dedup=False):
        """One gradient step on a host minibatch; returns (loss, priorities of the first B rows)."""
        B, n_w = obs.size(0), sampled_w.size(0)
        w = sampled_w.repeat_interleave(B, 0)  # first tiling (envelope.py:284-291)
        obs_t, nobs_t = obs.repeat(n_w, 1), next_obs.repeat(n_w, 1)
        act_t, rew_t, done_t = actions.repeat(n_w, 1), rewards.repeat(n_w, 1), dones.repeat(n_w, 1)
        with th.no_grad():
            target = self.envelope_target_dedup(next_obs, sampled_w) if dedup else self.envelope_target(nobs_t, w, sampled_w)
            target_q = rew_t + (1 - done_t) * self.gamma * target
        q_values = self._q(self.q_net, obs_t, w)
        q_value = q_values.gather(1, act_t.long().reshape(-1, 1, 1).expand(-1, 1, self.D)).reshape(-1, self.D)
        loss = F.mse_loss(q_value, target_q)
        if homotopy_lambda > 0:
            aux = F.mse_loss(th.einsum("br,br->b", q_value, w), th.einsum("br,br->b", target_q, w))
            loss = (1 - homotopy_lambda) * loss + homotopy_lambda * aux
        self.optim.zero_grad()
        loss.backward()
        if self.max_grad_norm is not None:
            th.nn.utils.clip_grad_norm_(self.q_net.parameters(), self.max_grad_norm)
        self.optim.step()
        td = (q_value[:B] - target_q[:B]).detach()
        prio = th.einsum("sr,sr->s", td, w[:B]).abs()
        return float(loss.item()), prio.numpy()


from morl_baselines_b200.testing import synthetic_store  # noqa: E402,F401  (kept importable from here for the tests)
