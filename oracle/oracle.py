"""TEST INFRASTRUCTURE ONLY -- ctypes binding of oracle/morl_oracle.c (the CPU checker) over numpy arrays.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.
Every wrapper mirrors the signature of the matching C-ABI entry point in include/morl_b200.h, on HOST arrays.
"""

from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_build", "libmorl_oracle.so")

DOT_UNFUSED, DOT_FMA, DOT_PAIRFMA = 0, 1, 2
MAP_TILE, MAP_BLOCK = 0, 1
ROWS_REFERENCE, ROWS_BMAJOR = 0, 1
AC_ELEMENTWISE_MIN, AC_SCALAR_MIN, AC_ARGMIN_GATHER = 0, 1, 2

_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(HERE, "morl_oracle.c")
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        r = subprocess.run(["make", "-C", HERE, "-B" if force else "-s", "all"], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"oracle build failed:\n{r.stdout}\n{r.stderr}")
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB_PATH)
    return _lib


def _f32(x):
    return np.ascontiguousarray(x, dtype=np.float32)


def _p(a, t=C.c_float):
    return None if a is None else a.ctypes.data_as(C.POINTER(t))


def envelope_td(q_on, q_tg, wset, reward, done, gamma, mode=DOT_UNFUSED, row_order=ROWS_REFERENCE):
    q_on, q_tg, wset, reward, done = map(_f32, (q_on, q_tg, wset, reward, done))
    B, W, A, D = q_on.shape
    target = np.empty((W * B, D), np.float32)
    pref = np.empty(W * B, np.int32)
    act = np.empty(W * B, np.int32)
    lib().oracle_envelope_td(_p(q_on), _p(q_tg), _p(wset), _p(reward), _p(done.reshape(-1)), C.c_float(gamma), B, W, A, D, mode,
                             row_order, _p(target), _p(pref, C.c_int32), _p(act, C.c_int32))
    return target, pref, act


def greedy_td(q_sel, q_eval, w, reward, done, gamma, mode=DOT_UNFUSED, w_map=MAP_BLOCK, r_map=MAP_TILE):
    q_sel, q_eval, w = map(_f32, (q_sel, q_eval, w))
    N, A, D = q_sel.shape
    w = w.reshape(-1, D)
    out = np.empty((N, D), np.float32)
    act = np.empty(N, np.int32)
    if reward is not None:
        reward, done = _f32(reward).reshape(-1, D), _f32(done).reshape(-1)
        r_rows = reward.shape[0]
    else:
        r_rows = N
    lib().oracle_greedy_td(_p(q_sel), _p(q_eval), _p(w), w.shape[0], w_map, _p(reward), _p(done) if reward is not None else None,
                           r_rows, r_map, C.c_float(gamma), N, A, D, mode, _p(out), _p(act, C.c_int32))
    return out, act


def critic_min_td(q_nets, w, reward, done, gamma, mode=DOT_UNFUSED, w_map=MAP_BLOCK, r_map=MAP_TILE):
    q_nets, w = _f32(q_nets), _f32(w)
    n_nets, N, A, D = q_nets.shape
    w = w.reshape(-1, D)
    out = np.empty((N, D), np.float32)
    act = np.empty(N, np.int32)
    if reward is not None:
        reward, done = _f32(reward).reshape(-1, D), _f32(done).reshape(-1)
        r_rows = reward.shape[0]
    else:
        r_rows = N
    lib().oracle_critic_min_td(_p(q_nets), n_nets, _p(w), w.shape[0], w_map, _p(reward), _p(done) if reward is not None else None,
                               r_rows, r_map, C.c_float(gamma), N, A, D, mode, _p(out), _p(act, C.c_int32))
    return out, act


def gpi_envelope(q_nets, w, reward=None, done=None, gamma=0.0, mode=DOT_UNFUSED, w_map=MAP_BLOCK, r_map=MAP_TILE):
    q_nets, w = _f32(q_nets), _f32(w)
    n_nets, B, P, A, D = q_nets.shape
    w = w.reshape(-1, D)
    out = np.empty((B, D), np.float32)
    pol = np.empty(B, np.int32)
    act = np.empty(B, np.int32)
    if reward is not None:
        reward, done = _f32(reward).reshape(-1, D), _f32(done).reshape(-1)
        r_rows = reward.shape[0]
    else:
        r_rows = B
    lib().oracle_gpi_envelope(_p(q_nets), n_nets, _p(w), w.shape[0], w_map, _p(reward), _p(done) if reward is not None else None,
                              r_rows, r_map, C.c_float(gamma), B, P, A, D, mode, _p(out), _p(pol, C.c_int32), _p(act, C.c_int32))
    return out, pol, act


def actor_critic_td(q_nets, w, reward, done, logp, alpha, gamma, variant, w_map=MAP_BLOCK):
    q_nets = _f32(q_nets)
    n_nets, N, D = q_nets.shape
    reward, done = _f32(reward).reshape(N, D), _f32(done).reshape(-1)
    if w is not None:
        w = _f32(w).reshape(-1, D)
    if logp is not None:
        logp = _f32(logp).reshape(-1)
    out = np.empty((N,) if variant == AC_SCALAR_MIN else (N, D), np.float32)
    lib().oracle_actor_critic_td(_p(q_nets), n_nets, _p(w), 0 if w is None else w.shape[0], w_map, _p(reward), _p(done), _p(logp),
                                 C.c_float(alpha), C.c_float(gamma), N, D, variant, _p(out))
    return out


def td_mse(q_values, action, target_q, wset, lam, B, W, row_order=ROWS_REFERENCE, want_grad=True):
    q_values, target_q, wset = map(_f32, (q_values, target_q, wset))
    N, A, D = q_values.shape
    action = np.ascontiguousarray(action, dtype=np.int32).reshape(-1)
    loss = np.zeros(1, np.float32)
    grad = np.empty_like(q_values) if want_grad else None
    q_taken = np.empty((N, D), np.float32)
    prio = np.empty(B, np.float32)
    lib().oracle_td_mse(_p(q_values), _p(action, C.c_int32), _p(target_q), _p(wset), C.c_float(lam), B, W, A, D, row_order, _p(loss),
                        _p(grad), _p(q_taken), _p(prio))
    return float(loss[0]), grad, q_taken, prio


def td_huber(q_values, action, target_q, target_gpi, w, min_priority, p_rows, w_map=MAP_BLOCK, want_grad=True):
    q_values, target_q = _f32(q_values), _f32(target_q)
    n_nets, N, A, D = q_values.shape
    action = np.ascontiguousarray(action, dtype=np.int32).reshape(-1)
    if target_gpi is not None:
        target_gpi = _f32(target_gpi)
    w = _f32(w).reshape(-1, D)
    loss = np.zeros(1, np.float32)
    grad = np.empty_like(q_values) if want_grad else None
    prio = np.empty(p_rows, np.float32)
    lib().oracle_td_huber(_p(q_values), n_nets, _p(action, C.c_int32), action.shape[0], _p(target_q), _p(target_gpi), _p(w), w.shape[0],
                          w_map, C.c_float(min_priority), N, A, D, p_rows, _p(loss), _p(grad), _p(prio))
    return float(loss[0]), grad, prio


def pareto_mask(pts, remove_duplicates=True):
    pts = np.ascontiguousarray(pts)
    if pts.dtype not in (np.float32, np.float64):
        pts = pts.astype(np.float64)
    N, D = pts.shape
    keep = np.zeros(N, np.uint8)
    if pts.dtype == np.float32:
        lib().oracle_pareto_mask_f32(_p(pts), N, D, int(remove_duplicates), _p(keep, C.c_uint8))
    else:
        lib().oracle_pareto_mask_f64(_p(pts, C.c_double), N, D, int(remove_duplicates), _p(keep, C.c_uint8))
    return keep.astype(bool)


def sumtree_levels(max_size):
    n_levels = int(np.ceil(np.log2(max_size))) + 1
    return np.zeros(2**n_levels - 1, np.float64), n_levels


def sumtree_sample(levels, n_levels, query):
    query = np.ascontiguousarray(query, dtype=np.float64)
    idx = np.empty(query.shape[0], np.int64)
    lib().oracle_sumtree_sample(_p(levels, C.c_double), n_levels, _p(query, C.c_double), query.shape[0], _p(idx, C.c_int64))
    return idx


def sumtree_batch_set(levels, n_levels, idx, prio):
    idx = np.ascontiguousarray(idx, dtype=np.int64)
    prio = np.ascontiguousarray(prio, dtype=np.float64)
    lib().oracle_sumtree_batch_set(_p(levels, C.c_double), n_levels, _p(idx, C.c_int64), _p(prio, C.c_double), idx.shape[0])


def polyak(param, target, tau):
    param = _f32(param)
    assert target.dtype == np.float32 and target.flags.c_contiguous
    lib().oracle_polyak(_p(param), _p(target), C.c_int64(param.size), C.c_double(tau))
    return target
