// gpi_td.cu -- per-row-weight TD targets: Double-DQN, GPI-PD critic-min, GPI envelope / policy-set evaluation,
// and the continuous-action (CAPQL / MOSAC / TD3-style GPI-PD) vector targets.  See include/morl_b200.h for the
// reference lines each entry point replaces.  These operators stream every Q element exactly once (HBM-bound).
#include <limits.h>

#include "common.cuh"

namespace morl {

// ---- Double-DQN: a* from q_select, value from q_eval (envelope.py:442-463, gpi_pd.py:648-656) ---------------
template <int D, int MODE>
__global__ void __launch_bounds__(256) greedy_td_kernel(const float* __restrict__ q_sel, const float* __restrict__ q_eval,
                                                        const float* __restrict__ w, int w_rows, int w_map,
                                                        const float* __restrict__ reward, const float* __restrict__ done,
                                                        int r_rows, int r_map, float gamma, int N, int A,
                                                        float* __restrict__ out, int32_t* __restrict__ act_out) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= N) return;
    float wv[D];
    const int wi = map_row(k, w_rows, N, w_map);
#pragma unroll
    for (int r = 0; r < D; ++r) wv[r] = __ldg(w + (size_t)wi * D + r);
    const float* qs = q_sel + (size_t)k * A * D;
    float best = -INFINITY;
    int ba = INT_MAX;
    for (int a = 0; a < A; ++a) {
        float q[D];
#pragma unroll
        for (int r = 0; r < D; ++r) q[r] = __ldg(qs + a * D + r);
        const float s = dotw<D, MODE>(wv, q);
        if (s > best) {
            best = s;
            ba = a;
        }
    }
    if (ba == INT_MAX) ba = 0;
    const float* qe = q_eval + ((size_t)k * A + ba) * D;
    if (reward) {
        const int ri = map_row(k, r_rows, N, r_map);
        const float dn = __ldg(done + ri);
#pragma unroll
        for (int r = 0; r < D; ++r) out[(size_t)k * D + r] = bellman(__ldg(reward + (size_t)ri * D + r), dn, gamma, __ldg(qe + r));
    } else {
#pragma unroll
        for (int r = 0; r < D; ++r) out[(size_t)k * D + r] = __ldg(qe + r);
    }
    if (act_out) act_out[k] = ba;
}

// ---- GPI-PD update target: scalarised first-argmin over critics per action, then greedy (gpi_pd.py:445-463) ---
template <int D, int MODE>
__global__ void __launch_bounds__(256) critic_min_td_kernel(const float* __restrict__ q_nets, int n_nets,
                                                            const float* __restrict__ w, int w_rows, int w_map,
                                                            const float* __restrict__ reward, const float* __restrict__ done,
                                                            int r_rows, int r_map, float gamma, int N, int A,
                                                            float* __restrict__ out, int32_t* __restrict__ act_out) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= N) return;
    float wv[D];
    const int wi = map_row(k, w_rows, N, w_map);
#pragma unroll
    for (int r = 0; r < D; ++r) wv[r] = __ldg(w + (size_t)wi * D + r);
    const size_t net_stride = (size_t)N * A * D;
    const float* q0 = q_nets + (size_t)k * A * D;
    float best = -INFINITY;
    int ba = INT_MAX, bn = 0;
    for (int a = 0; a < A; ++a) {
        float smin = 0.f;
        int nmin = 0;
        for (int n = 0; n < n_nets; ++n) {
            float q[D];
#pragma unroll
            for (int r = 0; r < D; ++r) q[r] = __ldg(q0 + n * net_stride + a * D + r);
            const float s = dotw<D, MODE>(wv, q);
            if (n == 0 || s < smin) {
                smin = s;
                nmin = n;
            }
        }
        if (a == 0) bn = nmin;  // fallback when no candidate beats -inf: th.argmax -> action 0
        if (smin > best) {
            best = smin;
            ba = a;
            bn = nmin;
        }
    }
    if (ba == INT_MAX) ba = 0;
    const float* qe = q0 + bn * net_stride + ba * D;
    if (reward) {
        const int ri = map_row(k, r_rows, N, r_map);
        const float dn = __ldg(done + ri);
#pragma unroll
        for (int r = 0; r < D; ++r) out[(size_t)k * D + r] = bellman(__ldg(reward + (size_t)ri * D + r), dn, gamma, __ldg(qe + r));
    } else {
#pragma unroll
        for (int r = 0; r < D; ++r) out[(size_t)k * D + r] = __ldg(qe + r);
    }
    if (act_out) act_out[k] = ba;
}

// ---- GPI envelope over a support set: one warp per row, lanes stride the P*A candidates (gpi_pd.py:662-690, 564-582)
template <int D, int MODE>
__global__ void __launch_bounds__(256) gpi_envelope_kernel(const float* __restrict__ q_nets, int n_nets,
                                                           const float* __restrict__ w, int w_rows, int w_map,
                                                           const float* __restrict__ reward, const float* __restrict__ done,
                                                           int r_rows, int r_map, float gamma, int B, int P, int A,
                                                           float* __restrict__ out, int32_t* __restrict__ policy_out,
                                                           int32_t* __restrict__ act_out) {
    const int lane = threadIdx.x & 31;
    const int b = (int)(((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (b >= B) return;  // warp-uniform
    float wv[D];
    const int wi = map_row(b, w_rows, B, w_map);
#pragma unroll
    for (int r = 0; r < D; ++r) wv[r] = __ldg(w + (size_t)wi * D + r);
    const int PA = P * A;
    const size_t net_stride = (size_t)B * PA * D;
    const float* qb = q_nets + (size_t)b * PA * D;

    auto critic_min = [&](int c, float& smin, int& nmin) {
        smin = 0.f;
        nmin = 0;
        for (int n = 0; n < n_nets; ++n) {
            float q[D];
#pragma unroll
            for (int r = 0; r < D; ++r) q[r] = __ldg(qb + n * net_stride + (size_t)c * D + r);
            const float s = dotw<D, MODE>(wv, q);
            if (n == 0 || s < smin) {
                smin = s;
                nmin = n;
            }
        }
    };

    float best = -INFINITY;
    int bidx = INT_MAX, bnet = 0;
    for (int c = lane; c < PA; c += 32) {
        float smin;
        int nmin;
        critic_min(c, smin, nmin);
        if (smin > best) {
            best = smin;
            bidx = c;
            bnet = nmin;
        }
    }
    // warp merge with (value desc, index asc); the winning lane's net index rides along
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const float v2 = __shfl_xor_sync(0xffffffffu, best, off);
        const int i2 = __shfl_xor_sync(0xffffffffu, bidx, off);
        const int n2 = __shfl_xor_sync(0xffffffffu, bnet, off);
        if (v2 > best || (v2 == best && i2 < bidx)) {
            best = v2;
            bidx = i2;
            bnet = n2;
        }
    }
    if (lane == 0) {
        if (bidx == INT_MAX) {  // all candidates -inf / NaN: th.argmax -> (0, 0)
            bidx = 0;
            float s;
            critic_min(0, s, bnet);
        }
        const float* qe = qb + bnet * net_stride + (size_t)bidx * D;
        if (reward) {
            const int ri = map_row(b, r_rows, B, r_map);
            const float dn = __ldg(done + ri);
#pragma unroll
            for (int r = 0; r < D; ++r)
                out[(size_t)b * D + r] = bellman(__ldg(reward + (size_t)ri * D + r), dn, gamma, __ldg(qe + r));
        } else if (out) {
#pragma unroll
            for (int r = 0; r < D; ++r) out[(size_t)b * D + r] = __ldg(qe + r);
        }
        if (policy_out) policy_out[b] = bidx / A;
        if (act_out) act_out[b] = bidx % A;
    }
}

// ---- continuous-action vector targets -----------------------------------------------
template <int D>
__global__ void __launch_bounds__(256) actor_critic_td_kernel(const float* __restrict__ q_nets, int n_nets,
                                                              const float* __restrict__ w, int w_rows, int w_map,
                                                              const float* __restrict__ reward, const float* __restrict__ done,
                                                              const float* __restrict__ logp, float alpha, float gamma, int N,
                                                              int variant, float* __restrict__ out) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= N) return;
    const size_t net_stride = (size_t)N * D;
    float wv[D];
    if (variant != MORL_AC_ELEMENTWISE_MIN) {
        const int wi = map_row(k, w_rows, N, w_map);
#pragma unroll
        for (int r = 0; r < D; ++r) wv[r] = __ldg(w + (size_t)wi * D + r);
    }
    const float dn = __ldg(done + k);
    const float ent = logp ? __fmul_rn(alpha, __ldg(logp + k)) : 0.f;
    if (variant == MORL_AC_ELEMENTWISE_MIN) {
        // CAPQL (capql.py:329-331): th.min over critics per objective, minus alpha*logp, vector Bellman
#pragma unroll
        for (int r = 0; r < D; ++r) {
            float m = __ldg(q_nets + (size_t)k * D + r);
            for (int n = 1; n < n_nets; ++n) m = fminf(m, __ldg(q_nets + n * net_stride + (size_t)k * D + r));
            const float soft = __fsub_rn(m, ent);
            out[(size_t)k * D + r] = bellman(__ldg(reward + (size_t)k * D + r), dn, gamma, soft);
        }
    } else if (variant == MORL_AC_SCALAR_MIN) {
        // MOSAC (mosac_continuous_action.py:438-442): th.matmul(q, w) per critic, min, - alpha*logp, scalar target
        float m = 0.f;
        for (int n = 0; n < n_nets; ++n) {
            float q[D];
#pragma unroll
            for (int r = 0; r < D; ++r) q[r] = __ldg(q_nets + n * net_stride + (size_t)k * D + r);
            const float s = dotw<D, MORL_DOT_UNFUSED>(wv, q);
            m = (n == 0) ? s : fminf(m, s);
        }
        float rv[D];
#pragma unroll
        for (int r = 0; r < D; ++r) rv[r] = __ldg(reward + (size_t)k * D + r);
        const float rs = dotw<D, MORL_DOT_UNFUSED>(wv, rv);
        out[k] = bellman(rs, dn, gamma, __fsub_rn(m, ent));
    } else {
        // TD3-style GPI-PD continuous (gpi_pd_continuous_action.py:397-403): first argmin_n of w.q_n, gather vector
        float smin = 0.f;
        int nmin = 0;
        for (int n = 0; n < n_nets; ++n) {
            float q[D];
#pragma unroll
            for (int r = 0; r < D; ++r) q[r] = __ldg(q_nets + n * net_stride + (size_t)k * D + r);
            const float s = dotw<D, MORL_DOT_UNFUSED>(wv, q);
            if (n == 0 || s < smin) {
                smin = s;
                nmin = n;
            }
        }
#pragma unroll
        for (int r = 0; r < D; ++r) {
            const float qv = __fsub_rn(__ldg(q_nets + nmin * net_stride + (size_t)k * D + r), ent);
            out[(size_t)k * D + r] = bellman(__ldg(reward + (size_t)k * D + r), dn, gamma, qv);
        }
    }
}

static int check_common(const char* fn, int N, int A, int D, int dot_mode, int w_rows, int r_rows, bool has_reward) {
    MORL_REQUIRE(N > 0 && A > 0 && D > 0, MORL_ERR_SHAPE, "%s: bad shape N=%d A=%d D=%d", fn, N, A, D);
    MORL_REQUIRE(D <= MORL_MAX_D, MORL_ERR_UNSUPPORTED, "%s: D=%d > %d", fn, D, MORL_MAX_D);
    MORL_REQUIRE(dot_mode >= 0 && dot_mode <= 2, MORL_ERR_UNSUPPORTED, "%s: bad dot_mode %d", fn, dot_mode);
    MORL_REQUIRE(w_rows > 0 && w_rows <= N && N % w_rows == 0, MORL_ERR_SHAPE, "%s: w_rows=%d must divide N=%d", fn, w_rows, N);
    if (has_reward)
        MORL_REQUIRE(r_rows > 0 && r_rows <= N && N % r_rows == 0, MORL_ERR_SHAPE, "%s: r_rows=%d must divide N=%d", fn, r_rows, N);
    return MORL_OK;
}

}  // namespace morl

extern "C" int morl_greedy_td_f32(const float* q_select, const float* q_eval, const float* w, int w_rows, int w_map,
                                  const float* reward, const float* done, int r_rows, int r_map, float gamma, int N, int A,
                                  int D, int dot_mode, float* target_out, int32_t* act_out, void* stream) {
    using namespace morl;
    MORL_REQUIRE(q_select && q_eval && w && target_out, MORL_ERR_NULL, "morl_greedy_td_f32: NULL pointer argument");
    MORL_REQUIRE(!reward || done, MORL_ERR_NULL, "morl_greedy_td_f32: reward given without done");
    int rc = check_common("morl_greedy_td_f32", N, A, D, dot_mode, w_rows, r_rows, reward != nullptr);
    if (rc) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int threads = 128;
    const int blocks = (N + threads - 1) / threads;
    MORL_DISPATCH_D(D, MORL_DISPATCH_MODE(dot_mode, (greedy_td_kernel<kD, kMode><<<blocks, threads, 0, st>>>(
                                                        q_select, q_eval, w, w_rows, w_map, reward, done, r_rows, r_map,
                                                        gamma, N, A, target_out, act_out))));
    return check_launch("morl_greedy_td_f32");
}

extern "C" int morl_critic_min_td_f32(const float* q_nets, int n_nets, const float* w, int w_rows, int w_map,
                                      const float* reward, const float* done, int r_rows, int r_map, float gamma, int N,
                                      int A, int D, int dot_mode, float* target_out, int32_t* act_out, void* stream) {
    using namespace morl;
    MORL_REQUIRE(q_nets && w && target_out, MORL_ERR_NULL, "morl_critic_min_td_f32: NULL pointer argument");
    MORL_REQUIRE(!reward || done, MORL_ERR_NULL, "morl_critic_min_td_f32: reward given without done");
    MORL_REQUIRE(n_nets > 0, MORL_ERR_SHAPE, "morl_critic_min_td_f32: n_nets=%d", n_nets);
    int rc = check_common("morl_critic_min_td_f32", N, A, D, dot_mode, w_rows, r_rows, reward != nullptr);
    if (rc) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int threads = 128;
    const int blocks = (N + threads - 1) / threads;
    MORL_DISPATCH_D(D, MORL_DISPATCH_MODE(dot_mode, (critic_min_td_kernel<kD, kMode><<<blocks, threads, 0, st>>>(
                                                        q_nets, n_nets, w, w_rows, w_map, reward, done, r_rows, r_map,
                                                        gamma, N, A, target_out, act_out))));
    return check_launch("morl_critic_min_td_f32");
}

extern "C" int morl_gpi_envelope_f32(const float* q_nets, int n_nets, const float* w, int w_rows, int w_map,
                                     const float* reward, const float* done, int r_rows, int r_map, float gamma, int B,
                                     int P, int A, int D, int dot_mode, float* out, int32_t* policy_out, int32_t* act_out,
                                     void* stream) {
    using namespace morl;
    MORL_REQUIRE(q_nets && w, MORL_ERR_NULL, "morl_gpi_envelope_f32: NULL pointer argument");
    MORL_REQUIRE(out || policy_out || act_out, MORL_ERR_NULL, "morl_gpi_envelope_f32: no output requested");
    MORL_REQUIRE(!reward || (done && out), MORL_ERR_NULL, "morl_gpi_envelope_f32: reward given without done/out");
    MORL_REQUIRE(n_nets > 0 && P > 0, MORL_ERR_SHAPE, "morl_gpi_envelope_f32: n_nets=%d P=%d", n_nets, P);
    MORL_REQUIRE((long long)P * A < INT_MAX, MORL_ERR_UNSUPPORTED, "morl_gpi_envelope_f32: P*A overflows int32");
    int rc = check_common("morl_gpi_envelope_f32", B, A, D, dot_mode, w_rows, r_rows, reward != nullptr);
    if (rc) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int threads = 128;  // 4 rows per CTA
    const long long total = (long long)B * 32;
    const int blocks = (int)((total + threads - 1) / threads);
    MORL_DISPATCH_D(D, MORL_DISPATCH_MODE(dot_mode, (gpi_envelope_kernel<kD, kMode><<<blocks, threads, 0, st>>>(
                                                        q_nets, n_nets, w, w_rows, w_map, reward, done, r_rows, r_map,
                                                        gamma, B, P, A, out, policy_out, act_out))));
    return check_launch("morl_gpi_envelope_f32");
}

extern "C" int morl_actor_critic_td_f32(const float* q_nets, int n_nets, const float* w, int w_rows, int w_map,
                                        const float* reward, const float* done, const float* logp, float alpha,
                                        float gamma, int N, int D, int variant, float* target_out, void* stream) {
    using namespace morl;
    MORL_REQUIRE(q_nets && reward && done && target_out, MORL_ERR_NULL, "morl_actor_critic_td_f32: NULL pointer argument");
    MORL_REQUIRE(variant >= 0 && variant <= 2, MORL_ERR_UNSUPPORTED, "morl_actor_critic_td_f32: bad variant %d", variant);
    MORL_REQUIRE(variant == MORL_AC_ELEMENTWISE_MIN || w, MORL_ERR_NULL, "morl_actor_critic_td_f32: w required for variant %d", variant);
    MORL_REQUIRE(n_nets > 0 && N > 0 && D > 0, MORL_ERR_SHAPE, "morl_actor_critic_td_f32: bad shape n_nets=%d N=%d D=%d", n_nets, N, D);
    MORL_REQUIRE(D <= MORL_MAX_D, MORL_ERR_UNSUPPORTED, "morl_actor_critic_td_f32: D=%d > %d", D, MORL_MAX_D);
    if (variant != MORL_AC_ELEMENTWISE_MIN)
        MORL_REQUIRE(w_rows > 0 && w_rows <= N && N % w_rows == 0, MORL_ERR_SHAPE,
                     "morl_actor_critic_td_f32: w_rows=%d must divide N=%d", w_rows, N);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int threads = 128;
    const int blocks = (N + threads - 1) / threads;
    MORL_DISPATCH_D(D, (actor_critic_td_kernel<kD><<<blocks, threads, 0, st>>>(q_nets, n_nets, w, w_rows, w_map, reward, done,
                                                                                logp, alpha, gamma, N, variant, target_out)));
    return check_launch("morl_actor_critic_td_f32");
}
