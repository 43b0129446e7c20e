// envelope_td.cu -- fused envelope-max TD target (SURVEY.md K3+K4).
//
// Replaces Envelope.envelope_target (reference multi_policy/envelope/envelope.py:404-440) and the vector Bellman
// line (envelope.py:298).  The reference runs both Q-nets on B*W^2 tiled rows; here the operator consumes the
// Q tensors of the B*W DISTINCT (s'_b, w_j) rows and performs, per output row (i, b),
//     (j*, a*) = first argmax_{j,a} wset[i] . Q_on[b, j, a, :]        (joint first-occurrence == th.max(dim=2) then th.argmax(dim=1))
//     out      = r[b] + ((1 - done[b]) * gamma) * Q_tg[b, j*, a*, :]
//
// Mapping (one CTA per transition b): the Q_on[b] block (W*A*D floats, 6 KB at the north-star shape) is staged in
// shared memory with 128-bit loads; thread (js, i) owns scalarising weight i and scans the j = js, js+JS, ...
// candidates, so every shared-memory read is a warp-wide broadcast (all lanes of a warp share js) and no shuffle
// is needed in the hot loop; the JS partial results per weight are merged through shared memory with the
// (value desc, flat index asc) order, which preserves first-occurrence semantics.  Q_tg is only touched at the
// winning (j*, a*) (a D-float gather per output row), never streamed.
#include <limits.h>

#include "common.cuh"

namespace morl {

template <int D, int MODE, bool VEC4>
__global__ void __launch_bounds__(512) envelope_td_kernel(const float* __restrict__ q_on, const float* __restrict__ q_tg,
                                                          const float* __restrict__ wset, const float* __restrict__ reward,
                                                          const float* __restrict__ done, float gamma, int B, int W, int A,
                                                          int Wc, int JS, int JT, int row_order,
                                                          float* __restrict__ target_out, int32_t* __restrict__ pref_out,
                                                          int32_t* __restrict__ act_out) {
    extern __shared__ __align__(16) float smem[];
    const int AD = A * D;
    const int tile_floats = (JT * AD + 3) & ~3;
    float* tile = smem;
    float* pv = smem + tile_floats;                         // [JS][Wc] partial best values
    int* pi = reinterpret_cast<int*>(pv + JS * Wc);         // [JS][Wc] partial best flat indices

    const int b = blockIdx.x;
    const int il = threadIdx.x % Wc;
    const int js = threadIdx.x / Wc;
    const int i = blockIdx.y * Wc + il;
    const bool active = i < W;

    float w[D];
#pragma unroll
    for (int r = 0; r < D; ++r) w[r] = active ? __ldg(wset + (size_t)i * D + r) : 0.f;

    float best = -INFINITY;
    int bidx = INT_MAX;

    const float* qb = q_on + (size_t)b * W * AD;
    for (int j0 = 0; j0 < W; j0 += JT) {
        const int jt = min(JT, W - j0);
        const int n = jt * AD;
        const float* src = qb + (size_t)j0 * AD;
        if (j0 > 0) __syncthreads();  // previous tile fully consumed
        if (((reinterpret_cast<uintptr_t>(src) & 15u) == 0) && (n % 4 == 0)) {
            const float4* s4 = reinterpret_cast<const float4*>(src);
            float4* d4 = reinterpret_cast<float4*>(tile);
            for (int t = threadIdx.x; t < n / 4; t += blockDim.x) d4[t] = __ldg(s4 + t);
        } else {
            for (int t = threadIdx.x; t < n; t += blockDim.x) tile[t] = __ldg(src + t);
        }
        __syncthreads();
        if (active) {
            for (int jj = js; jj < jt; jj += JS) {
                const float* qj = tile + jj * AD;
                const int base = (j0 + jj) * A;
                if constexpr (VEC4) {
                    for (int a4 = 0; a4 < A; a4 += 4) {
                        float f[4 * D];
                        const float4* p = reinterpret_cast<const float4*>(qj + a4 * D);
#pragma unroll
                        for (int v = 0; v < D; ++v) {
                            const float4 x = p[v];
                            f[4 * v + 0] = x.x;
                            f[4 * v + 1] = x.y;
                            f[4 * v + 2] = x.z;
                            f[4 * v + 3] = x.w;
                        }
#pragma unroll
                        for (int t = 0; t < 4; ++t) {
                            float q[D];
#pragma unroll
                            for (int r = 0; r < D; ++r) q[r] = f[t * D + r];
                            const float s = dotw<D, MODE>(w, q);
                            if (s > best) {
                                best = s;
                                bidx = base + a4 + t;
                            }
                        }
                    }
                } else {
                    for (int a = 0; a < A; ++a) {
                        float q[D];
#pragma unroll
                        for (int r = 0; r < D; ++r) q[r] = qj[a * D + r];
                        const float s = dotw<D, MODE>(w, q);
                        if (s > best) {
                            best = s;
                            bidx = base + a;
                        }
                    }
                }
            }
        }
    }

    if (JS > 1) {
        pv[js * Wc + il] = best;
        pi[js * Wc + il] = bidx;
        __syncthreads();
    }
    if (js == 0 && active) {
        for (int s = 1; s < JS; ++s) argmax_merge(best, bidx, pv[s * Wc + il], pi[s * Wc + il]);
        if (bidx == INT_MAX) bidx = 0;  // every candidate was -inf (or NaN): th.argmax returns 0
        const int jstar = bidx / A;
        const int astar = bidx - jstar * A;
        const float* qt = q_tg + (((size_t)b * W + jstar) * A + astar) * D;
        const size_t k = (row_order == MORL_ROWS_REFERENCE) ? ((size_t)i * B + b) : ((size_t)b * W + i);
        const float dn = __ldg(done + b);
#pragma unroll
        for (int r = 0; r < D; ++r)
            target_out[k * D + r] = bellman(__ldg(reward + (size_t)b * D + r), dn, gamma, __ldg(qt + r));
        if (pref_out) pref_out[k] = jstar;
        if (act_out) act_out[k] = astar;
    }
}

struct EnvelopePlan {
    int Wc, JS, JT;
    dim3 grid, block;
    size_t smem;
};

static EnvelopePlan plan_envelope(int B, int W, int A, int D) {
    EnvelopePlan p;
    const int Wp = (W + 31) / 32 * 32;
    p.Wc = Wp < 256 ? Wp : 256;
    int js = 256 / p.Wc;
    if (js > 8) js = 8;
    if (js > W) js = W;
    if (js < 1) js = 1;
    p.JS = js;
    const int AD = A * D;
    int jt = 10240 / AD;  // <= 40 KB of tile
    if (jt < 1) jt = 1;
    if (jt > W) jt = W;
    p.JT = jt;
    p.grid = dim3((unsigned)B, (unsigned)((W + p.Wc - 1) / p.Wc), 1);
    p.block = dim3((unsigned)(p.Wc * p.JS), 1, 1);
    const size_t tile_floats = ((size_t)jt * AD + 3) & ~(size_t)3;
    p.smem = (tile_floats + 2 * (size_t)p.JS * p.Wc) * sizeof(float);
    return p;
}

}  // namespace morl

extern "C" int morl_envelope_td_f32(const float* q_online, const float* q_target, const float* wset, const float* reward,
                                    const float* done, float gamma, int B, int W, int A, int D, int dot_mode,
                                    int row_order, float* target_out, int32_t* pref_out, int32_t* act_out, void* stream) {
    using namespace morl;
    MORL_REQUIRE(q_online && q_target && wset && reward && done && target_out, MORL_ERR_NULL,
                 "morl_envelope_td_f32: NULL pointer argument");
    MORL_REQUIRE(B > 0 && W > 0 && A > 0 && D > 0, MORL_ERR_SHAPE, "morl_envelope_td_f32: bad shape B=%d W=%d A=%d D=%d", B,
                 W, A, D);
    MORL_REQUIRE(D <= MORL_MAX_D && A * D <= 10240, MORL_ERR_UNSUPPORTED,
                 "morl_envelope_td_f32: unsupported D=%d (max %d) or A*D=%d (max 10240)", D, MORL_MAX_D, A * D);
    MORL_REQUIRE((long long)W * A < INT_MAX, MORL_ERR_UNSUPPORTED, "morl_envelope_td_f32: W*A overflows int32");
    MORL_REQUIRE(dot_mode >= 0 && dot_mode <= 2, MORL_ERR_UNSUPPORTED, "morl_envelope_td_f32: bad dot_mode %d", dot_mode);
    MORL_REQUIRE(row_order == MORL_ROWS_REFERENCE || row_order == MORL_ROWS_BMAJOR, MORL_ERR_UNSUPPORTED,
                 "morl_envelope_td_f32: bad row_order %d", row_order);
    MORL_REQUIRE(aligned16(q_online) && aligned16(q_target) && aligned16(target_out), MORL_ERR_ALIGN,
                 "morl_envelope_td_f32: q_online/q_target/target_out must be 16-byte aligned");
    const EnvelopePlan p = plan_envelope(B, W, A, D);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const bool vec4 = (A % 4 == 0);
    bool launched = false;
    MORL_DISPATCH_D(D, MORL_DISPATCH_MODE(dot_mode, {
                        if (vec4)
                            envelope_td_kernel<kD, kMode, true><<<p.grid, p.block, p.smem, st>>>(
                                q_online, q_target, wset, reward, done, gamma, B, W, A, p.Wc, p.JS, p.JT, row_order,
                                target_out, pref_out, act_out);
                        else
                            envelope_td_kernel<kD, kMode, false><<<p.grid, p.block, p.smem, st>>>(
                                q_online, q_target, wset, reward, done, gamma, B, W, A, p.Wc, p.JS, p.JT, row_order,
                                target_out, pref_out, act_out);
                        launched = true;
                    }));
    MORL_REQUIRE(launched, MORL_ERR_UNSUPPORTED, "morl_envelope_td_f32: no kernel for D=%d mode=%d", D, dot_mode);
    return check_launch("morl_envelope_td_f32");
}
