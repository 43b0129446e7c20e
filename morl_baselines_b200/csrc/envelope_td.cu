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
#include <stdlib.h>

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


// =================================================================================================================
// v2 fast path: packed-FP32 (FMUL2 / FFMA2) scalarisation, 3-input FMNMX3 max tree, persistent CTAs with a
// register-prefetched double buffer.  Bit-identical to the scalar arithmetic above:
//   * mul.rn.f32x2 is two IEEE multiplies; the unfused add is issued as fma.rn.f32x2(acc, ONE, p) with ONE = 1.0f passed
//     as a RUNTIME kernel argument -- ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 even with -fmad=false
//     (observed with CUDA 12.9), but it cannot fold a multiplier it does not know; fl(acc * 1 + p) == fl(acc + p);
//   * the Q_on[b] block is transposed once into shared memory as SoA planes Qs[r][c] (c = j*A + a), so one LDS.128 yields
//     the r-th objective of four consecutive candidates already sitting in aligned register pairs;
//   * candidates are scanned in groups of 8: four packed dot products, max of 8 with FMNMX3, and only the GROUP index of
//     the running maximum is tracked (strict '>' keeps the first group); the exact (first) position inside the winning
//     group is recovered afterwards by re-evaluating its 8 scores with the scalar path and testing equality.
// =================================================================================================================
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk2(float lo, float hi) {
    u64 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void upk2(u64 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) {
    u64 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) {
    u64 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ float max3(float a, float b, float c) {
    float r;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}

// packed w . q for two candidates; `one2` = (1.0f, 1.0f) built from the runtime kernel argument
template <int D, int MODE>
__device__ __forceinline__ u64 dotw2(const u64 (&w2)[D], const u64 (&q2)[D], u64 one2) {
    if constexpr (MODE == MORL_DOT_UNFUSED) {
        u64 acc = mul2(w2[0], q2[0]);
#pragma unroll
        for (int r = 1; r < D; ++r) acc = fma2(acc, one2, mul2(w2[r], q2[r]));
        return acc;
    } else if constexpr (MODE == MORL_DOT_FMA) {
        u64 acc = mul2(w2[0], q2[0]);
#pragma unroll
        for (int r = 1; r < D; ++r) acc = fma2(w2[r], q2[r], acc);
        return acc;
    } else {
        u64 acc = 0;
#pragma unroll
        for (int r = 0; r + 1 < D; r += 2) {
            const u64 p = fma2(w2[r + 1], q2[r + 1], mul2(w2[r], q2[r]));
            acc = (r == 0) ? p : fma2(acc, one2, p);
        }
        if constexpr (D % 2 == 1) {
            const u64 t = mul2(w2[D - 1], q2[D - 1]);
            acc = (D == 1) ? t : fma2(acc, one2, t);
        }
        return acc;
    }
}

// ---- 1-D bulk async copy (TMA, UBLKCP) + mbarrier helpers -----------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src),
                 "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// v3 kernel.  On top of the packed arithmetic described above:
//   * FILTER (MODE != FMA): the scan uses the 3-op FMA chain t = fma(w2,q2, fma(w1,q1, w0*q0)) as an approximation of the
//     contract arithmetic e (5 roundings).  |t - e| <= 6u * sum_r |w_r q_r| < eps := 2^-21 * (sum_r |w_r|) * max|Q_on[b]|.
//     Per weight the scan keeps the best group maximum, its (first) group, and the runner-up group maximum.  If
//     runner_up < best - thr (thr = 2^-19 * ..., a 4x margin over 2 eps) every candidate outside the best group is exactly
//     smaller than the best group's maximum, so the exact first argmax lies in that group and is found by evaluating its 8
//     candidates in the contract arithmetic.  Otherwise (near tie, ~1e-4 of the rows on continuous data; always for constant
//     or non-finite Q) the row is re-scanned exactly by its whole warp.  The result is bit-identical to the exact scan.
//   * Q_tg[b] is staged with one 1-D bulk async copy (TMA) that overlaps the scan; reward/done are pre-loaded;
//   * the transposing Q_on load walks candidates (no div/mod), bank-conflict free.
template <int D, int MODE>
__global__ void __launch_bounds__(256) envelope_td_v3_kernel(const float* __restrict__ q_on, const float* __restrict__ q_tg,
                                                             const float* __restrict__ wset, const float* __restrict__ reward,
                                                             const float* __restrict__ done, float gamma, float one, int B, int W,
                                                             int A, int Cp, int CS, int gps, int row_order,
                                                             float* __restrict__ target_out, int32_t* __restrict__ pref_out,
                                                             int32_t* __restrict__ act_out) {
    constexpr bool FILTER = (MODE != MORL_DOT_FMA);
    constexpr int SCAN_MODE = MORL_DOT_FMA;  // arithmetic of the scan: the FMA chain (exact when MODE == FMA)
    extern __shared__ __align__(16) float smem[];
    const int C = W * A;
    const int plane = Cp;
    float* Qs = smem;                                        // [D][Cp] SoA planes of Q_on[b]
    float* Qt = smem + D * plane;                            // [C*D] AoS copy of Q_tg[b] (bulk async copy)
    const int WI = blockDim.x / CS;
    float* red_v = Qt + ((C * D + 3) & ~3);                  // [CS][WI] best
    float* red_s = red_v + CS * WI;                          // [CS][WI] runner-up
    int* red_g = reinterpret_cast<int*>(red_s + CS * WI);    // [CS][WI] group of best
    float* red_amax = reinterpret_cast<float*>(red_g + CS * WI);  // [32] per-warp max |Q|
    uint64_t* bar = reinterpret_cast<uint64_t*>(red_amax + 32);
    uint64_t* bar_on = bar + 1;
    float* Qa = reinterpret_cast<float*>(bar + 2);           // [C*D] AoS staging of Q_on[b] (bulk async copy)

    const int il = threadIdx.x % WI;
    const int cs = threadIdx.x / WI;
    const int i = blockIdx.y * WI + il;
    const bool active = i < W;
    const int lane = threadIdx.x & 31;
    const int nwarps = blockDim.x >> 5;
    const u64 one2 = pk2(one, one);

    float w[D];
    u64 w2[D];
    float wsum = 0.f;
#pragma unroll
    for (int r = 0; r < D; ++r) {
        w[r] = active ? __ldg(wset + (size_t)i * D + r) : 0.f;
        w2[r] = pk2(w[r], w[r]);
        wsum += fabsf(w[r]);
    }
    const int ngroups = (C + 7) / 8;
    const int g_begin = cs * gps;
    const int g_end = min(g_begin + gps, ngroups);
    const int g_full = min(g_end, C / 8);
    const uint32_t qt_bytes = (uint32_t)(C * D) * 4u;  // multiple of 16 (launcher)

    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        mbar_init(bar_on, 1);
    }
    for (int t = threadIdx.x; t < (Cp - C) * D; t += blockDim.x) Qs[(t / (Cp - C)) * plane + C + t % (Cp - C)] = 0.f;
    __syncthreads();

    uint32_t parity = 0;
    for (int b = blockIdx.x; b < B; b += gridDim.x, parity ^= 1u) {
        if (threadIdx.x == 0) {
            mbar_expect_tx(bar_on, qt_bytes);
            bulk_g2s(Qa, q_on + (size_t)b * C * D, qt_bytes, bar_on);
            mbar_expect_tx(bar, qt_bytes);
            bulk_g2s(Qt, q_tg + (size_t)b * C * D, qt_bytes, bar);
        }
        // pre-load the per-transition scalars of the epilogue
        float rw[D];
        float dn = 0.f;
        if (cs == 0) {
            dn = __ldg(done + b);
#pragma unroll
            for (int r = 0; r < D; ++r) rw[r] = __ldg(reward + (size_t)b * D + r);
        }
        // ---- Q_on[b] arrives by the same bulk async copy (AoS staging); transpose smem -> smem into the SoA planes ----
        // (reads at stride D words are bank-conflict free for odd D; the running max |q| feeds the filter threshold)
        float amax = 0.f;
        {
            mbar_wait(bar_on, parity);
            for (int c = threadIdx.x; c < C; c += blockDim.x) {
                float x[D];
#pragma unroll
                for (int r = 0; r < D; ++r) x[r] = Qa[c * D + r];
#pragma unroll
                for (int r = 0; r < D; ++r) {
                    Qs[r * plane + c] = x[r];
                    amax = fmaxf(amax, fabsf(x[r]));
                    if (!(x[r] == x[r])) amax = INFINITY;  // NaN in the block: force the exact path
                }
            }
            if (FILTER) {
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, off));
                if (lane == 0) red_amax[threadIdx.x >> 5] = amax;
            }
        }
        __syncthreads();
        float qmax = 0.f;
        if (FILTER)
            for (int k = 0; k < nwarps; ++k) qmax = fmaxf(qmax, red_amax[k]);

        // ---- scan: groups of 8 candidates, packed FMA-chain scores, best / runner-up group maxima ----
        float best = -INFINITY, second = -INFINITY;
        int bg = INT_MAX;
        auto scan_group = [&](int g, bool tail) {
            const int c0 = 8 * g;
            u64 q2[4][D];
#pragma unroll
            for (int r = 0; r < D; ++r) {
                const float4 lo = *reinterpret_cast<const float4*>(Qs + r * plane + c0);
                const float4 hi = *reinterpret_cast<const float4*>(Qs + r * plane + c0 + 4);
                q2[0][r] = pk2(lo.x, lo.y);
                q2[1][r] = pk2(lo.z, lo.w);
                q2[2][r] = pk2(hi.x, hi.y);
                q2[3][r] = pk2(hi.z, hi.w);
            }
            float x[8];
#pragma unroll
            for (int p = 0; p < 4; ++p) upk2(dotw2<D, SCAN_MODE>(w2, q2[p], one2), x[2 * p], x[2 * p + 1]);
            if (tail) {
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    if (c0 + k >= C) x[k] = -INFINITY;
            }
            const float m = fmaxf(max3(x[0], x[1], x[2]), max3(max3(x[3], x[4], x[5]), x[6], x[7]));
            if (FILTER) second = fmaxf(second, fminf(best, m));
            if (m > best) {
                best = m;
                bg = g;
            }
        };
#pragma unroll 2
        for (int g = g_begin; g < g_full; ++g) scan_group(g, false);
        for (int g = max(g_begin, g_full); g < g_end; ++g) scan_group(g, true);

        if (CS > 1) {
            red_v[cs * WI + il] = best;
            red_g[cs * WI + il] = bg;
            if (FILTER) red_s[cs * WI + il] = second;
            __syncthreads();
        }
        if (cs == 0) {  // warp-uniform: WI is a multiple of 32
            for (int s = 1; s < CS; ++s) {
                const float v2 = red_v[s * WI + il];
                const int g2 = red_g[s * WI + il];
                if (FILTER) second = fmaxf(fmaxf(second, red_s[s * WI + il]), fminf(best, v2));
                argmax_merge(best, bg, v2, g2);
            }
            int cstar = 0;
            bool amb = false;
            if (FILTER) {
                const float thr = 1.9073486328125e-06f * wsum * qmax;  // 2^-19 * sum|w| * max|Q|
                amb = active && !(second < best - thr);                // also true for NaN / inf
            }
            if (!amb && bg != INT_MAX) {
                // exact first argmax inside the winning group, contract arithmetic (strict '>' from the left)
                const int c0 = 8 * bg;
                float ev = -INFINITY;
                int kf = 0;
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    float q[D];
#pragma unroll
                    for (int r = 0; r < D; ++r) q[r] = Qs[r * plane + c0 + k];
                    const float s = dotw<D, MODE>(w, q);
                    if (c0 + k < C && s > ev) {
                        ev = s;
                        kf = k;
                    }
                }
                cstar = c0 + kf;
            }
            if (FILTER) {
                // near ties: the whole warp re-scans the row exactly (rare)
                unsigned ambmask = __ballot_sync(0xffffffffu, amb);
                while (ambmask) {
                    const int L = __ffs(ambmask) - 1;
                    ambmask &= ambmask - 1;
                    float wl[D];
#pragma unroll
                    for (int r = 0; r < D; ++r) wl[r] = __shfl_sync(0xffffffffu, w[r], L);
                    float bv = -INFINITY;
                    int bc = INT_MAX;
                    for (int c = lane; c < C; c += 32) {
                        float q[D];
#pragma unroll
                        for (int r = 0; r < D; ++r) q[r] = Qs[r * plane + c];
                        const float s = dotw<D, MODE>(wl, q);
                        if (s > bv) {
                            bv = s;
                            bc = c;
                        }
                    }
                    warp_argmax(bv, bc);
                    if (lane == L) cstar = (bc == INT_MAX) ? 0 : bc;
                }
            }
            mbar_wait(bar, parity);  // Q_tg[b] has landed in shared memory
            if (active) {
                const int jstar = cstar / A;
                const int astar = cstar - jstar * A;
                const float* qt = Qt + (size_t)cstar * D;
                const size_t k = (row_order == MORL_ROWS_REFERENCE) ? ((size_t)i * B + b) : ((size_t)b * W + i);
#pragma unroll
                for (int r = 0; r < D; ++r) target_out[k * D + r] = bellman(rw[r], dn, gamma, qt[r]);
                if (pref_out) pref_out[k] = jstar;
                if (act_out) act_out[k] = astar;
            }
        }
        __syncthreads();  // Qs / Qt / red_* are rewritten by the next transition
    }
}

// =================================================================================================================
// v5 kernel ("weight pairs"): the same FMA-chain filter + exact re-check as v3, re-blocked so that
//   * a thread owns TWO scalarising weights packed in f32x2 registers and the candidate value is the scalar operand of the packed
//     FMUL2 / FFMA2 (SASS form  FFMA2 Rd, Rw.F32x2, Rq.F32, Rc.F32x2): one LDS.128 of the AoS block feeds two weights, so the
//     shared-memory traffic per score halves and Q_on[b] is consumed exactly as the bulk copy (TMA) delivered it -- the
//     AoS -> SoA transposition pass of v3 (and its barrier) disappears;
//   * a warp covers all 64 weights of a weight block (lane = weight pair) and one quarter of the candidates, so every
//     shared-memory read is a warp-wide broadcast;
//   * groups of 16 candidates (book-keeping amortised over twice as many scores); the winner group is re-evaluated in the
//     contract arithmetic by two threads per weight.
// 2.9 instructions per score instead of 3.8.  Shapes: W*A a multiple of 16, 16-byte multiple Q blocks, |W| > 32 (below that half of
// the lanes would idle and v3 is used).  Bit-identical to v1 / v3 / the oracle (tests/test_kernels_gpu.py).
// =================================================================================================================
template <int D, int MODE>
__global__ void __launch_bounds__(128) envelope_td_wp_kernel(const float* __restrict__ q_on, const float* __restrict__ q_tg,
                                                             const float* __restrict__ wset, const float* __restrict__ reward,
                                                             const float* __restrict__ done, float gamma, int B, int W, int A,
                                                             int row_order, float* __restrict__ target_out,
                                                             int32_t* __restrict__ pref_out, int32_t* __restrict__ act_out, int pdl) {
    constexpr bool FILTER = (MODE != MORL_DOT_FMA);
    extern __shared__ __align__(16) float smem[];
    const int C = W * A;
    const int CDp = (C * D + 3) & ~3;
    if (pdl) {
        // programmatic dependent launch: this grid was allowed to become resident while its predecessor in the stream was still
        // draining; let OUR successor do the same, then wait until the predecessor's results are visible before any global read
        asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
        asm volatile("griddepcontrol.wait;" ::: "memory");
    }
    float* Qa = smem;                                         // [C*D] AoS Q_on[b]  (bulk async copy)
    float* Qt = Qa + CDp;                                     // [C*D] AoS Q_tg[b]  (bulk async copy)
    float* red_v = Qt + CDp;                                  // [4][64] best group maximum per (candidate quarter, weight)
    float* red_s = red_v + 256;                               // [4][64] runner-up
    int* red_g = reinterpret_cast<int*>(red_s + 256);         // [4][64] group of best
    unsigned* red_amax = reinterpret_cast<unsigned*>(red_g + 256);  // [4]
    uint64_t* bar_on = reinterpret_cast<uint64_t*>(red_amax + 4);
    uint64_t* bar_tg = bar_on + 1;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wbase = blockIdx.y * 64;
    // scan role: lane = weight pair, warp = candidate quarter
    const int i0 = wbase + 2 * lane, i1 = i0 + 1;
    u64 wp2[D];
#pragma unroll
    for (int r = 0; r < D; ++r) wp2[r] = pk2(i0 < W ? __ldg(wset + (size_t)i0 * D + r) : 0.f, i1 < W ? __ldg(wset + (size_t)i1 * D + r) : 0.f);
    // finish role: thread pair (2 k, 2 k + 1) finishes weight wbase + k
    const int fi = wbase + (tid >> 1), part = tid & 1;
    const bool f_active = fi < W;
    float fw[D];
    float wsum = 0.f;
#pragma unroll
    for (int r = 0; r < D; ++r) {
        fw[r] = f_active ? __ldg(wset + (size_t)fi * D + r) : 0.f;
        wsum += fabsf(fw[r]);
    }
    const int ngroups = C >> 4;
    const int gpw = (ngroups + 3) >> 2;
    const int g_begin = warp * gpw, g_end = min(g_begin + gpw, ngroups);
    const uint32_t q_bytes = (uint32_t)(C * D) * 4u;  // multiple of 16 (launcher)

    if (tid == 0) {
        mbar_init(bar_on, 1);
        mbar_init(bar_tg, 1);
    }
    __syncthreads();

    uint32_t parity = 0;
    for (int b = blockIdx.x; b < B; b += gridDim.x, parity ^= 1u) {
        if (tid == 0) {
            mbar_expect_tx(bar_on, q_bytes);
            bulk_g2s(Qa, q_on + (size_t)b * C * D, q_bytes, bar_on);
            mbar_expect_tx(bar_tg, q_bytes);
            bulk_g2s(Qt, q_tg + (size_t)b * C * D, q_bytes, bar_tg);
        }
        float rw[D];
        const float dn = __ldg(done + b);
#pragma unroll
        for (int r = 0; r < D; ++r) rw[r] = __ldg(reward + (size_t)b * D + r);
        mbar_wait(bar_on, parity);

        // ---- max |Q_on[b]| for the filter threshold (integer max of the magnitude bits: NaN / inf sort above every finite value) ----
        if (FILTER) {
            unsigned am = 0u;
            for (int t = tid; t < (C * D) >> 2; t += 128) {
                const uint4 x = *reinterpret_cast<const uint4*>(Qa + 4 * t);
                am = max(max(am, x.x & 0x7FFFFFFFu), max(x.y & 0x7FFFFFFFu, max(x.z & 0x7FFFFFFFu, x.w & 0x7FFFFFFFu)));
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) am = max(am, __shfl_xor_sync(0xffffffffu, am, off));
            if (lane == 0) red_amax[warp] = am;
        }

        // ---- scan: groups of 16 candidates, two weights per thread, FMA-chain scores ----
        float best0 = -INFINITY, second0 = -INFINITY, best1 = -INFINITY, second1 = -INFINITY;
        int bg0 = INT_MAX, bg1 = INT_MAX;
#pragma unroll 1
        for (int g = g_begin; g < g_end; ++g) {
            const float4* src = reinterpret_cast<const float4*>(Qa + (size_t)g * 16 * D);
            float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
            for (int sub = 0; sub < 4; ++sub) {
                float f[4 * D];
#pragma unroll
                for (int v = 0; v < D; ++v) {
                    const float4 x = src[sub * D + v];
                    f[4 * v + 0] = x.x;
                    f[4 * v + 1] = x.y;
                    f[4 * v + 2] = x.z;
                    f[4 * v + 3] = x.w;
                }
                float lo[4], hi[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    u64 acc = mul2(wp2[0], pk2(f[k * D], f[k * D]));
#pragma unroll
                    for (int r = 1; r < D; ++r) acc = fma2(wp2[r], pk2(f[k * D + r], f[k * D + r]), acc);
                    upk2(acc, lo[k], hi[k]);
                }
                m0 = max3(max3(m0, lo[0], lo[1]), lo[2], lo[3]);
                m1 = max3(max3(m1, hi[0], hi[1]), hi[2], hi[3]);
            }
            if (FILTER) {
                second0 = fmaxf(second0, fminf(best0, m0));
                second1 = fmaxf(second1, fminf(best1, m1));
            }
            if (m0 > best0) {
                best0 = m0;
                bg0 = g;
            }
            if (m1 > best1) {
                best1 = m1;
                bg1 = g;
            }
        }
        {
            const int k0 = warp * 64 + 2 * lane;
            *reinterpret_cast<float2*>(red_v + k0) = make_float2(best0, best1);
            *reinterpret_cast<int2*>(red_g + k0) = make_int2(bg0, bg1);
            if (FILTER) *reinterpret_cast<float2*>(red_s + k0) = make_float2(second0, second1);
        }
        __syncthreads();

        // ---- finish weight fi: merge the four quarters (candidate order), exact re-check of the winning group ----
        const int fl = tid >> 1;
        float bb = -INFINITY, ss = -INFINITY;
        int g = INT_MAX;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float pb = red_v[k * 64 + fl];
            if (FILTER) ss = fmaxf(fmaxf(ss, red_s[k * 64 + fl]), fminf(bb, pb));
            if (pb > bb) {
                bb = pb;
                g = red_g[k * 64 + fl];
            }
        }
        bool amb = false;
        if (FILTER) {
            const unsigned am = max(max(red_amax[0], red_amax[1]), max(red_amax[2], red_amax[3]));
            const float qmax = am >= 0x7F800000u ? INFINITY : __uint_as_float(am);
            const float thr = 1.9073486328125e-06f * wsum * qmax;  // 2^-19 * sum|w| * max|Q|
            amb = f_active && !(ss < bb - thr);                    // also true for NaN / inf
        }
        int cstar = 0;
        {
            const int gg = (g == INT_MAX) ? 0 : g;  // every candidate was -inf / NaN: th.argmax returns 0 (found by the re-check below)
            const int c0 = 16 * gg + 8 * part;
            float ev = -INFINITY;
            int ei = INT_MAX;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                float q[D];
#pragma unroll
                for (int r = 0; r < D; ++r) q[r] = Qa[(c0 + k) * D + r];
                const float s = dotw<D, MODE>(fw, q);
                if (s > ev) {
                    ev = s;
                    ei = c0 + k;
                }
            }
            const float ev2 = __shfl_xor_sync(0xffffffffu, ev, 1);
            const int ei2 = __shfl_xor_sync(0xffffffffu, ei, 1);
            argmax_merge(ev, ei, ev2, ei2);
            cstar = (ei == INT_MAX) ? 0 : ei;
        }
        if (FILTER) {
            // near ties: the whole warp re-scans the row exactly (rare)
            unsigned ambmask = __ballot_sync(0xffffffffu, amb && part == 0);
            while (ambmask) {
                const int L = __ffs(ambmask) - 1;
                ambmask &= ambmask - 1;
                float wl[D];
#pragma unroll
                for (int r = 0; r < D; ++r) wl[r] = __shfl_sync(0xffffffffu, fw[r], L);
                float bv = -INFINITY;
                int bc = INT_MAX;
                for (int c = lane; c < C; c += 32) {
                    float q[D];
#pragma unroll
                    for (int r = 0; r < D; ++r) q[r] = Qa[c * D + r];
                    const float s = dotw<D, MODE>(wl, q);
                    if (s > bv) {
                        bv = s;
                        bc = c;
                    }
                }
                warp_argmax(bv, bc);
                if ((lane & ~1) == L) cstar = (bc == INT_MAX) ? 0 : bc;
            }
        }
        mbar_wait(bar_tg, parity);  // Q_tg[b] has landed in shared memory
        if (f_active) {
            const size_t k = (row_order == MORL_ROWS_REFERENCE) ? ((size_t)fi * B + b) : ((size_t)b * W + fi);
            if (part == 0) {
                const float* qt = Qt + (size_t)cstar * D;
#pragma unroll
                for (int r = 0; r < D; ++r) target_out[k * D + r] = bellman(rw[r], dn, gamma, qt[r]);
            } else {
                const int jstar = cstar / A;
                if (pref_out) pref_out[k] = jstar;
                if (act_out) act_out[k] = cstar - jstar * A;
            }
        }
        __syncthreads();  // Qa / Qt / red_* are rewritten by the next transition
    }
}

struct EnvelopeV2Plan {
    bool ok;
    int Cp, CS, gps, WI;
    dim3 grid, block;
    size_t smem;
};

static EnvelopeV2Plan plan_envelope_v2(int B, int W, int A, int D, int sm_count) {
    EnvelopeV2Plan p{};
    (void)B;
    (void)sm_count;
    const long long C = (long long)W * A;
    p.ok = false;
    if ((C * D) % 4 != 0) return p;  // Q_tg[b] must be a 16-byte multiple for the bulk copy
    const int Wp = (W + 31) / 32 * 32;
    p.WI = Wp < 256 ? Wp : 256;
    const int warps_i = p.WI / 32;
    int cs = 4 / warps_i;
    if (cs < 1) cs = 1;
    const int ngroups = (int)((C + 7) / 8);
    if (cs > ngroups) cs = ngroups;
    p.CS = cs;
    p.gps = (ngroups + cs - 1) / cs;
    p.Cp = ngroups * 8;
    p.block = dim3((unsigned)(p.WI * p.CS), 1, 1);
    const size_t qs = (size_t)D * p.Cp * sizeof(float);
    const size_t qt = (((size_t)C * D + 3) & ~(size_t)3) * sizeof(float);
    p.smem = qs + 2 * qt + 3 * (size_t)p.CS * p.WI * sizeof(float) + 32 * sizeof(float) + 16;
    if (p.smem > 96 * 1024) return p;
    p.grid = dim3(1u, (unsigned)((W + p.WI - 1) / p.WI), 1);  // grid.x is set by the launcher from the measured occupancy
    p.ok = true;
    return p;
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

// Path selection, read on every call: MORL_ENVELOPE_PATH = "v1" (generic kernel), "v3" (CUDA-core fast path), "wp" (v5, weight-pair
// re-blocking of v3; the default whenever the shape fits, then v3, then v1).  (A tensor-core scoring path existed in round 1; it was
// bit-identical but measured SLOWER on B200 -- reading the 128 KB score tile back from tensor memory is limited to 64 B/clk per SM,
// profiles/r01_s3_envelope_tc_ncu.txt -- and has been removed from the library)  MORL_ENVELOPE_FORCE_V1 is the
// older spelling of "v1".
static int envelope_path_override() {
    if (getenv("MORL_ENVELOPE_FORCE_V1") != nullptr) return 1;
    const char* e = getenv("MORL_ENVELOPE_PATH");
    if (!e) return 0;
    if (e[0] == 'v' && e[1] == '1') return 1;
    if (e[0] == 'v' && e[1] == '3') return 3;
    if (e[0] == 'w' && e[1] == 'p') return 5;
    return 0;
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
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    static int sm_count_cached = 0;
    if (sm_count_cached == 0) {
        int dev = 0, n = 0;
        if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
            sm_count_cached = n;
        else {
            (void)cudaGetLastError();
            sm_count_cached = 148;
        }
    }
    const int path = envelope_path_override();
    // v5 (weight pairs): the default fast path when the shape fits; "wp" forces it, "v3" / "v1" skip it
    const long long Cw = (long long)W * A;
    const bool wp_ok = W > 32 && (Cw % 16) == 0 && ((Cw * D) % 4) == 0 && (2 * ((Cw * D + 3) & ~3LL) + 3 * 256 + 16) * 4 <= 96 * 1024;
    MORL_REQUIRE(path != 5 || wp_ok, MORL_ERR_UNSUPPORTED, "morl_envelope_td_f32: MORL_ENVELOPE_PATH=wp but the shape W=%d A=%d D=%d is outside the weight-pair path", W, A, D);
    if (wp_ok && (path == 0 || path == 5)) {
        bool launched5 = false;
        const size_t smem5 = (size_t)(2 * ((Cw * D + 3) & ~3LL) + 3 * 256 + 16) * 4;
        MORL_DISPATCH_D(D, MORL_DISPATCH_MODE(dot_mode, {
                            auto kern = envelope_td_wp_kernel<kD, kMode>;
                            if (smem5 > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem5);
                            int occ = 0;
                            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 128, smem5);
                            if (occ < 1) occ = 1;
                            const unsigned gy = (unsigned)((W + 63) / 64);
                            long long gx = (long long)sm_count_cached * occ / gy;
                            if (gx < 1) gx = 1;
                            if (gx > B) gx = B;
                            // measured (profiles/r01_s3_pdl_ab.txt): helps when one CTA per SM is resident (B = 148: 3.8 -> 3.35 us) and in
                            // launch-bound python loops, hurts at the north-star shape where the next grid's early CTAs compete with the
                            // 7 resident CTAs per SM of the running grid (9.85 -> 10.65 us); whole step unchanged -> opt-in only
                            static const bool want_pdl = [] { const char* e = getenv("MORL_ENVELOPE_PDL"); return e && e[0] == '1'; }();
                            if (want_pdl) {
                                // programmatic stream serialisation: the grid may start (barrier init, CTA residency) while the previous
                                // kernel of the stream drains; the kernel's griddepcontrol.wait restores the data dependency
                                cudaLaunchConfig_t cfg = {};
                                cfg.gridDim = dim3((unsigned)gx, gy, 1);
                                cfg.blockDim = dim3(128, 1, 1);
                                cfg.dynamicSmemBytes = smem5;
                                cfg.stream = st;
                                cudaLaunchAttribute attr[1];
                                attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
                                attr[0].val.programmaticStreamSerializationAllowed = 1;
                                cfg.attrs = attr;
                                cfg.numAttrs = 1;
                                cudaLaunchKernelEx(&cfg, kern, q_online, q_target, wset, reward, done, gamma, B, W, A, row_order, target_out, pref_out,
                                                   act_out, 1);
                            } else {
                                kern<<<dim3((unsigned)gx, gy, 1), 128, smem5, st>>>(q_online, q_target, wset, reward, done, gamma, B, W, A, row_order,
                                                                                  target_out, pref_out, act_out, 0);
                            }
                            launched5 = true;
                        }));
        MORL_REQUIRE(launched5, MORL_ERR_UNSUPPORTED, "morl_envelope_td_f32: no weight-pair kernel for D=%d mode=%d", D, dot_mode);
        return check_launch("morl_envelope_td_f32(wp)");
    }
    EnvelopeV2Plan p2 = plan_envelope_v2(B, W, A, D, sm_count_cached);
    if (p2.ok && path != 1) {
        bool launched2 = false;
        MORL_DISPATCH_D(D, MORL_DISPATCH_MODE(dot_mode, {
                            auto kern = envelope_td_v3_kernel<kD, kMode>;
                            if (p2.smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p2.smem);
                            int occ = 0;
                            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, (int)p2.block.x, p2.smem);
                            if (occ < 1) occ = 1;
                            // all transitions co-resident in one wave when they fit, else persistent CTAs striding over b
                            long long gx = (long long)sm_count_cached * occ / (long long)p2.grid.y;
                            if (gx < 1) gx = 1;
                            if (gx > B) gx = B;
                            p2.grid.x = (unsigned)gx;
                            kern<<<p2.grid, p2.block, p2.smem, st>>>(q_online, q_target, wset, reward, done, gamma, 1.0f, B, W, A, p2.Cp,
                                                                      p2.CS, p2.gps, row_order, target_out, pref_out, act_out);
                            launched2 = true;
                        }));
        MORL_REQUIRE(launched2, MORL_ERR_UNSUPPORTED, "morl_envelope_td_f32: no fast-path kernel for D=%d mode=%d", D, dot_mode);
        return check_launch("morl_envelope_td_f32(v3)");
    }
    const EnvelopePlan p = plan_envelope(B, W, A, D);
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
