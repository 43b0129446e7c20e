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

constexpr int kV2MaxPrefetch = 4;  // float4 per thread held in registers for the next tile

template <int D, int MODE, bool PIPE>
__global__ void __launch_bounds__(256) envelope_td_v2_kernel(const float* __restrict__ q_on, const float* __restrict__ q_tg,
                                                             const float* __restrict__ wset, const float* __restrict__ reward,
                                                             const float* __restrict__ done, float gamma, float one, int B, int W,
                                                             int A, int Cp, int CS, int gps, int row_order,
                                                             float* __restrict__ target_out, int32_t* __restrict__ pref_out,
                                                             int32_t* __restrict__ act_out) {
    extern __shared__ __align__(16) float smem[];
    const int C = W * A;               // candidates per transition
    const int plane = Cp;              // floats per objective plane (multiple of 8, 16-byte aligned rows)
    const int stage_floats = D * plane;
    float* stage0 = smem;
    float* stage1 = smem + stage_floats;
    const int WI = blockDim.x / CS;    // weights handled per CTA (multiple of 32)
    float* red_v = smem + 2 * stage_floats;                   // [CS][WI]
    int* red_g = reinterpret_cast<int*>(red_v + CS * WI);     // [CS][WI]

    const int il = threadIdx.x % WI;
    const int cs = threadIdx.x / WI;
    const int i = blockIdx.y * WI + il;
    const bool active = i < W;
    const u64 one2 = pk2(one, one);

    float w[D];
    u64 w2[D];
#pragma unroll
    for (int r = 0; r < D; ++r) {
        w[r] = active ? __ldg(wset + (size_t)i * D + r) : 0.f;
        w2[r] = pk2(w[r], w[r]);
    }

    const int n4 = (C * D) / 4;  // the launcher guarantees C*D % 4 == 0 and 16-byte aligned rows
    const int ngroups = (C + 7) / 8;
    const int g_begin = cs * gps;
    const int g_end = min(g_begin + gps, ngroups);

    // zero the padding columns [C, Cp) of both stages once (scores of padded candidates are masked, but must be finite)
    for (int t = threadIdx.x; t < (Cp - C) * D; t += blockDim.x) {
        const int r = t / (Cp - C), c = C + t % (Cp - C);
        stage0[r * plane + c] = 0.f;
        stage1[r * plane + c] = 0.f;
    }

    auto store_tile = [&](float* st, const float4& v, int e4) {
        const float x[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int e = 4 * e4 + k;
            const int c = e / D, r = e - c * D;
            st[r * plane + c] = x[k];
        }
    };

    // prologue: first tile straight to stage 0
    int b = blockIdx.x;
    if (b < B) {
        const float4* src = reinterpret_cast<const float4*>(q_on + (size_t)b * C * D);
        for (int t = threadIdx.x; t < n4; t += blockDim.x) store_tile(stage0, __ldg(src + t), t);
    }
    __syncthreads();

    const bool can_prefetch = PIPE && (n4 <= kV2MaxPrefetch * (int)blockDim.x);
    int it = 0;
    for (; b < B; b += gridDim.x, ++it) {
        float* cur = (it & 1) ? stage1 : stage0;
        float* nxt = (it & 1) ? stage0 : stage1;
        const int bn = b + gridDim.x;
        float4 pf[PIPE ? kV2MaxPrefetch : 1];
        if (PIPE && can_prefetch && bn < B) {
            const float4* src = reinterpret_cast<const float4*>(q_on + (size_t)bn * C * D);
#pragma unroll
            for (int k = 0; k < kV2MaxPrefetch; ++k) {
                const int t = threadIdx.x + k * blockDim.x;
                if (t < n4) pf[k] = __ldg(src + t);
            }
        }

        // ---- scan: groups of 8 candidates, packed arithmetic ----
        float best = -INFINITY;
        int bg = INT_MAX;
        auto scan_group = [&](int g, bool tail) {
            const int c0 = 8 * g;
            u64 q2[4][D];
#pragma unroll
            for (int r = 0; r < D; ++r) {
                const float4 lo = *reinterpret_cast<const float4*>(cur + r * plane + c0);
                const float4 hi = *reinterpret_cast<const float4*>(cur + r * plane + c0 + 4);
                q2[0][r] = pk2(lo.x, lo.y);
                q2[1][r] = pk2(lo.z, lo.w);
                q2[2][r] = pk2(hi.x, hi.y);
                q2[3][r] = pk2(hi.z, hi.w);
            }
            float x[8];
#pragma unroll
            for (int p = 0; p < 4; ++p) upk2(dotw2<D, MODE>(w2, q2[p], one2), x[2 * p], x[2 * p + 1]);
            if (tail) {  // last, partially filled group: padded candidates can never win
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    if (c0 + k >= C) x[k] = -INFINITY;
            }
            const float m = fmaxf(max3(x[0], x[1], x[2]), max3(max3(x[3], x[4], x[5]), x[6], x[7]));
            if (m > best) {
                best = m;
                bg = g;
            }
        };
        const int g_full = min(g_end, C / 8);  // groups entirely inside [0, C)
        for (int g = g_begin; g < g_full; ++g) scan_group(g, false);
        for (int g = max(g_begin, g_full); g < g_end; ++g) scan_group(g, true);

        if (CS > 1) {
            red_v[cs * WI + il] = best;
            red_g[cs * WI + il] = bg;
            __syncthreads();
        }
        if (cs == 0 && active) {
            for (int s = 1; s < CS; ++s) argmax_merge(best, bg, red_v[s * WI + il], red_g[s * WI + il]);
            int cstar = 0;
            if (bg != INT_MAX) {
                // exact first position inside the winning group: re-evaluate its scores with the scalar arithmetic
                const int c0 = 8 * bg;
                int kf = -1;
#pragma unroll
                for (int k = 7; k >= 0; --k) {
                    float q[D];
#pragma unroll
                    for (int r = 0; r < D; ++r) q[r] = cur[r * plane + c0 + k];
                    const float s = dotw<D, MODE>(w, q);
                    if (s == best && c0 + k < C) kf = k;
                }
                cstar = c0 + (kf < 0 ? 0 : kf);
            }
            const int jstar = cstar / A;
            const int astar = cstar - jstar * A;
            const float* qt = q_tg + (((size_t)b * W + jstar) * A + astar) * D;
            const size_t k = (row_order == MORL_ROWS_REFERENCE) ? ((size_t)i * B + b) : ((size_t)b * W + i);
            const float dn = __ldg(done + b);
#pragma unroll
            for (int r = 0; r < D; ++r)
                target_out[k * D + r] = bellman(__ldg(reward + (size_t)b * D + r), dn, gamma, __ldg(qt + r));
            if (pref_out) pref_out[k] = jstar;
            if (act_out) act_out[k] = astar;
        }

        // ---- publish the next tile ----
        if (bn < B) {
            if (PIPE && can_prefetch) {
#pragma unroll
                for (int k = 0; k < (PIPE ? kV2MaxPrefetch : 1); ++k) {
                    const int t = threadIdx.x + k * blockDim.x;
                    if (t < n4) store_tile(nxt, pf[k], t);
                }
            } else {
                const float4* src = reinterpret_cast<const float4*>(q_on + (size_t)bn * C * D);
                for (int t = threadIdx.x; t < n4; t += blockDim.x) store_tile(nxt, __ldg(src + t), t);
            }
        }
        __syncthreads();
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
    const long long C = (long long)W * A;
    p.ok = false;
    if ((C * D) % 4 != 0) return p;  // rows of Q_on[b] must stay 16-byte aligned for the float4 loads
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
    const size_t stage = (size_t)D * p.Cp * sizeof(float);
    p.smem = 2 * stage + 2 * (size_t)p.CS * p.WI * sizeof(float);
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

static const bool g_force_v1 = (getenv("MORL_ENVELOPE_FORCE_V1") != nullptr);

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
    EnvelopeV2Plan p2 = plan_envelope_v2(B, W, A, D, sm_count_cached);
    if (p2.ok && !g_force_v1) {
        bool launched2 = false;
        MORL_DISPATCH_D(D, MORL_DISPATCH_MODE(dot_mode, {
                            auto k1 = envelope_td_v2_kernel<kD, kMode, false>;  // one transition per CTA, everything co-resident
                            auto kp = envelope_td_v2_kernel<kD, kMode, true>;   // persistent, register-prefetched double buffer
                            if (p2.smem > 48 * 1024) {
                                cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p2.smem);
                                cudaFuncSetAttribute(kp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p2.smem);
                            }
                            int occ1 = 0, occp = 0;
                            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ1, k1, (int)p2.block.x, p2.smem);
                            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occp, kp, (int)p2.block.x, p2.smem);
                            const long long ctas_per_b = p2.grid.y;
                            if (occ1 > 0 && (long long)B * ctas_per_b <= (long long)sm_count_cached * occ1) {
                                p2.grid.x = (unsigned)B;  // a single wave: no tail, no pipeline needed
                                k1<<<p2.grid, p2.block, p2.smem, st>>>(q_online, q_target, wset, reward, done, gamma, 1.0f, B, W, A, p2.Cp,
                                                                        p2.CS, p2.gps, row_order, target_out, pref_out, act_out);
                            } else {
                                long long gx = (long long)sm_count_cached * (occp > 0 ? occp : 1) / ctas_per_b;
                                if (gx < 1) gx = 1;
                                if (gx > B) gx = B;
                                p2.grid.x = (unsigned)gx;
                                kp<<<p2.grid, p2.block, p2.smem, st>>>(q_online, q_target, wset, reward, done, gamma, 1.0f, B, W, A, p2.Cp,
                                                                        p2.CS, p2.gps, row_order, target_out, pref_out, act_out);
                            }
                            launched2 = true;
                        }));
        MORL_REQUIRE(launched2, MORL_ERR_UNSUPPORTED, "morl_envelope_td_f32: no v2 kernel for D=%d mode=%d", D, dot_mode);
        return check_launch("morl_envelope_td_f32(v2)");
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
