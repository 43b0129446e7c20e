// envelope_td_tc.cu -- envelope-max TD target with the SCORING on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a.
//
// Same operator and same bit-exact result as envelope_td.cu (reference multi_policy/envelope/envelope.py:404-440 + :298):
//     (j*, a*) = first argmax_{j,a} wset[i] . Q_on[b, j, a, :]     (contract arithmetic MODE, first occurrence)
//     out[i,b] = r[b] + ((1 - done[b]) * gamma) * Q_tg[b, j*, a*, :]
// The CUDA-core kernel spends ~90 % of its instructions on the W x (W*A) score matrix of a transition.  Here that matrix is a
// tensor-core product used as a FILTER, and only the winners are re-evaluated in the contract arithmetic:
//
//   * every fp32 value is split exactly into three bf16 terms x = x0 + x1 + x2; all nine cross terms w_x[r] * q_y[r] of the D <= 3
//     objectives are laid out ALONG K (slot 10x + 3y + r, 30 of 32 slots used), so ONE K = 32 bf16 MMA row-by-column product is the
//     full-precision score: every bf16 x bf16 product is exact in fp32 and the 27 terms are accumulated in fp32 in tensor memory;
//   * M = 128 is filled with a block-diagonal trick: A = [[W_split, 0], [0, W_split]] (128 x 64), B = [Q_split(cands 0..255) ;
//     Q_split(cands 256..511)] (64 x 256): TMEM lane m = 64 h + i holds the scores of weight i against candidates 256 h + [0,256);
//   * each worker thread owns one TMEM lane: it streams its 256 scores with tcgen05.ld, keeps the best group-of-16 maximum, its
//     group and the runner-up group maximum (8 FMNMX3 + 5 bookkeeping instructions per 16 scores);
//   * |score_tc - score_contract| <= eps := 2^-17 * sum|w| * max|Q_on[b]| (27-term fp32 accumulation incl. truncation, plus the 5
//     roundings of the contract arithmetic; thr = 2^-15 * ... is a 2x margin over 2 eps).  If runner_up < best - thr the exact first
//     argmax lies in the best group: its 16 candidates are evaluated in the contract arithmetic by two threads.  Otherwise (near tie,
//     ~1 % of the rows on continuous data; always for constant / non-finite Q) the row is re-scanned exactly by a whole warp.
//     The result is therefore bit-identical to the exact scan for every input.
//
// Pipeline (one persistent CTA per SM, 576 threads, roles linked by mbarrier rings so that four transitions are in flight per SM --
// every stage is a latency chain, and the TMEM read port (64 B/clk per quadrant) and the 4 MMAs both cost ~512 cycles a transition):
//   warp 0     : producer   -- 1-D bulk async copies (TMA) of Q_on[b] and Q_tg[b] into a 4-deep staging ring;
//   warp 1     : one thread issues the 4 tcgen05.mma (M=128, N=256, K=16) of a transition into one of two TMEM accumulator stages;
//   warps 2-5  : converters -- Q_on[b] -> split bf16 K-major operand tiles (64-byte swizzle), two operand stages;
//   warps 6-13 : scanners   -- tcgen05.ld of the 128 x 256 scores (two warps per TMEM quadrant), best / runner-up group maxima;
//   warps 14-17: finishers  -- merge, exact re-check in the contract arithmetic, near-tie re-scan, Bellman epilogue.
#include <cuda.h>
#include <cuda_bf16.h>
#include <limits.h>
#include <stdlib.h>

#include "common.cuh"

namespace morl {
namespace etc {

constexpr int kQStages = 4;     // Q_on / Q_tg staging ring (TMA prefetch depth)
constexpr int kPStages = 2;     // partial-maxima ring between scanners and finishers
constexpr int kConvWarps = 4, kScanWarps = 8, kFinWarps = 4;
constexpr int kWarpProducer = 0, kWarpMma = 1, kWarpConv0 = 2, kWarpScan0 = 6, kWarpFin0 = 14;  // (kWarpScan0 + k) % 4 == TMEM quadrant
constexpr int kThreads = 32 * (2 + kConvWarps + kScanWarps + kFinWarps);   // 576
constexpr int kMaxC = 512;
constexpr uint32_t kTileA = 8192;       // 128 rows x 64 B
constexpr uint32_t kTileB = 16384;      // 256 rows x 64 B
constexpr uint32_t kStageHalf = 6144;   // 512 candidates x 3 objectives x 4 B
constexpr uint32_t kOffAX = 0, kOffAY = kTileA, kOffB = 2 * kTileA;
constexpr uint32_t kOffStage = kOffB + 2 * 2 * kTileB;                  // two operand stages of (X, Y) tiles
constexpr uint32_t kOffMisc = kOffStage + kQStages * 2 * kStageHalf;
constexpr uint32_t kSmemBytes = kOffMisc + 8192 + 1024;                 // + misc + alignment slack

__device__ __forceinline__ uint32_t s_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mb_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mb_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mb_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s_u32(bar)) : "memory");
}
// bounded spin: a protocol bug becomes a trap (launch error) instead of a hung GPU
__device__ __forceinline__ void mb_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok = 0;
    for (uint32_t it = 0; it < (1u << 24); ++it) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.b32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(s_u32(bar)), "r"(parity)
            : "memory");
        if (ok) return;
    }
    __trap();
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s_u32(dst)), "l"(src),
                 "r"(bytes), "r"(s_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s_u32(bar)) : "memory");
}
__device__ __forceinline__ void mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tm_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, "
        "%20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
          "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
          "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]),
          "=r"(v[31])
        : "r"(taddr));
}
__device__ __forceinline__ void tm_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ float fmax3(float a, float b, float c) {
    float r;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}
// K-major canonical layout, 64-byte swizzle: rows of 64 B (32 bf16), 8-row groups of 512 B (same descriptor as gemm_bf16x3.cu)
__device__ __forceinline__ uint64_t desc_k_sw64(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(512 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)4 << 61;
    return d;
}

// exact three-way bf16 split of one fp32 value: x = x0 + x1 + x2, every term with <= 8 significant bits (a bf16 value whose fp32 bit
// pattern has a zero low half).  Two flavours, both exact for every finite x below the overflow guard of the filter (|x| < 1e37):
//   * truncation: mask the low 16 bits, subtract exactly, repeat (3 LOP + 2 FADD);
//   * Veltkamp:   p = x * (2^16 + 1), h = p - (p - x) rounds x to 8 bits, the residual is exact (8 FMA-pipe ops, no ALU-pipe op).
// The scanners saturate the ALU pipe with FMNMX3, so the FMA-pipe flavour is the default.  (The first version used F2F.BF16 round-to-
// nearest conversions, which sit on a slow pipe with long-scoreboard latency.)
#ifndef MORL_ETC_VELTKAMP
#define MORL_ETC_VELTKAMP 1
#endif
__device__ __forceinline__ void split3t(float x, uint32_t& h0, uint32_t& h1, uint32_t& h2) {
#if MORL_ETC_VELTKAMP
    const float C = 65537.0f;
    const float p0 = __fmul_rn(x, C);
    const float a0 = __fsub_rn(p0, __fsub_rn(p0, x));
    const float r1 = __fsub_rn(x, a0);
    const float p1 = __fmul_rn(r1, C);
    const float a1 = __fsub_rn(p1, __fsub_rn(p1, r1));
    const float r2 = __fsub_rn(r1, a1);
    h0 = __float_as_uint(a0);
    h1 = __float_as_uint(a1);
    h2 = __float_as_uint(r2);
#else
    h0 = __float_as_uint(x) & 0xFFFF0000u;
    const float r1 = __fsub_rn(x, __uint_as_float(h0));
    h1 = __float_as_uint(r1) & 0xFFFF0000u;
    const float r2 = __fsub_rn(r1, __uint_as_float(h1));
    h2 = __float_as_uint(r2) & 0xFFFF0000u;
#endif
}
// running max |.| over packed bf16 pairs, NaN-propagating: max.NaN.xorsign.abs keeps max(|a|, |b|) per half (sign = xor, ignored)
__device__ __forceinline__ uint32_t absmax_bf16x2(uint32_t acc, uint32_t v) {
    uint32_t r;
    asm("max.NaN.xorsign.abs.bf16x2 %0, %1, %2;" : "=r"(r) : "r"(acc), "r"(v));
    return r;
}
// (lo, hi) bf16 pair from the HIGH halves of two fp32 bit patterns: one PRMT
__device__ __forceinline__ uint32_t pack_hi(uint32_t lo, uint32_t hi) { return __byte_perm(lo, hi, 0x7632); }

// one 64-byte operand row = 16 words; 16-byte chunk q of row `row` lives at chunk (q ^ ((row >> 1) & 3))  (SWIZZLE_64B)
__device__ __forceinline__ void store_row_sw64(uint8_t* tile, int row, const uint32_t (&wd)[16]) {
    uint8_t* p = tile + row * 64;
    const int sw = (row >> 1) & 3;
#pragma unroll
    for (int q = 0; q < 4; ++q)
        *reinterpret_cast<uint4*>(p + ((q ^ sw) << 4)) = make_uint4(wd[4 * q], wd[4 * q + 1], wd[4 * q + 2], wd[4 * q + 3]);
}

// slot 10x + 3y + r of an operand row; `s[r][k]` = k-th split term of objective r.  WEIGHT side holds w_x[r] (term index = x),
// CANDIDATE side holds q_y[r] (term index = y); per x the 10 halfwords are (., ., ., ., ., ., ., ., ., 0).
__device__ __forceinline__ void build_row_w(const uint32_t (&s)[3][3], uint32_t (&wd)[16]) {
#pragma unroll
    for (int x = 0; x < 3; ++x) {
        const uint32_t a = s[0][x], b = s[1][x], c = s[2][x];
        wd[5 * x + 0] = pack_hi(a, b);
        wd[5 * x + 1] = pack_hi(c, a);
        wd[5 * x + 2] = pack_hi(b, c);
        wd[5 * x + 3] = pack_hi(a, b);
        wd[5 * x + 4] = c >> 16;
    }
    wd[15] = 0u;
}
__device__ __forceinline__ void build_row_q(const uint32_t (&s)[3][3], uint32_t (&wd)[16]) {
    wd[0] = pack_hi(s[0][0], s[1][0]);
    wd[1] = pack_hi(s[2][0], s[0][1]);
    wd[2] = pack_hi(s[1][1], s[2][1]);
    wd[3] = pack_hi(s[0][2], s[1][2]);
    wd[4] = s[2][2] >> 16;
#pragma unroll
    for (int q = 0; q < 5; ++q) {
        wd[5 + q] = wd[q];
        wd[10 + q] = wd[q];
    }
    wd[15] = 0u;
}

template <int D, int MODE>
__global__ void __launch_bounds__(kThreads, 1)
envelope_td_tc_kernel(const float* __restrict__ q_on, const float* __restrict__ q_tg, const float* __restrict__ wset,
                      const float* __restrict__ reward, const float* __restrict__ done, float gamma, int B, int W, int A, int row_order,
                      float* __restrict__ target_out, int32_t* __restrict__ pref_out, int32_t* __restrict__ act_out,
                      unsigned long long* __restrict__ stats) {
    const long long t_start = stats ? clock64() : 0;
    extern __shared__ __align__(16) uint8_t raw_smem[];
    // 1 KB alignment by pointer arithmetic on the __shared__ array (keeps the address space known to the compiler: LDS / STS)
    uint8_t* sm = raw_smem + ((1024u - (s_u32(raw_smem) & 1023u)) & 1023u);
    uint8_t* tAX = sm + kOffAX;
    uint8_t* tAY = sm + kOffAY;
    uint8_t* tB0 = sm + kOffB;       // operand ring: stage s at tB0 + s * 2 * kTileB  (X tile, then Y tile)
    uint8_t* stage0 = sm + kOffStage;
    float* wsm = reinterpret_cast<float*>(sm + kOffMisc);              // [64][3] fp32 weights (zero padded)
    float* part_best = wsm + 192;                                      // [kPStages][2 colhalf][128 rows]
    float* part_second = part_best + kPStages * 256;
    int* part_g = reinterpret_cast<int*>(part_second + kPStages * 256);
    float* amax_ring = reinterpret_cast<float*>(part_g + kPStages * 256);   // [kQStages][4]
    uint64_t* qfull = reinterpret_cast<uint64_t*>(amax_ring + kQStages * 4);
    uint64_t* qempty = qfull + kQStages;
    uint64_t* bfull = qempty + kQStages;
    uint64_t* bempty = bfull + 2;
    uint64_t* tfull = bempty + 2;
    uint64_t* tempty = tfull + 2;
    uint64_t* pfull = tempty + 2;
    uint64_t* pempty = pfull + kPStages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pempty + kPStages);

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int C = W * A;
    const uint32_t q_bytes = (uint32_t)(C * D) * 4u;  // multiple of 16 (launcher)
    const int n_local = (B - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

    if (tid == 0) {
        for (int s = 0; s < kQStages; ++s) {
            mb_init(&qfull[s], 1);
            mb_init(&qempty[s], kFinWarps);
        }
        for (int s = 0; s < 2; ++s) {
            mb_init(&bfull[s], kConvWarps);
            mb_init(&bempty[s], 1);
            mb_init(&tfull[s], 1);
            mb_init(&tempty[s], kScanWarps);
        }
        for (int s = 0; s < kPStages; ++s) {
            mb_init(&pfull[s], kScanWarps);
            mb_init(&pempty[s], kFinWarps);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    auto issue_load = [&](int it) {
        const int b = (int)blockIdx.x + it * (int)gridDim.x;
        const int sq = it % kQStages;
        mb_expect_tx(&qfull[sq], 2u * q_bytes);
        uint8_t* st = stage0 + sq * 2 * kStageHalf;
        bulk_g2s(st, q_on + (size_t)b * C * D, q_bytes, &qfull[sq]);
        bulk_g2s(st + kStageHalf, q_tg + (size_t)b * C * D, q_bytes, &qfull[sq]);
    };
    // the first Q blocks are requested before anything else is set up: their HBM latency overlaps the prologue
    if (warp == kWarpProducer && lane == 0) {
        const int n0 = n_local < kQStages ? n_local : kQStages;
        for (int it = 0; it < n0; ++it) issue_load(it);
    }
    if (warp == kWarpMma) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(s_u32(tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (warp >= kWarpConv0 && warp < kWarpConv0 + kConvWarps) {
        // rows ct of A_X and A_Y:  A_X[i] = split(w_i) for i < 64, A_Y[64 + i] = split(w_i), everything else zero
        const int ct = tid - kWarpConv0 * 32;
        const int i = ct & 63;
        float wv[3] = {0.f, 0.f, 0.f};
        if (i < W) {
#pragma unroll
            for (int r = 0; r < D; ++r) wv[r] = __ldg(wset + (size_t)i * D + r);
        }
        if (ct < 64) {
            wsm[3 * i + 0] = wv[0];
            wsm[3 * i + 1] = wv[1];
            wsm[3 * i + 2] = wv[2];
        }
        uint32_t s[3][3];
#pragma unroll
        for (int r = 0; r < 3; ++r) split3t(wv[r], s[r][0], s[r][1], s[r][2]);
        uint32_t wd[16], zero[16];
        build_row_w(s, wd);
#pragma unroll
        for (int k = 0; k < 16; ++k) zero[k] = 0u;
        store_row_sw64(tAX, ct, ct < 64 ? wd : zero);
        store_row_sw64(tAY, ct, ct < 64 ? zero : wd);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem_base = *tmem_slot;
    long long c_wait = 0, c_busy = 0;

    if (warp == kWarpProducer) {
        // ================= producer: keeps kQStages Q blocks in flight =================
        if (lane == 0) {
            for (int it = kQStages; it < n_local; ++it) {
                const int sq = it % kQStages;
                mb_wait(&qempty[sq], ((uint32_t)(it / kQStages) & 1u) ^ 1u);
                issue_load(it);
            }
        }
    } else if (warp == kWarpMma) {
        // ================= MMA issuer =================
        if (lane == 0) {
            // instruction descriptor: D = f32, A = B = bf16, both K-major, N = 256, M = 128
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            const uint32_t ax = s_u32(tAX), ay = s_u32(tAY);
            for (int it = 0; it < n_local; ++it) {
                const uint32_t sb = (uint32_t)it & 1u, ph = ((uint32_t)it >> 1) & 1u;
                const long long c0 = stats ? clock64() : 0;
                mb_wait(&bfull[sb], ph);
                mb_wait(&tempty[sb], ph ^ 1u);
                if (stats) c_wait += clock64() - c0;
                fence_after();
                const uint32_t bx = s_u32(tB0 + sb * 2 * kTileB), by = bx + kTileB;
                const uint32_t d = tmem_base + sb * 256u;
                mma_bf16(d, desc_k_sw64(ax), desc_k_sw64(bx), idesc, 0u);
                mma_bf16(d, desc_k_sw64(ax + 32), desc_k_sw64(bx + 32), idesc, 1u);
                mma_bf16(d, desc_k_sw64(ay), desc_k_sw64(by), idesc, 1u);
                mma_bf16(d, desc_k_sw64(ay + 32), desc_k_sw64(by + 32), idesc, 1u);
                mma_commit(&bempty[sb]);  // the operand stage may be rewritten
                mma_commit(&tfull[sb]);   // the scores are in tensor memory
            }
            if (stats) atomicAdd(&stats[2], (unsigned long long)c_wait);
        }
    } else if (warp < kWarpConv0 + kConvWarps) {
        // ================= converters: Q_on[b] (fp32 AoS) -> split bf16 operand rows =================
        const int ct = tid - kWarpConv0 * 32;
        const int cw = warp - kWarpConv0;
        for (int it = 0; it < n_local; ++it) {
            const int sq = it % kQStages;
            const uint32_t sb = (uint32_t)it & 1u;
            long long c0 = stats ? clock64() : 0;
            mb_wait(&qfull[sq], (uint32_t)(it / kQStages) & 1u);
            mb_wait(&bempty[sb], (((uint32_t)it >> 1) & 1u) ^ 1u);
            long long c1 = stats ? clock64() : 0;
            const float* Qa = reinterpret_cast<const float*>(stage0 + sq * 2 * kStageHalf);
            uint8_t* tBX = tB0 + sb * 2 * kTileB;
            uint32_t amax2 = 0u;
#pragma unroll
            for (int k = 0; k < kMaxC / (kConvWarps * 32); ++k) {
                const int c = ct + k * (kConvWarps * 32);
                float x[3] = {0.f, 0.f, 0.f};
                if (c < C) {
#pragma unroll
                    for (int r = 0; r < D; ++r) x[r] = Qa[c * D + r];
                }
                uint32_t s[3][3];
#pragma unroll
                for (int r = 0; r < 3; ++r) split3t(x[r], s[r][0], s[r][1], s[r][2]);
                uint32_t wd[16];
                build_row_q(s, wd);
                // words 0 and 1 hold the leading terms q0[0], q0[1], q0[2] (and q1[0], which is smaller than q0[0]):
                // max |leading term| >= max|q| / (1 + 2^-7); NaN propagates, inf / overflow show up as inf
                amax2 = absmax_bf16x2(absmax_bf16x2(amax2, wd[0]), wd[1]);
                store_row_sw64(tBX + (c >> 8) * kTileB, c & 255, wd);
            }
            const float am_lo = __uint_as_float((amax2 & 0x7FFFu) << 16), am_hi = __uint_as_float(amax2 & 0x7FFF0000u);
            float amax = fmaxf(am_lo, am_hi);
            if (!(am_lo == am_lo) || !(am_hi == am_hi)) amax = INFINITY;  // NaN in the block: force the exact path for this transition
            // a NaN / inf whose payload sits only in the low 16 bits, or a finite value the split overflows on, leaves inf / NaN terms
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, off));
            if (lane == 0) amax_ring[sq * 4 + cw] = amax;
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) mb_arrive(&bfull[sb]);
            if (stats) {
                c_wait += c1 - c0;
                c_busy += clock64() - c1;
            }
        }
        if (stats && ct == 0) {
            atomicAdd(&stats[0], (unsigned long long)c_wait);
            atomicAdd(&stats[1], (unsigned long long)c_busy);
        }
    } else if (warp < kWarpScan0 + kScanWarps) {
        // ================= scanners: warp w reads TMEM lanes [32 (w % 4), +32), two warps per quadrant split the 256 columns =========
        const int quad = warp & 3;
        const int colhalf = (warp - kWarpScan0) >> 2;
        const int row = quad * 32 + lane;
        const int h = quad >> 1;  // candidate half of this TMEM lane
        const int nvg = min(max(C - 256 * h - 128 * colhalf, 0), 128) >> 4;  // valid groups of 16 (C % 16 == 0), warp-uniform
        const int nch = (nvg + 1) >> 1;
        const int g0 = 16 * h + 8 * colhalf;
        for (int it = 0; it < n_local; ++it) {
            const uint32_t st = (uint32_t)it & 1u, ph = ((uint32_t)it >> 1) & 1u;
            const int sp = it % kPStages;
            long long c0 = stats ? clock64() : 0;
            mb_wait(&tfull[st], ph);
            long long c1 = stats ? clock64() : 0;
            fence_after();
            const uint32_t t_lane = tmem_base + ((uint32_t)(quad * 32) << 16) + st * 256u + (uint32_t)(colhalf * 128);
            float best = -INFINITY, second = -INFINITY;
            int bg = g0;
            auto group = [&](const uint32_t* v, int gidx) {
                const float t0 = fmax3(__uint_as_float(v[0]), __uint_as_float(v[1]), __uint_as_float(v[2]));
                const float t1 = fmax3(__uint_as_float(v[3]), __uint_as_float(v[4]), __uint_as_float(v[5]));
                const float t2 = fmax3(__uint_as_float(v[6]), __uint_as_float(v[7]), __uint_as_float(v[8]));
                const float t3 = fmax3(__uint_as_float(v[9]), __uint_as_float(v[10]), __uint_as_float(v[11]));
                const float t4 = fmax3(__uint_as_float(v[12]), __uint_as_float(v[13]), __uint_as_float(v[14]));
                const float m = fmaxf(fmax3(t0, t1, t2), fmax3(t3, t4, __uint_as_float(v[15])));
                second = fmaxf(second, fminf(best, m));
                if (m > best) {
                    best = m;
                    bg = gidx;
                }
            };
            {
                uint32_t va[32], vb[32];
                if (nch > 0) tm_ld32(t_lane, va);
                if (nch > 1) tm_ld32(t_lane + 32u, vb);
                if (nch > 0) {
                    tm_ld_wait();  // both loads (issued back to back) have landed
                    group(va, g0);
                    if (1 < nvg) group(va + 16, g0 + 1);
                }
                if (nch > 2) tm_ld32(t_lane + 64u, va);
                if (nch > 1) {
                    group(vb, g0 + 2);
                    if (3 < nvg) group(vb + 16, g0 + 3);
                }
                if (nch > 3) tm_ld32(t_lane + 96u, vb);
                if (nch > 2) {
                    tm_ld_wait();
                    group(va, g0 + 4);
                    if (5 < nvg) group(va + 16, g0 + 5);
                }
                if (nch > 3) {
                    group(vb, g0 + 6);
                    if (7 < nvg) group(vb + 16, g0 + 7);
                }
            }
            fence_before();  // this thread's TMEM loads are ordered before the MMA that reuses the accumulator stage
            mb_wait(&pempty[sp], (((uint32_t)(it / kPStages)) & 1u) ^ 1u);
            const int pi = (sp * 2 + colhalf) * 128 + row;
            part_best[pi] = best;
            part_second[pi] = second;
            part_g[pi] = bg;
            __syncwarp();
            if (lane == 0) {
                mb_arrive(&tempty[st]);
                mb_arrive(&pfull[sp]);
            }
            if (stats) {
                c_wait += c1 - c0;
                c_busy += clock64() - c1;
            }
        }
        if (stats && warp == kWarpScan0 && lane == 0) {
            atomicAdd(&stats[3], (unsigned long long)c_wait);
            atomicAdd(&stats[4], (unsigned long long)c_busy);
        }
    } else {
        // ================= finishers: merge the partial maxima, exact re-check of the winning group (or exact re-scan), epilogue ======
        const int ft = tid - kWarpFin0 * 32;
        const int fi = ft >> 1, part = ft & 1;  // thread pair (2 i, 2 i + 1) finishes weight i
        const bool f_active = fi < W;
        float fw[D];
        float wsum = 0.f;
#pragma unroll
        for (int r = 0; r < D; ++r) {
            fw[r] = wsm[3 * fi + r];
            wsum += fabsf(fw[r]);
        }
        for (int it = 0; it < n_local; ++it) {
            const int b = (int)blockIdx.x + it * (int)gridDim.x;
            const int sq = it % kQStages;
            const int sp = it % kPStages;
            const float* Qa = reinterpret_cast<const float*>(stage0 + sq * 2 * kStageHalf);
            const float* Qt = reinterpret_cast<const float*>(stage0 + sq * 2 * kStageHalf + kStageHalf);
            float rw[D];
            const float dn = __ldg(done + b);
#pragma unroll
            for (int r = 0; r < D; ++r) rw[r] = __ldg(reward + (size_t)b * D + r);
            long long c0 = stats ? clock64() : 0;
            mb_wait(&qfull[sq], (uint32_t)(it / kQStages) & 1u);  // completed long ago: makes the TMA-written block visible here
            mb_wait(&pfull[sp], (uint32_t)(it / kPStages) & 1u);
            long long c1 = stats ? clock64() : 0;
            float qmax = 0.f;
#pragma unroll
            for (int k = 0; k < 4; ++k) qmax = fmaxf(qmax, amax_ring[sq * 4 + k]);
            qmax *= 1.0078125f;  // amax is taken on the leading bf16 terms: max|q| <= amax * (1 + 2^-7)
            // partials in candidate order: (h = 0, cols 0..127), (h = 0, cols 128..255), (h = 1, ...), (h = 1, ...)
            float bb = -INFINITY, ss = -INFINITY;
            int g = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int pi = (sp * 2 + (k & 1)) * 128 + 64 * (k >> 1) + fi;
                const float pb = part_best[pi];
                ss = fmaxf(fmaxf(ss, part_second[pi]), fminf(bb, pb));
                if (pb > bb) {
                    bb = pb;
                    g = part_g[pi];
                }
            }
            const float scale = wsum * qmax;
            const float thr = 3.0517578125e-05f * scale;  // 2^-15 * sum|w| * max|Q|
            // ambiguous (exact re-scan): near ties, NaN / inf / all-equal rows, and magnitudes where bf16 products could
            // underflow or overflow (the filter's error bound assumes normal fp32 arithmetic)
            const bool amb = f_active && (!(ss < bb - thr) || !(scale > 1.0e-30f && scale < 1.0e37f));
            int cstar = 0;
            {
                // both threads of the pair evaluate 8 candidates of the winning group (harmless when the row is ambiguous)
                const int c0g = 16 * g + 8 * part;
                float ev = -INFINITY;
                int ei = INT_MAX;
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    float q[D];
#pragma unroll
                    for (int r = 0; r < D; ++r) q[r] = Qa[(c0g + k) * D + r];
                    const float s = dotw<D, MODE>(fw, q);
                    if (s > ev) {
                        ev = s;
                        ei = c0g + k;
                    }
                }
                const float ev2 = __shfl_xor_sync(0xffffffffu, ev, 1);
                const int ei2 = __shfl_xor_sync(0xffffffffu, ei, 1);
                argmax_merge(ev, ei, ev2, ei2);
                cstar = (ei == INT_MAX) ? 16 * g : ei;
            }
            {
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
            __syncwarp();
            if (lane == 0) {
                mb_arrive(&pempty[sp]);
                mb_arrive(&qempty[sq]);  // the staging slot (Q_on / Q_tg of this transition) may be refilled
            }
            if (stats) {
                c_wait += c1 - c0;
                c_busy += clock64() - c1;
            }
        }
        if (stats && ft == 0) {
            atomicAdd(&stats[5], (unsigned long long)c_wait);
            atomicAdd(&stats[6], (unsigned long long)c_busy);
            atomicAdd(&stats[7], (unsigned long long)(clock64() - t_start));
        }
    }

    fence_before();
    __syncthreads();
    if (warp == kWarpMma) {
        fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
    }
}

__device__ unsigned long long g_stats[8];

}  // namespace etc

// Returns 1 when the tensor-core path handled the call (kernel enqueued), 0 when the shape is outside its envelope
// (W <= 64, W*A <= 512 and a multiple of 16, D <= 3, 16-byte-multiple Q blocks); the caller then uses the CUDA-core kernels.
int envelope_td_tc_try_launch(const float* q_online, const float* q_target, const float* wset, const float* reward, const float* done,
                              float gamma, int B, int W, int A, int D, int dot_mode, int row_order, float* target_out, int32_t* pref_out,
                              int32_t* act_out, cudaStream_t st, int sm_count) {
    const long long C = (long long)W * A;
    if (W > 64 || C > etc::kMaxC || (C % 16) != 0 || D > 3 || ((C * D) % 4) != 0) return 0;
    bool launched = false;
    static const bool want_stats = [] { const char* e = getenv("MORL_ENVELOPE_STATS"); return e && e[0] == '1'; }();
    unsigned long long* stats = nullptr;
    if (want_stats) {
        void* sp = nullptr;
        cudaGetSymbolAddress(&sp, etc::g_stats);
        stats = static_cast<unsigned long long*>(sp);
    }
    MORL_DISPATCH_D(D, MORL_DISPATCH_MODE(dot_mode, {
                        if constexpr (kD <= 3) {
                            auto kern = etc::envelope_td_tc_kernel<kD, kMode>;
                            static bool configured = false;
                            if (!configured) {
                                cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)etc::kSmemBytes);
                                cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
                                configured = true;
                            }
                            int grid = sm_count;
                            if (grid > B) grid = B;
                            kern<<<grid, etc::kThreads, etc::kSmemBytes, st>>>(q_online, q_target, wset, reward, done, gamma, B, W, A, row_order,
                                                                              target_out, pref_out, act_out, stats);
                            launched = true;
                        }
                    }));
    return launched ? 1 : 0;
}

}  // namespace morl

// Diagnostics: per-phase cycle counters of worker thread 0, summed over CTAs and launches since the last reset (collected only when
// MORL_ENVELOPE_STATS=1 was set before the first call).
extern "C" int morl_debug_envelope_stats(unsigned long long* out8, int reset) {
    using namespace morl;
    MORL_REQUIRE(out8, MORL_ERR_NULL, "morl_debug_envelope_stats: NULL pointer argument");
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out8, etc::g_stats, 8 * sizeof(unsigned long long));
    if (reset) {
        unsigned long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        cudaMemcpyToSymbol(etc::g_stats, z, sizeof(z));
    }
    return check_launch("morl_debug_envelope_stats");
}
