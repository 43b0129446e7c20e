// td_loss.cu -- fused TD loss + gradient seed + PER priority.
//
// Envelope (reference multi_policy/envelope/envelope.py:301-313, 329-331): gather Q(s, a_taken), MSE against the
// target, optional homotopy auxiliary loss on the scalarised values, d loss / d q_values, |w . td| priorities of the
// rows that carry weight index 0 -- one pass over q_values / target_q instead of ~12 eager kernels.
// GPI-PD (reference multi_policy/gpi_pd/gpi_pd.py:469-487, 507-520): Huber-style loss per critic and
// | w . max_n |delta_n| | priorities.
//
// The reduction is deterministic: fixed-shape block partials (float) + a single-block final sum in double.
#include "common.cuh"

namespace morl {

constexpr int kTdThreads = 256;

__device__ __forceinline__ float block_sum(float v, float* red) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    float t = 0.f;
    if (warp == 0) {
        t = (lane < (blockDim.x >> 5)) ? red[lane] : 0.f;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) t += __shfl_xor_sync(0xffffffffu, t, off);
    }
    return t;  // valid in thread 0
}

template <int D>
__device__ __forceinline__ void write_grad_row(float* __restrict__ grow, int A, int a_taken, const float (&g)[D]) {
    // dense row of A*D floats: zero except the taken action
    const int AD = A * D;
    if ((AD % 4 == 0) && ((reinterpret_cast<uintptr_t>(grow) & 15u) == 0)) {
        float4* g4 = reinterpret_cast<float4*>(grow);
        const int lo = a_taken * D, hi = lo + D;
        for (int v = 0; v < AD / 4; ++v) {
            float x[4];
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const int c = 4 * v + t;
                float val = 0.f;
                if (c >= lo && c < hi) {
#pragma unroll
                    for (int r = 0; r < D; ++r)
                        if (c - lo == r) val = g[r];
                }
                x[t] = val;
            }
            g4[v] = make_float4(x[0], x[1], x[2], x[3]);
        }
    } else {
        for (int a = 0; a < A; ++a)
#pragma unroll
            for (int r = 0; r < D; ++r) grow[a * D + r] = (a == a_taken) ? g[r] : 0.f;
    }
}

template <int D>
__global__ void __launch_bounds__(kTdThreads) td_mse_kernel(const float* __restrict__ q_values, const int32_t* __restrict__ action,
                                                            const float* __restrict__ target_q, const float* __restrict__ wset,
                                                            float lambda_arg, const float* __restrict__ lambda_dev, int B, int W, int A,
                                                            int row_order, float* __restrict__ grad_q, float* __restrict__ q_taken,
                                                            float* __restrict__ prio_out, float* __restrict__ partials) {
    pdl_enter();
    __shared__ float red[kTdThreads / 32];
    const float lambda = lambda_dev ? __ldg(lambda_dev) : lambda_arg;  // device-resident schedule value: a captured graph stays valid while it decays
    const long long N = (long long)B * W;
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    float sq = 0.f, aux2 = 0.f;
    if (k < N) {
        int i, b;
        if (row_order == MORL_ROWS_REFERENCE) {
            i = (int)(k / B);
            b = (int)(k - (long long)i * B);
        } else {
            b = (int)(k / W);
            i = (int)(k - (long long)b * W);
        }
        const int a = __ldg(action + b);
        float q[D], t[D], d[D], w[D];
#pragma unroll
        for (int r = 0; r < D; ++r) {
            q[r] = __ldg(q_values + ((size_t)k * A + a) * D + r);
            t[r] = __ldg(target_q + (size_t)k * D + r);
            w[r] = __ldg(wset + (size_t)i * D + r);
            d[r] = __fsub_rn(q[r], t[r]);
            sq += d[r] * d[r];
        }
        float aux = 0.f;
        if (lambda > 0.f) {
            // th.einsum("br,br->b", q_value, w) and (target_q, w) separately, then the difference (envelope.py:310-312)
            aux = __fsub_rn(dotw<D, MORL_DOT_UNFUSED>(q, w), dotw<D, MORL_DOT_UNFUSED>(t, w));
            aux2 = aux * aux;
        }
        if (q_taken) {
#pragma unroll
            for (int r = 0; r < D; ++r) q_taken[(size_t)k * D + r] = q[r];
        }
        if (grad_q) {
            const float c1 = (1.0f - lambda) * 2.0f / (float)((double)N * D);
            const float c2 = lambda * 2.0f / (float)N;
            float g[D];
#pragma unroll
            for (int r = 0; r < D; ++r) g[r] = c1 * d[r] + c2 * aux * w[r];
            write_grad_row<D>(grad_q + (size_t)k * A * D, A, a, g);
        }
        if (prio_out && i == 0) prio_out[b] = fabsf(dotw<D, MORL_DOT_UNFUSED>(d, w));  // envelope.py:330-331
    }
    const float s1 = block_sum(sq, red);
    const float s2 = block_sum(aux2, red);
    if (threadIdx.x == 0) {
        partials[2 * blockIdx.x + 0] = s1;
        partials[2 * blockIdx.x + 1] = s2;
    }
}

__global__ void __launch_bounds__(kTdThreads) td_mse_finalize_kernel(const float* __restrict__ partials, int n_blocks, float lambda_arg,
                                                                     const float* __restrict__ lambda_dev, double inv_nd, double inv_n,
                                                                     float* __restrict__ loss_out) {
    pdl_enter();
    __shared__ double red[2][kTdThreads];
    const float lambda = lambda_dev ? __ldg(lambda_dev) : lambda_arg;
    double a = 0.0, b = 0.0;
    for (int t = threadIdx.x; t < n_blocks; t += blockDim.x) {
        a += (double)partials[2 * t + 0];
        b += (double)partials[2 * t + 1];
    }
    red[0][threadIdx.x] = a;
    red[1][threadIdx.x] = b;
    __syncthreads();
    for (int s = blockDim.x / 2; s > 0; s >>= 1) {
        if (threadIdx.x < s) {
            red[0][threadIdx.x] += red[0][threadIdx.x + s];
            red[1][threadIdx.x] += red[1][threadIdx.x + s];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const double mse = red[0][0] * inv_nd;
        const double auxl = red[1][0] * inv_n;
        const double l = (lambda > 0.f) ? ((1.0 - (double)lambda) * mse + (double)lambda * auxl) : mse;
        loss_out[0] = (float)l;
    }
}

// ---- GPI-PD Huber-style loss ----------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(kTdThreads) td_huber_kernel(const float* __restrict__ q_values, int n_nets,
                                                              const int32_t* __restrict__ action, int a_rows,
                                                              const float* __restrict__ target_q, const float* __restrict__ target_gpi,
                                                              const float* __restrict__ w, int w_rows, int w_map, float min_priority,
                                                              int N, int A, int p_rows, float* __restrict__ grad_q,
                                                              float* __restrict__ prio_out, float* __restrict__ partials) {
    __shared__ float red[kTdThreads / 32];
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    float lsum = 0.f;
    if (k < N) {
        const int a = __ldg(action + (k % a_rows));
        float t[D], tg[D], emax[D];
#pragma unroll
        for (int r = 0; r < D; ++r) {
            t[r] = __ldg(target_q + (size_t)k * D + r);
            tg[r] = target_gpi ? __ldg(target_gpi + (size_t)k * D + r) : 0.f;
            emax[r] = 0.f;
        }
        const size_t net_stride = (size_t)N * A * D;
        const float gscale = 1.0f / (float)((double)N * D) / (float)n_nets;
        for (int n = 0; n < n_nets; ++n) {
            float g[D];
#pragma unroll
            for (int r = 0; r < D; ++r) {
                const float q = __ldg(q_values + n * net_stride + ((size_t)k * A + a) * D + r);
                const float d = __fsub_rn(q, t[r]);
                const float x = fabsf(d);
                // huber(x) = where(x < mp, 0.5 x^2, mp x)   (common/networks.py:90-100)
                lsum += (x < min_priority) ? 0.5f * x * x : min_priority * x;
                g[r] = ((x < min_priority) ? d : copysignf(min_priority, d) * (d != 0.f ? 1.f : 0.f)) * gscale;
                const float e = target_gpi ? fabsf(__fsub_rn(q, tg[r])) : x;
                emax[r] = (n == 0) ? e : fmaxf(emax[r], e);  // th.max over the stacked |errors| (gpi_pd.py:509, 516)
            }
            if (grad_q) write_grad_row<D>(grad_q + n * net_stride + (size_t)k * A * D, A, a, g);
        }
        if (prio_out && k < p_rows) {
            float wv[D];
            const int wi = map_row(k, w_rows, N, w_map);
#pragma unroll
            for (int r = 0; r < D; ++r) wv[r] = __ldg(w + (size_t)wi * D + r);
            prio_out[k] = fabsf(dotw<D, MORL_DOT_UNFUSED>(wv, emax));  // einsum("br,br->b", w, err).abs()
        }
    }
    const float s = block_sum(lsum, red);
    if (threadIdx.x == 0) partials[blockIdx.x] = s;
}

__global__ void __launch_bounds__(kTdThreads) td_huber_finalize_kernel(const float* __restrict__ partials, int n_blocks, double scale,
                                                                       float* __restrict__ loss_out) {
    __shared__ double red[kTdThreads];
    double a = 0.0;
    for (int t = threadIdx.x; t < n_blocks; t += blockDim.x) a += (double)partials[t];
    red[threadIdx.x] = a;
    __syncthreads();
    for (int s = blockDim.x / 2; s > 0; s >>= 1) {
        if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) loss_out[0] = (float)(red[0] * scale);
}

}  // namespace morl

extern "C" size_t morl_td_workspace_bytes(int n_rows) {
    if (n_rows <= 0) return 0;
    const size_t blocks = ((size_t)n_rows + morl::kTdThreads - 1) / morl::kTdThreads;
    return blocks * 2 * sizeof(float) + 16;
}

extern "C" int morl_td_mse_priority_f32(const float* q_values, const int32_t* action, const float* target_q, const float* wset,
                                        float homotopy_lambda, const float* homotopy_lambda_dev, int B, int W, int A, int D, int row_order,
                                        float* loss_out, float* grad_q, float* q_taken, float* prio_out, void* workspace, void* stream) {
    using namespace morl;
    MORL_REQUIRE(q_values && action && target_q && wset && loss_out && workspace, MORL_ERR_NULL,
                 "morl_td_mse_priority_f32: NULL pointer argument");
    MORL_REQUIRE(B > 0 && W > 0 && A > 0 && D > 0, MORL_ERR_SHAPE, "morl_td_mse_priority_f32: bad shape B=%d W=%d A=%d D=%d", B, W, A, D);
    MORL_REQUIRE(D <= MORL_MAX_D, MORL_ERR_UNSUPPORTED, "morl_td_mse_priority_f32: D=%d > %d", D, MORL_MAX_D);
    MORL_REQUIRE((long long)B * W < (1ll << 31), MORL_ERR_UNSUPPORTED, "morl_td_mse_priority_f32: B*W overflows int32");
    MORL_REQUIRE(row_order == MORL_ROWS_REFERENCE || row_order == MORL_ROWS_BMAJOR, MORL_ERR_UNSUPPORTED,
                 "morl_td_mse_priority_f32: bad row_order %d", row_order);
    MORL_REQUIRE(homotopy_lambda >= 0.f && homotopy_lambda <= 1.f, MORL_ERR_SHAPE, "morl_td_mse_priority_f32: lambda=%f outside [0,1]",
                 (double)homotopy_lambda);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const long long N = (long long)B * W;
    const int blocks = (int)((N + kTdThreads - 1) / kTdThreads);
    float* partials = static_cast<float*>(workspace);
    MORL_DISPATCH_D(D, (launch_k(td_mse_kernel<kD>, dim3(blocks), dim3(kTdThreads), 0, st, q_values, action, target_q, wset, homotopy_lambda,
                                                                         homotopy_lambda_dev, B, W, A, row_order, grad_q, q_taken, prio_out,
                                                                         partials)));
    int rc = check_launch("morl_td_mse_priority_f32");
    if (rc) return rc;
    launch_k(td_mse_finalize_kernel, dim3(1), dim3(kTdThreads), 0, st, partials, blocks, homotopy_lambda, homotopy_lambda_dev, 1.0 / ((double)N * D),
                                                     1.0 / (double)N, loss_out);
    return check_launch("morl_td_mse_priority_f32(finalize)");
}

extern "C" int morl_td_huber_priority_f32(const float* q_values, int n_nets, const int32_t* action, int a_rows, const float* target_q,
                                          const float* target_q_gpi, const float* w, int w_rows, int w_map, float min_priority, int N,
                                          int A, int D, int p_rows, float* loss_out, float* grad_q, float* prio_out, void* workspace,
                                          void* stream) {
    using namespace morl;
    MORL_REQUIRE(q_values && action && target_q && loss_out && workspace, MORL_ERR_NULL, "morl_td_huber_priority_f32: NULL pointer argument");
    MORL_REQUIRE(!prio_out || w, MORL_ERR_NULL, "morl_td_huber_priority_f32: prio_out requires w");
    MORL_REQUIRE(n_nets > 0 && N > 0 && A > 0 && D > 0 && a_rows > 0, MORL_ERR_SHAPE,
                 "morl_td_huber_priority_f32: bad shape n_nets=%d N=%d A=%d D=%d a_rows=%d", n_nets, N, A, D, a_rows);
    MORL_REQUIRE(D <= MORL_MAX_D, MORL_ERR_UNSUPPORTED, "morl_td_huber_priority_f32: D=%d > %d", D, MORL_MAX_D);
    MORL_REQUIRE(N % a_rows == 0, MORL_ERR_SHAPE, "morl_td_huber_priority_f32: a_rows=%d must divide N=%d", a_rows, N);
    MORL_REQUIRE(p_rows >= 0 && p_rows <= N, MORL_ERR_SHAPE, "morl_td_huber_priority_f32: p_rows=%d outside [0,N]", p_rows);
    if (prio_out) MORL_REQUIRE(w_rows > 0 && w_rows <= N && N % w_rows == 0, MORL_ERR_SHAPE,
                               "morl_td_huber_priority_f32: w_rows=%d must divide N=%d", w_rows, N);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int blocks = (N + kTdThreads - 1) / kTdThreads;
    float* partials = static_cast<float*>(workspace);
    MORL_DISPATCH_D(D, (td_huber_kernel<kD><<<blocks, kTdThreads, 0, st>>>(q_values, n_nets, action, a_rows, target_q, target_q_gpi, w,
                                                                           w_rows, w_map, min_priority, N, A, p_rows, grad_q, prio_out,
                                                                           partials)));
    int rc = check_launch("morl_td_huber_priority_f32");
    if (rc) return rc;
    td_huber_finalize_kernel<<<1, kTdThreads, 0, st>>>(partials, blocks, 1.0 / ((double)N * D) / (double)n_nets, loss_out);
    return check_launch("morl_td_huber_priority_f32(finalize)");
}
