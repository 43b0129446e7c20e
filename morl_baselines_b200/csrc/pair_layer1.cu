// pair_layer1.cu -- the separable first layer of the weight-conditioned Q-network on the pair batch: both input products in one small
// launch instead of two library sgemms and their epilogue kernels.
//
// Reference: QNet.forward builds [s || w] rows and runs nn.Linear on B*W (reference: B*W^2) of them
// (multi_policy/envelope/envelope.py:59-77).  Here  W1 [s || w] + b1 = W1_s s + (W1_w w + b1)  (DESIGN.md section 2), so only
//     u[b, :] = W1_s feats[b]                 (B rows, K = F)
//     v[j, :] = W1_w wset[j] + b1            (W rows, K = D)
// are computed (morl_pair_layer1_uv_f32: 6.3 us; the library path was two SIMT sgemms plus their epilogue kernels, ~15 us).  The
// parameter gradients dW1 = [dU^T feats | dV^T wset], db1 = colsum(dV) stay on the library path: a hand-written chunked reduction was
// measured at 56 us against ~30 us for the two library GEMMs and was dropped.
// fp32 FMA chains in a fixed order: deterministic, and within the 1e-5 parity bar of the dense layers (tests/test_gemm_gpu.py).
#include "common.cuh"

namespace morl {

constexpr int kL1Rows = 8;  // rows (transitions or weight vectors) per block of the forward kernel

// grid.x = ceil((B + W) / kL1Rows), block = 256 threads striding the H outputs.  W1 is row-major [H, F + D].
__global__ void __launch_bounds__(256) pair_layer1_uv_kernel(const float* __restrict__ feats, const float* __restrict__ wset, const float* __restrict__ W1,
                                                             const float* __restrict__ b1, int B, int W, int F, int D, int H, float* __restrict__ u,
                                                             float* __restrict__ v) {
    extern __shared__ float xs[];  // [kL1Rows][max(F, D)] input rows of this block
    const int K = F + D;
    const int r0 = blockIdx.x * kL1Rows;
    const int kmax = F > D ? F : D;
    for (int t = threadIdx.x; t < kL1Rows * kmax; t += blockDim.x) {
        const int rr = t / kmax, k = t - rr * kmax;
        const int r = r0 + rr;
        float x = 0.f;
        if (r < B) {
            if (k < F) x = __ldg(feats + (size_t)r * F + k);
        } else if (r < B + W) {
            if (k < D) x = __ldg(wset + (size_t)(r - B) * D + k);
        }
        xs[t] = x;
    }
    __syncthreads();
    for (int h = threadIdx.x; h < H; h += blockDim.x) {
        const float* wrow = W1 + (size_t)h * K;
        float accu[kL1Rows];
#pragma unroll
        for (int rr = 0; rr < kL1Rows; ++rr) accu[rr] = 0.f;
        // rows of this block are either all transitions, all weight vectors, or (one block at most) mixed: handle per row
        const bool any_u = r0 < B, any_v = r0 + kL1Rows > B;
        if (any_u) {
            for (int k = 0; k < F; ++k) {
                const float wk = __ldg(wrow + k);
#pragma unroll
                for (int rr = 0; rr < kL1Rows; ++rr)
                    if (r0 + rr < B) accu[rr] = __fmaf_rn(xs[rr * kmax + k], wk, accu[rr]);
            }
        }
        if (any_v) {
            const float bias = __ldg(b1 + h);
#pragma unroll
            for (int rr = 0; rr < kL1Rows; ++rr)
                if (r0 + rr >= B) accu[rr] = bias;
            for (int k = 0; k < D; ++k) {
                const float wk = __ldg(wrow + F + k);
#pragma unroll
                for (int rr = 0; rr < kL1Rows; ++rr)
                    if (r0 + rr >= B) accu[rr] = __fmaf_rn(xs[rr * kmax + k], wk, accu[rr]);
            }
        }
#pragma unroll
        for (int rr = 0; rr < kL1Rows; ++rr) {
            const int r = r0 + rr;
            if (r < B)
                u[(size_t)r * H + h] = accu[rr];
            else if (r < B + W)
                v[(size_t)(r - B) * H + h] = accu[rr];
        }
    }
}

}  // namespace morl

extern "C" int morl_pair_layer1_uv_f32(const float* feats, const float* wset, const float* W1, const float* b1, int B, int W, int F, int D, int H, float* u,
                                       float* v, void* stream) {
    using namespace morl;
    MORL_REQUIRE(feats && wset && W1 && b1 && u && v, MORL_ERR_NULL, "morl_pair_layer1_uv_f32: NULL pointer argument");
    MORL_REQUIRE(B > 0 && W > 0 && F > 0 && D > 0 && H > 0, MORL_ERR_SHAPE, "morl_pair_layer1_uv_f32: bad shape B=%d W=%d F=%d D=%d H=%d", B, W, F, D, H);
    const int kmax = F > D ? F : D;
    const size_t smem = (size_t)kL1Rows * kmax * sizeof(float);
    MORL_REQUIRE(smem <= 48 * 1024, MORL_ERR_UNSUPPORTED, "morl_pair_layer1_uv_f32: feature dimension %d too large", kmax);
    const int blocks = (B + W + kL1Rows - 1) / kL1Rows;
    pair_layer1_uv_kernel<<<blocks, 256, smem, static_cast<cudaStream_t>(stream)>>>(feats, wset, W1, b1, B, W, F, D, H, u, v);
    return check_launch("morl_pair_layer1_uv_f32");
}
