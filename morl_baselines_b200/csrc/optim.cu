// optim.cu -- multi-tensor target-network sync.
//
// Replaces polyak_update (reference common/networks.py:121-139), which issues 1-2 tiny kernels per parameter tensor
// (10 tensors for the Envelope Q-net, 2 nets x ~14 tensors for GPI-PD), with ONE launch over a device-side table of
// (param, target, size) entries.  Arithmetic is the reference's: tau == 1 -> copy, else
//   target.mul_(1 - tau); th.add(target, param, alpha=tau, out=target)    i.e.   fma(tau, p, fl(t * (1 - tau)))
// (ATen's CPU add-with-alpha kernel is a vectorised fmadd -- probed bit-exact in the build container).
#include "common.cuh"

namespace morl {

__global__ void __launch_bounds__(256) polyak_kernel(const float* const* __restrict__ params, float* const* __restrict__ targets,
                                                     const int64_t* __restrict__ sizes, float tau, float one_minus_tau) {
    const int t = blockIdx.y;
    const int64_t n = sizes[t];
    const float* __restrict__ p = params[t];
    float* __restrict__ q = targets[t];
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += stride) {
        if (tau == 1.0f)
            q[e] = p[e];
        else
            q[e] = __fmaf_rn(tau, p[e], __fmul_rn(q[e], one_minus_tau));
    }
}

}  // namespace morl

extern "C" int morl_polyak_f32(const float* const* params, float* const* targets, const int64_t* sizes, int n_tensors, int64_t max_size,
                               double tau, void* stream) {
    using namespace morl;
    MORL_REQUIRE(params && targets && sizes, MORL_ERR_NULL, "morl_polyak_f32: NULL pointer argument");
    MORL_REQUIRE(n_tensors > 0 && n_tensors <= 65535 && max_size > 0, MORL_ERR_SHAPE, "morl_polyak_f32: bad n_tensors=%d max_size=%lld",
                 n_tensors, (long long)max_size);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    long long bx = (max_size + 255) / 256;
    if (bx > 148 * 4) bx = 148 * 4;
    const dim3 grid((unsigned)bx, (unsigned)n_tensors, 1);
    // (1 - tau) is formed in double then rounded, like Python's `1.0 - tau` handed to Tensor.mul_
    const float omt = (float)(1.0 - tau);
    polyak_kernel<<<grid, 256, 0, st>>>(params, targets, sizes, (float)tau, omt);
    return check_launch("morl_polyak_f32");
}
