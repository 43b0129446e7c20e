// dyna.cu -- GPI-PD's Dyna planning step on the device.
//
// morl_ensemble_sample_f32 fuses everything between the last EnsembleLayer of the probabilistic ensemble and the imagined transition
// (reference common/model_based/probabilistic_ensemble.py:115-154 + common/model_based/utils.py:162-170):
//     mean, logvar = chunk(out, 2)                                                              (:115)
//     logvar = max_logvar - softplus(max_logvar - logvar);  logvar = min_logvar + softplus(logvar - min_logvar)   (:118-119)
//     samples = mean + exp(0.5 logvar) * noise                                                  (:127-128; noise = th.randn, or none: deterministic)
//     vars = exp(logvar); mean_ens = mean_e(mean); var_ens = mean_e(mean^2 + vars) - mean_ens^2  (:140-147)
//     uncertainty = sum_o sqrt(var_ens + 1e-12)                                                  (:148-149)
//     sample / var of the elite model drawn for the row (model_inds, :143, :152-154), sample[:, rew_dim:] += obs (utils.py:165)
// The reference moves three [E, N, O] tensors to the host and does this in numpy; here the raw [E, N, 2 O] output is read once.
// One warp per row; bound: HBM (E * 2 O * 4 bytes per row read, 2 O * 4 + 4 written).
#include "common.cuh"

namespace morl {

__device__ __forceinline__ float softplus_t(float x) { return x > 20.f ? x : log1pf(expf(x)); }  // torch F.softplus (beta 1, threshold 20)

__global__ void __launch_bounds__(256) ensemble_sample_kernel(const float* __restrict__ out, const float* __restrict__ max_logvar,
                                                              const float* __restrict__ min_logvar, const int32_t* __restrict__ model_idx,
                                                              const float* __restrict__ noise, const float* __restrict__ obs, int rew_dim, int E, int N,
                                                              int O, float* __restrict__ sample_out, float* __restrict__ var_out,
                                                              float* __restrict__ unc_out) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= N) return;
    const int n = warp;
    const int pick = model_idx[n];
    float unc = 0.f;
    for (int o = lane; o < O; o += 32) {
        const float hi = __ldg(max_logvar + o), lo = __ldg(min_logvar + o);
        float sum_m = 0.f, sum_s = 0.f, s_pick = 0.f, v_pick = 0.f;
        for (int e = 0; e < E; ++e) {
            const float* row = out + ((size_t)e * N + n) * (2 * O);
            const float m = __ldg(row + o);
            float lv = __ldg(row + O + o);
            lv = __fsub_rn(hi, softplus_t(__fsub_rn(hi, lv)));
            lv = __fadd_rn(lo, softplus_t(__fsub_rn(lv, lo)));
            const float var = expf(lv);
            // ensemble moments in the reference's order: sequential sums over the models, one divide at the end (numpy mean over axis 0)
            sum_m = e == 0 ? m : __fadd_rn(sum_m, m);
            const float t = __fadd_rn(__fmul_rn(m, m), var);
            sum_s = e == 0 ? t : __fadd_rn(sum_s, t);
            if (e == pick) {
                v_pick = var;
                s_pick = noise ? __fadd_rn(m, __fmul_rn(expf(__fmul_rn(0.5f, lv)), __ldg(noise + ((size_t)e * N + n) * O + o))) : m;
            }
        }
        const float mean_e = __fdiv_rn(sum_m, (float)E);
        const float var_e = __fsub_rn(__fdiv_rn(sum_s, (float)E), __fmul_rn(mean_e, mean_e));
        unc += sqrtf(__fadd_rn(var_e, 1e-12f));
        if (obs && o >= rew_dim) s_pick = __fadd_rn(s_pick, __ldg(obs + (size_t)n * (O - rew_dim) + (o - rew_dim)));
        sample_out[(size_t)n * O + o] = s_pick;
        var_out[(size_t)n * O + o] = v_pick;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) unc += __shfl_xor_sync(0xffffffffu, unc, off);
    if (lane == 0) unc_out[n] = unc;
}

}  // namespace morl

extern "C" int morl_ensemble_sample_f32(const float* out, const float* max_logvar, const float* min_logvar, const int32_t* model_idx, const float* noise,
                                        const float* obs, int rew_dim, int E, int N, int O, float* sample_out, float* var_out, float* uncertainty_out,
                                        void* stream) {
    using namespace morl;
    MORL_REQUIRE(out && max_logvar && min_logvar && model_idx && sample_out && var_out && uncertainty_out, MORL_ERR_NULL,
                 "morl_ensemble_sample_f32: NULL pointer argument");
    MORL_REQUIRE(E > 0 && N > 0 && O > 0 && rew_dim >= 0 && rew_dim <= O, MORL_ERR_SHAPE, "morl_ensemble_sample_f32: bad shape E=%d N=%d O=%d rew_dim=%d", E, N, O,
                 rew_dim);
    const int warps_per_block = 8;
    const unsigned grid = (unsigned)((N + warps_per_block - 1) / warps_per_block);
    ensemble_sample_kernel<<<grid, warps_per_block * 32, 0, static_cast<cudaStream_t>(stream)>>>(out, max_logvar, min_logvar, model_idx, noise, obs, rew_dim, E, N, O,
                                                                                                 sample_out, var_out, uncertainty_out);
    return check_launch("morl_ensemble_sample_f32");
}
