// KL( N(mean, std) || N(0, 1) ) summed over a [n] block of a data encoder's posterior, and its
// gradient (sm_100a).
//
// The reference evaluates  D.kl_divergence(u, prior).sum(dim=1).mean() / n_features  per modality
// (scglue/models/scglue.py:349-353) -- with the standard-normal prior (scglue/models/sc.py:640-660)
//   kl = 0.5 (std^2 + mean^2 - 1) - log std
// per element.  Through torch.distributions that is ~10 elementwise launches forward and ~15
// backward per modality, all on the serial tail of the training step; here it is one launch
// each way.  [B, K] is a few thousand elements: one block, fixed reduction order.
#include "common.cuh"

namespace glue {
namespace {

constexpr int KL_THREADS = 512;

__global__ void __launch_bounds__(KL_THREADS) kl_unit_normal_fwd_kernel(const float* __restrict__ mean,
                                                                        const float* __restrict__ std_, int64_t n,
                                                                        float scale, float* __restrict__ out) {
    __shared__ float red[KL_THREADS / 32];
    float acc = 0.f;
    for (int64_t i = threadIdx.x; i < n; i += KL_THREADS) {
        const float m = mean[i], s = std_[i];
        acc += 0.5f * (fmaf(s, s, m * m) - 1.f) - logf(s);
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        float v = threadIdx.x < KL_THREADS / 32 ? red[threadIdx.x] : 0.f;
        v = warp_sum(v);
        if (threadIdx.x == 0) out[0] = v * scale;
    }
}

__global__ void __launch_bounds__(256) kl_unit_normal_bwd_kernel(const float* __restrict__ mean,
                                                                 const float* __restrict__ std_, int64_t n, float scale,
                                                                 const float* __restrict__ upstream,
                                                                 float* __restrict__ gmean, float* __restrict__ gstd) {
    const float c = upstream[0] * scale;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float s = std_[i];
        gmean[i] = c * mean[i];
        gstd[i] = c * (s - 1.f / s);
    }
}

}  // namespace
}  // namespace glue

extern "C" int glue_kl_unit_normal_fwd(const float* mean, const float* std_, int64_t n, float scale, float* out,
                                       void* stream) {
    if (!mean || !std_ || !out || n <= 0) return GLUE_E_BADARG;
    glue::kl_unit_normal_fwd_kernel<<<1, glue::KL_THREADS, 0, (cudaStream_t)stream>>>(mean, std_, n, scale, out);
    GLUE_CHECK_LAUNCH();
    return 0;
}

extern "C" int glue_kl_unit_normal_bwd(const float* mean, const float* std_, int64_t n, float scale,
                                       const float* upstream, float* grad_mean, float* grad_std, void* stream) {
    if (!mean || !std_ || !upstream || !grad_mean || !grad_std || n <= 0) return GLUE_E_BADARG;
    const int64_t want = (n + 255) / 256;
    const int cap = 4 * glue::sm_count();
    const int grid = (int)(want > cap ? cap : want);
    glue::kl_unit_normal_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(mean, std_, n, scale, upstream, grad_mean,
                                                                            grad_std);
    GLUE_CHECK_LAUNCH();
    return 0;
}
