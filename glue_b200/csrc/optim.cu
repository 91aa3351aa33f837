// Fused RMSprop update over a flat parameter range (sm_100a).
//
// One pass over (param, grad, square_avg) replaces the 6-7 multi-tensor launches per optimizer
// step of torch.optim.RMSprop (momentum = 0, centered = False, weight_decay = 0 -- the
// configuration scglue uses, scglue/models/scglue.py:889-942):
//
//   sq  = alpha * sq + (1 - alpha) * g * g
//   p  -= lr * g / (sqrt(sq) + eps)
//
// HBM-bound: 20 bytes per parameter.  Each element is independent, so the result does not
// depend on the launch geometry.
#include "common.cuh"

namespace glue {
namespace {

__device__ __forceinline__ void rms1(float& p, float g, float& sq, float lr, float alpha, float oma, float eps) {
    sq = fmaf(oma * g, g, alpha * sq);      // as ATen: mul_(alpha) then addcmul_(g, g, 1 - alpha)
    const float avg = sqrtf(sq) + eps;
    p = fmaf(-lr, g / avg, p);              // addcdiv_(g, avg, value = -lr)
}

__global__ void __launch_bounds__(256) rmsprop_flat_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                           float* __restrict__ sq, int64_t n, float lr, float alpha,
                                                           float eps) {
    const float oma = 1.f - alpha;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nth = (int64_t)gridDim.x * blockDim.x;
    const bool vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) |
                       reinterpret_cast<uintptr_t>(sq)) & 15u) == 0;
    int64_t done = 0;
    if (vec) {
        const int64_t n4 = n >> 2;
        float4* p4 = reinterpret_cast<float4*>(p);
        const float4* g4 = reinterpret_cast<const float4*>(g);
        float4* s4 = reinterpret_cast<float4*>(sq);
        for (int64_t i = tid; i < n4; i += nth) {
            float4 pv = p4[i], sv = s4[i];
            const float4 gv = g4[i];
            rms1(pv.x, gv.x, sv.x, lr, alpha, oma, eps);
            rms1(pv.y, gv.y, sv.y, lr, alpha, oma, eps);
            rms1(pv.z, gv.z, sv.z, lr, alpha, oma, eps);
            rms1(pv.w, gv.w, sv.w, lr, alpha, oma, eps);
            p4[i] = pv, s4[i] = sv;
        }
        done = n4 << 2;
    }
    for (int64_t i = done + tid; i < n; i += nth) {
        float pv = p[i], sv = sq[i];
        rms1(pv, g[i], sv, lr, alpha, oma, eps);
        p[i] = pv, sq[i] = sv;
    }
}

}  // namespace
}  // namespace glue

extern "C" int glue_rmsprop_flat(float* param, const float* grad, float* square_avg, int64_t n, float lr,
                                 float alpha, float eps, void* stream) {
    if (!param || !grad || !square_avg || n <= 0) return GLUE_E_BADARG;
    const int64_t want = (n / 4 + 255) / 256;
    const int cap = 8 * glue::sm_count();
    const int grid = (int)(want < 1 ? 1 : (want > cap ? cap : want));
    glue::rmsprop_flat_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(param, grad, square_avg, n, lr, alpha, eps);
    GLUE_CHECK_LAUNCH();
    return 0;
}
