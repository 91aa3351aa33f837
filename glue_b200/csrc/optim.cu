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

// ---- minibatch -> static input buffers of a captured step: up to 16 device-to-device copies in ONE launch ----
// (14 individual copies were 60 us of launch latency in front of every replayed step).  Blocks are dealt out
// to the segments in proportion to their size; 16-byte vectors when both ends are aligned.
constexpr int MC_MAX = 16;
struct MultiCopy {
    const void* src[MC_MAX];
    void* dst[MC_MAX];
    long long bytes[MC_MAX];
    int first_block[MC_MAX + 1];  // segment s owns blocks [first_block[s], first_block[s + 1])
    int n;
};

__global__ void __launch_bounds__(256) multi_copy_kernel(const MultiCopy a) {
    int s = 0;
    while (s + 1 < a.n && (int)blockIdx.x >= a.first_block[s + 1]) ++s;
    const int nb = a.first_block[s + 1] - a.first_block[s], b = blockIdx.x - a.first_block[s];
    const char* src = static_cast<const char*>(a.src[s]);
    char* dst = static_cast<char*>(a.dst[s]);
    const long long bytes = a.bytes[s];
    const long long tid = (long long)b * blockDim.x + threadIdx.x, nth = (long long)nb * blockDim.x;
    long long done = 0;
    if (((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15u) == 0) {
        const long long n16 = bytes >> 4;
        const uint4* s4 = reinterpret_cast<const uint4*>(src);
        uint4* d4 = reinterpret_cast<uint4*>(dst);
        for (long long i = tid; i < n16; i += nth) d4[i] = __ldg(s4 + i);
        done = n16 << 4;
    }
    for (long long i = done + tid; i < bytes; i += nth) dst[i] = src[i];
}

}  // namespace
}  // namespace glue

extern "C" int glue_multi_copy(int32_t n, const void* const* src, void* const* dst, const int64_t* bytes, void* stream) {
    using namespace glue;
    if (n <= 0 || n > MC_MAX || !src || !dst || !bytes) return GLUE_E_BADARG;
    MultiCopy a;
    a.n = n;
    int total = 0;
    const int sms = sm_count();
    for (int s = 0; s < n; ++s) {
        if (!src[s] || !dst[s] || bytes[s] < 0) return GLUE_E_BADARG;
        a.src[s] = src[s], a.dst[s] = dst[s], a.bytes[s] = bytes[s];
        a.first_block[s] = total;
        // one block per 64 KB, at least one, at most four per SM for the largest segments
        long long nb = (bytes[s] + 65535) / 65536;
        if (nb < 1) nb = 1;
        if (nb > 4LL * sms) nb = 4LL * sms;
        total += (int)nb;
    }
    a.first_block[n] = total;
    multi_copy_kernel<<<total, 256, 0, (cudaStream_t)stream>>>(a);
    GLUE_CHECK_LAUNCH();
    return 0;
}

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
