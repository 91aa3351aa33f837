// Shared device helpers for libglue_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "glue_b200.h"

#define GLUE_EPS 1e-7f  // scglue/num.py:12

// Diagnostic only: number of kernels this library has launched in this process.
extern "C" void glue_count_launch_(void);

#define GLUE_CHECK_LAUNCH()                      \
    do {                                         \
        glue_count_launch_();                    \
        cudaError_t e__ = cudaPeekAtLastError(); \
        if (e__ != cudaSuccess) return (int)e__; \
    } while (0)

#define GLUE_CUDA(call)                          \
    do {                                         \
        cudaError_t e__ = (call);                \
        if (e__ != cudaSuccess) return (int)e__; \
    } while (0)

// Opt a kernel instantiation in to > 48 KB of dynamic shared memory, once per process (the
// call site is instantiated per kernel; the value is a per-function attribute).
#define GLUE_SMEM_ONCE(kernel, bytes)                                                            \
    do {                                                                                         \
        static bool done__ = false;                                                              \
        if (!done__) {                                                                           \
            GLUE_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,  \
                                           (int)(bytes)));                                       \
            done__ = true;                                                                       \
        }                                                                                        \
    } while (0)

namespace glue {

constexpr unsigned FULL = 0xffffffffu;

// ---- programmatic dependent launch (PDL) -------------------------------------------------------
// A kernel launched with launch_pdl may start (barrier set-up, staging of its own inputs) while the
// kernel in front of it on the stream is still draining; griddep_wait() blocks until that kernel
// has completed and its writes are visible, and has to precede the first access to anything the
// earlier kernel wrote or read.  griddep_launch() lets the kernel behind start early.  Both are
// no-ops without a programmatic dependency.
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void griddep_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(bool enable, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                              cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid, cfg.blockDim = block, cfg.dynamicSmemBytes = smem, cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr, cfg.numAttrs = enable ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

__device__ __forceinline__ float warp_sum(float x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(FULL, x, o);
    return x;
}
__device__ __forceinline__ float warp_max(float x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x = fmaxf(x, __shfl_xor_sync(FULL, x, o));
    return x;
}

// softplus with torch's threshold (F.softplus: x > 20 -> x)
__device__ __forceinline__ float softplusf(float x) { return x > 20.f ? x : log1pf(expf(x)); }
__device__ __forceinline__ float sigmoidf(float x) { return 1.f / (1.f + expf(-x)); }

// digamma for x > 0 (recurrence up to 6, then asymptotic series)
__device__ __forceinline__ float digammaf(float x) {
    float r = 0.f;
    while (x < 6.f) {
        r -= 1.f / x;
        x += 1.f;
    }
    float inv = 1.f / x, inv2 = inv * inv;
    // ln x - 1/(2x) - 1/(12x^2) + 1/(120x^4) - 1/(252x^6)
    float series = inv2 * (-1.f / 12.f + inv2 * (1.f / 120.f - inv2 * (1.f / 252.f)));
    return r + logf(x) - 0.5f * inv + series;
}

// For a count x >= 0 and theta > 0:
//   lg = lgamma(theta + x) - lgamma(theta) - lgamma(1 + x)
//   dg = digamma(theta + x) - digamma(theta)
// Small integer counts use the exact product form (no cancellation); everything
// else goes through lgammaf / digammaf.
__device__ __forceinline__ void nb_gamma_terms(float x, float theta, float& lg, float& dg) {
    if (x == 0.f) {
        lg = 0.f;
        dg = 0.f;
        return;
    }
    if (x <= 12.f && x == floorf(x)) {
        float prod = 1.f, h = 0.f, quot = 1.f;
        int n = (int)x;
        for (int m = 0; m < n; ++m) {
            float t = theta + (float)m;
            prod *= t;
            quot *= (float)(m + 1);
            h += 1.f / t;
        }
        // prod can overflow only for theta ~ 1e3+ with x = 12 (1e36 < FLT_MAX) -- safe
        lg = logf(prod) - logf(quot);
        dg = h;
        return;
    }
    lg = lgammaf(theta + x) - lgammaf(theta) - lgammaf(1.f + x);
    dg = digammaf(theta + x) - digammaf(theta);
}

__device__ __forceinline__ void cp_async4(void* smem, const void* gmem, bool pred) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    int sz = pred ? 4 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;\n" ::"r"(s), "l"(gmem), "r"(sz));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

inline int sm_count() {
    int dev = 0, n = 148;
    if (cudaGetDevice(&dev) == cudaSuccess)
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    return n > 0 ? n : 148;
}

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

}  // namespace glue
