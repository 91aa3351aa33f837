// Philox4x32-10 keep flags shared by the hidden-block kernels (mlp.cu, mlp_fused.cu): the same
// (seed, call, row, column) gives the same flag in both, so the fused and the per-op paths draw
// identical dropout masks.
#pragma once
#include <stdint.h>

namespace glue {

__device__ __forceinline__ uint32_t mulhilo(uint32_t a, uint32_t b, uint32_t& hi) {
    const uint64_t p = (uint64_t)a * b;
    hi = (uint32_t)(p >> 32);
    return (uint32_t)p;
}

// Philox4x32-10 (Salmon et al. 2011): counter (c0..c3), key (k0, k1) -> 4 x 32 random bits
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0, hi1;
        const uint32_t lo0 = mulhilo(0xD2511F53u, c.x, hi0);
        const uint32_t lo1 = mulhilo(0xCD9E8D57u, c.z, hi1);
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += 0x9E3779B9u, k.y += 0xBB67AE85u;
    }
    return c;
}

// keep flag of element (row, col): one Philox block per (row, col / 4)
__device__ __forceinline__ bool keep_flag(uint64_t seed, uint64_t call, int row, int col, uint32_t threshold) {
    const uint4 r = philox4x32_10(make_uint4((uint32_t)row, (uint32_t)(col >> 2), (uint32_t)call, (uint32_t)(call >> 32)),
                                  make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
    const uint32_t v = (col & 3) == 0 ? r.x : ((col & 3) == 1 ? r.y : ((col & 3) == 2 ? r.z : r.w));
    return v >= threshold;  // P(keep) = 1 - threshold / 2^32
}

}  // namespace glue
