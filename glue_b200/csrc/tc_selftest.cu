// Diagnostic entry point: runs a short tcgen05.mma sequence on caller-provided shared-memory
// operand images and returns the fp32 accumulator tile.  Used by tests/test_tc_selftest_gpu.py
// to pin the UMMA descriptor conventions (K-major / MN-major, SWIZZLE_128B) the fused decoder
// relies on against a plain matrix product.
#include "tc_common.cuh"

namespace glue {
namespace tc {

struct SelfTestArgs {
    const uint8_t* a_img;  // bytes copied verbatim to shared memory (1024 B aligned)
    const uint8_t* b_img;
    uint32_t a_bytes, b_bytes;
    uint64_t a_desc_base, b_desc_base;  // descriptor without the start address
    uint32_t idesc;
    uint32_t nk;  // number of MMAs (K steps)
    uint32_t a_off[16], b_off[16];  // start-address byte offset per K step
    uint32_t N;                     // accumulator columns to read back (multiple of 32, <= 256)
    float* out;                     // [128][N]
};

__global__ void __launch_bounds__(160, 1) tc_selftest_kernel(SelfTestArgs p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_sh;
    uint8_t* a_sm = smem;
    uint8_t* b_sm = smem + ((p.a_bytes + 1023u) & ~1023u);
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    for (uint32_t i = t * 16; i < p.a_bytes; i += blockDim.x * 16)
        *reinterpret_cast<uint4*>(a_sm + i) = *reinterpret_cast<const uint4*>(p.a_img + i);
    for (uint32_t i = t * 16; i < p.b_bytes; i += blockDim.x * 16)
        *reinterpret_cast<uint4*>(b_sm + i) = *reinterpret_cast<const uint4*>(p.b_img + i);
    if (t == 0) {
        mbar_init(&bar, 1);
        mbar_fence_init();
    }
    if (warp == 4) tmem_alloc(&tmem_base_sh, 256);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_sh;
    if (warp == 4 && lane == 0) {
        for (uint32_t ks = 0; ks < p.nk; ++ks) {
            const uint64_t da = smem_desc(p.a_desc_base, smem_u32(a_sm) + p.a_off[ks]);
            const uint64_t db = smem_desc(p.b_desc_base, smem_u32(b_sm) + p.b_off[ks]);
            umma_bf16(tmem_base, da, db, p.idesc, ks > 0 ? 1u : 0u);
        }
        umma_commit(&bar);
    }
    if (warp < 4) {
        mbar_wait(&bar, 0);
        tc_fence_after();
        for (uint32_t c0 = 0; c0 < p.N; c0 += 32) {
            uint32_t r[32];
            tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + c0, r);
            tmem_ld_wait();
#pragma unroll
            for (int q = 0; q < 32; ++q) p.out[(size_t)(warp * 32 + lane) * p.N + c0 + q] = __uint_as_float(r[q]);
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 4) tmem_dealloc(tmem_base, 256);
}

}  // namespace tc
}  // namespace glue

extern "C" int glue_tc_selftest(const void* a_img, uint32_t a_bytes, const void* b_img, uint32_t b_bytes,
                                uint64_t a_desc_base, uint64_t b_desc_base, uint32_t idesc, uint32_t nk,
                                const uint32_t* a_off, const uint32_t* b_off, uint32_t N, float* out,
                                void* stream) {
    using namespace glue::tc;
    if (!a_img || !b_img || !out || nk == 0 || nk > 16 || N == 0 || N > 256 || (N % 32) != 0) return GLUE_E_BADARG;
    if ((a_bytes % 16) != 0 || (b_bytes % 16) != 0) return GLUE_E_ALIGN;
    SelfTestArgs p;
    p.a_img = (const uint8_t*)a_img, p.b_img = (const uint8_t*)b_img;
    p.a_bytes = a_bytes, p.b_bytes = b_bytes;
    p.a_desc_base = a_desc_base, p.b_desc_base = b_desc_base;
    p.idesc = idesc, p.nk = nk, p.N = N, p.out = out;
    for (uint32_t i = 0; i < 16; ++i) p.a_off[i] = i < nk ? a_off[i] : 0, p.b_off[i] = i < nk ? b_off[i] : 0;
    const size_t smem = (size_t)((a_bytes + 1023u) & ~1023u) + ((b_bytes + 1023u) & ~1023u) + 1024;
    if (smem > 200 * 1024) return GLUE_E_UNSUPPORTED;
    GLUE_CUDA(cudaFuncSetAttribute(tc_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    tc_selftest_kernel<<<1, 160, smem, (cudaStream_t)stream>>>(p);
    GLUE_CHECK_LAUNCH();
    return 0;
}
