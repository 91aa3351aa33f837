// Data-encoder MLP (scglue/models/sc.py:62-178) as three launches forward and three backward (sm_100a).
//
// The reference builds an encoder as h_depth x [Linear -> LeakyReLU(0.2) -> BatchNorm1d -> Dropout] followed by
// the `loc` / `std_lin` heads.  On a 128-cell minibatch every Linear is a [128, 100..256] x [256, 256] product:
// through cuBLAS that is a 32x32-tile SIMT sgemm (7-8 us) plus a bias kernel per layer, and with the hidden-block
// kernel of mlp.cu and the element-wise tail of the heads an encoder pass is ~14 launches forward and ~20 backward
// of 2-10 us each -- launch latency on the critical path of the step, not work.  Here:
//
//   forward   mlp_hidden_fwd_kernel (per hidden block): Linear + bias + LeakyReLU + BatchNorm (training mode,
//             batch statistics, running statistics updated) + Dropout;  mlp_heads_fwd_kernel: both heads,
//             softplus + EPS on the second;
//   backward  mlp_heads_bwd_kernel: d/d(loc, std) -> G = [d loc | d std sigmoid(pre)], head weight / bias
//             gradients;  mlp_hidden_bwd_kernel (per hidden block, top down): the input gradient of the layer
//             above restricted to this CTA's columns (G_above W_above[:, cols]), Dropout / BatchNorm /
//             LeakyReLU backward, bias / gamma / beta gradients, and the weight gradient dz^T in.
//
// Every kernel gives a CTA 16 output columns and all (<= 128) rows: BatchNorm's column statistics, the bias
// gradients and the weight-gradient reductions over the batch are CTA-local and summed in a fixed order (results
// are bit-reproducible).  The products run on warp-level tensor cores, mma.sync m16n8k8 TF32 with every fp32
// product rebuilt from three TF32 products (small x big + big x small + big x big, fp32 accumulation: ~1e-6
// relative, two orders inside the 1e-4 parity bar; a first version on FFMA register tiles was bound by the
// FP32 pipe and shared-memory reads at 8 us per 256-wide layer on the 8 SMs it occupied).  Operands are staged in
// shared memory (cp.async for the fp32 side that a warp splits on the fly, a pre-split TF32 big / small image for
// the side all warps share).  Dropout masks are the Philox masks of mlp.cu (same seed / call / row / column).
//
// Fragment layouts (PTX ISA, m16n8k8 .tf32): g = lane >> 2, t = lane & 3;
//   A (16 x 8, row):  a0 (g, t)  a1 (g + 8, t)  a2 (g, t + 4)  a3 (g + 8, t + 4)
//   B (8 x 8, col):   b0 (k = t, n = g)  b1 (k = t + 4, n = g)
//   C (16 x 8):       c0 (g, 2 t)  c1 (g, 2 t + 1)  c2 (g + 8, 2 t)  c3 (g + 8, 2 t + 1)
#include "common.cuh"
#include "philox.cuh"
#include "tc_common.cuh"

namespace glue {
namespace {

constexpr int MF_THREADS = 512;  // 16 warps: row block w & 7, half w >> 3 of the reduction dimension
constexpr int MF_ROWS = 128;  // cells per call
constexpr int MF_COLS = 16;   // output columns per CTA
constexpr int MF_KMAX = 256;  // widest staged operand row
constexpr int MF_GS_LD = 24;  // pitch of the [128][16] gradient slab (8 mod 32: its transposed reads are conflict-free)

// operand pitches (floats) for k columns (k a multiple of 8): `row-major reads` (rows g, columns t) want
// 4 mod 32, `transposed reads` (rows t, columns g) want 8 mod 32
__host__ __device__ inline int pitch4(int k) { return k + ((36 - (k & 31)) & 31); }
__host__ __device__ inline int pitch8(int k) { return k + ((40 - (k & 31)) & 31); }
__host__ __device__ inline int round8(int k) { return (k + 7) & ~7; }

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ uint32_t to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ void split_tf32(float x, uint32_t& big, uint32_t& small) {
    big = to_tf32(x);
    small = to_tf32(x - __uint_as_float(big));
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
        "{%0, %1, %2, %3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// rows [0, 128) x k floats into dst[row][ld] (kc real columns, the rest zeros; rows beyond n_rows zeros), completing
// phase `bar`.  `bulk` (kc a multiple of 4, 16-byte aligned rows): ONE bulk copy per row (cp.async.bulk, completing
// on `bar`; a first version issued 16-byte cp.async and spent 4 k cycles on issuing them), padding by plain stores.
// Otherwise (e.g. the 50-wide input of the discriminator): plain loads and stores by all threads.  Every thread then
// waits on `bar` and the CTA synchronises.  `reuse`: the destination was read through the generic proxy before (the
// caller has synchronised the CTA): order those reads before the asynchronous writes.
template <typename RowPtr>
__device__ __forceinline__ void stage_rows(float* dst, int ld, int kc, int k, int n_rows, RowPtr row_ptr, uint64_t* bar,
                                           bool reuse, bool bulk) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (!bulk) {
        if (tid == 0) tc::mbar_arrive(bar);
        for (int idx = tid; idx < MF_ROWS * k; idx += MF_THREADS) {
            const int r = idx / k, j = idx - r * k;
            dst[r * ld + j] = (r < n_rows && j < kc) ? __ldg(row_ptr(r) + j) : 0.f;
        }
        return;
    }
    if (tid == 0) tc::mbar_arrive_expect_tx(bar, (uint32_t)(n_rows < MF_ROWS ? n_rows : MF_ROWS) * (uint32_t)kc * 4u);
    if (lane < 8) {  // row = 16 lane + warp: 8 copies per warp (a warp issues its bulk copies one lane at a time)
        const int r = 16 * lane + warp;
        float* d = dst + r * ld;
        if (r < n_rows) {
            if (reuse) tc::fence_proxy_async();
            tc::bulk_load(d, row_ptr(r), (uint32_t)kc * 4u, bar);
        } else {
            for (int q = 0; q < k; q += 4) *reinterpret_cast<float4*>(d + q) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    } else if (lane < 16 && k > kc) {
        const int r = 16 * (lane - 8) + warp;
        if (r < n_rows)
            for (int q = kc; q < k; q += 4) *reinterpret_cast<float4*>(dst + r * ld + q) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
}
__device__ __forceinline__ void init_bar(uint64_t* bar) {
    if (threadIdx.x == 0) {
        tc::mbar_init(bar, 1);
        tc::mbar_fence_init();
    }
    __syncthreads();
}

// C[16 wr + {g, g + 8}][8 nt + {2 t, 2 t + 1}] += sum_k A[row][k] Bt[col][k] over the 8-wide steps [s_begin, s_end);
// A fp32 [128][ld] split on the fly, Bt pre-split TF32 big / small [16][ld].  The three TF32 products of an fp32
// product go to three accumulators (one dependent mma per step and accumulator instead of three); the two small
// ones are added first, then the big one.
__device__ __forceinline__ void gemm_mma(const float* __restrict__ As, const uint32_t* __restrict__ Bb,
                                         const uint32_t* __restrict__ Bs, int ld, int s_begin, int s_end, int wr, int lane,
                                         float (&c)[2][4]) {
    const int g = lane >> 2, t = lane & 3;
    const float* a_lo = As + (16 * wr + g) * ld + t;
    const float* a_hi = a_lo + 8 * ld;
    const uint32_t* bb = Bb + g * ld + t;
    const uint32_t* bs = Bs + g * ld + t;
    float c1[2][4], c2[2][4];
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) c1[nt][e] = 0.f, c2[nt][e] = 0.f;
#pragma unroll 4
    for (int s = s_begin; s < s_end; ++s) {
        const int k0 = 8 * s;
        uint32_t ab[4], as[4];
        split_tf32(a_lo[k0], ab[0], as[0]);
        split_tf32(a_hi[k0], ab[1], as[1]);
        split_tf32(a_lo[k0 + 4], ab[2], as[2]);
        split_tf32(a_hi[k0 + 4], ab[3], as[3]);
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) {
            const int o = 8 * nt * ld + k0;
            mma_tf32(c1[nt], as, bb[o], bb[o + 4]);
            mma_tf32(c2[nt], ab, bs[o], bs[o + 4]);
            mma_tf32(c[nt], ab, bb[o], bb[o + 4]);
        }
    }
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) c[nt][e] += c1[nt][e] + c2[nt][e];
}

// The two warps that share a row block (w and w ^ 8) each hold a partial 16 x 16 tile over half of the reduction
// dimension.  Afterwards warp half h owns row g + 8 h of the block: v[j] = both partials of (that row, column j).
__device__ __forceinline__ void combine_halves(const float (&c)[2][4], float (*xch)[32][4], int warp, int lane,
                                               float (&v)[4]) {
    const bool h = (warp >> 3) != 0;
#pragma unroll
    for (int j = 0; j < 4; ++j)  // the other half's row
        xch[warp][lane][j] = h ? c[j >> 1][j & 1] : c[j >> 1][2 + (j & 1)];
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = (h ? c[j >> 1][2 + (j & 1)] : c[j >> 1][j & 1]) + xch[warp ^ 8][lane][j];
}

// out[o][n] = sum_r Gs[r][o] In[r][n] (o < 16, r < 128): warp w owns n-tiles w and w + 16 (n_tiles <= 32);
// Gs [128][MF_GS_LD], In [128][ld] with ld = 8 mod 32; both fp32, split on the fly.
//   A[m = o][k = r] = Gs[r][o];  B[k = r][n] = In[r][n]
__device__ __forceinline__ void wgrad_mma(const float* __restrict__ Gs, const float* __restrict__ In, int ld, int n_tiles,
                                          int warp, int lane, float (&c)[2][4]) {
    const int g = lane >> 2, t = lane & 3;
    if (warp >= n_tiles) return;
    float c1[2][4], c2[2][4];  // (the two small TF32 products: independent accumulators, as in gemm_mma)
#pragma unroll
    for (int q = 0; q < 2; ++q)
#pragma unroll
        for (int e = 0; e < 4; ++e) c1[q][e] = 0.f, c2[q][e] = 0.f;
#pragma unroll 2
    for (int k0 = 0; k0 < MF_ROWS; k0 += 8) {
        uint32_t ab[4], as[4];
        split_tf32(Gs[(k0 + t) * MF_GS_LD + g], ab[0], as[0]);
        split_tf32(Gs[(k0 + t) * MF_GS_LD + g + 8], ab[1], as[1]);
        split_tf32(Gs[(k0 + t + 4) * MF_GS_LD + g], ab[2], as[2]);
        split_tf32(Gs[(k0 + t + 4) * MF_GS_LD + g + 8], ab[3], as[3]);
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int nt = warp + 16 * q;
            if (nt < n_tiles) {
                uint32_t bb0, bs0, bb1, bs1;
                split_tf32(In[(k0 + t) * ld + 8 * nt + g], bb0, bs0);
                split_tf32(In[(k0 + t + 4) * ld + 8 * nt + g], bb1, bs1);
                mma_tf32(c1[q], as, bb0, bb1);
                mma_tf32(c2[q], ab, bs0, bs1);
                mma_tf32(c[q], ab, bb0, bb1);
            }
        }
    }
#pragma unroll
    for (int q = 0; q < 2; ++q)
#pragma unroll
        for (int e = 0; e < 4; ++e) c[q][e] += c1[q][e] + c2[q][e];
}

// column totals over the 128 rows of per-thread partials v[j] (columns 8 (j >> 1) + 2 t + (j & 1)): lanes of a warp
// with equal t, then the 8 warps in order; every thread ends up with the totals of its own 4 columns
__device__ __forceinline__ void col_reduce4(float (&v)[4], float (*red)[MF_COLS], int warp, int lane) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        v[j] += __shfl_xor_sync(FULL, v[j], 4);
        v[j] += __shfl_xor_sync(FULL, v[j], 8);
        v[j] += __shfl_xor_sync(FULL, v[j], 16);
    }
    __syncthreads();  // (red may still be read from the previous reduction)
    const int t = lane & 3;
    if (lane < 4) {
#pragma unroll
        for (int j = 0; j < 4; ++j) red[warp][8 * (j >> 1) + 2 * t + (j & 1)] = v[j];
    }  // (MF_THREADS / 32 = 16 rows of red)
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < MF_THREADS / 32; ++w) s += red[w][8 * (j >> 1) + 2 * t + (j & 1)];
        v[j] = s;
    }
}

__device__ __forceinline__ float leaky(float z, float slope) { return z > 0.f ? z : slope * z; }

// After combine_halves a thread owns row 16 (warp & 7) + g + 8 (warp >> 3) and columns c0 + 8 (j >> 1) + 2 t + (j & 1).

// keep flags of two adjacent columns (2 t, 2 t + 1 of an 8-column tile share a Philox block)
__device__ __forceinline__ void keep_flag2(uint64_t seed, uint64_t call, int row, int col_even, uint32_t threshold,
                                           bool& k0, bool& k1) {
    const uint4 r = philox4x32_10(
        make_uint4((uint32_t)row, (uint32_t)(col_even >> 2), (uint32_t)call, (uint32_t)(call >> 32)),
        make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
    const bool hi = (col_even & 2) != 0;
    k0 = (hi ? r.z : r.x) >= threshold;
    k1 = (hi ? r.w : r.y) >= threshold;
}

// ------------------------------------------------------------------------------------ forward
struct HidFwd {
    const float* in;  // [B, Din], row pitch ld_in
    long long ld_in;
    const float* W;     // [H, Din]
    const float* bias;  // [H]
    const float* gamma;  // [H] or NULL (no BatchNorm)
    const float* beta;
    float* running_mean;
    float* running_var;
    long long* num_batches;
    unsigned long long* rng;  // [3]: seed, call counter, ticket (NULL: no dropout)
    float* z;                 // [B, H] Linear output, kept for the backward pass (NULL: not kept)
    float* out;               // [B, H]
    float* save_mean;
    float* save_rstd;
    unsigned char* keep;  // [B, H] (NULL: not kept)
    int B, Din, H;
    float slope, eps, momentum, p;
    int bulk;  // rows of `in` can be fetched with bulk copies
};

// Bt slab of a Linear: rows = output columns c0 .. c0 + 15 of W [H, K] (row_ptr(r) NULL: zeros), pre-split.
// 32 threads per row, 4 floats each per sweep; every load of a thread is issued before the first use.  Rows that
// are not 16-byte aligned (K not a multiple of 4) are read with scalar loads.
template <typename RowPtr>
__device__ __forceinline__ void stage_split(uint32_t* Bb, uint32_t* Bs, int ld, int kc, int k, RowPtr row_ptr) {
    const int r = threadIdx.x >> 5, q = threadIdx.x & 31;  // (MF_THREADS / 32 = MF_COLS rows)
    const float* src = row_ptr(r);
    const bool vec = (reinterpret_cast<uintptr_t>(src) & 15u) == 0 && (kc & 3) == 0;
    float4 v[2];
#pragma unroll
    for (int m = 0; m < 2; ++m) {
        const int j = 4 * (q + 32 * m);
        v[m] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (src == nullptr || j >= kc) continue;
        if (vec) {
            v[m] = __ldg(reinterpret_cast<const float4*>(src + j));
        } else {
            v[m].x = __ldg(src + j);
            if (j + 1 < kc) v[m].y = __ldg(src + j + 1);
            if (j + 2 < kc) v[m].z = __ldg(src + j + 2);
            if (j + 3 < kc) v[m].w = __ldg(src + j + 3);
        }
    }
#pragma unroll
    for (int m = 0; m < 2; ++m) {
        const int j = 4 * (q + 32 * m);
        if (j >= k) continue;
        uint4 b, sm;
        split_tf32(v[m].x, b.x, sm.x), split_tf32(v[m].y, b.y, sm.y);
        split_tf32(v[m].z, b.z, sm.z), split_tf32(v[m].w, b.w, sm.w);
        *reinterpret_cast<uint4*>(Bb + r * ld + j) = b;
        *reinterpret_cast<uint4*>(Bs + r * ld + j) = sm;
    }
}

// shared-memory carve-up of the forward kernels: barrier (16 B) | A [128][ld] | Bt big, small [16][ld] each |
// exchange | red
struct FwdSmem {
    uint64_t* bar;
    float* As;
    uint32_t *Bb, *Bs;
    float (*xch)[32][4];
    float (*red)[MF_COLS];
};
__device__ __forceinline__ FwdSmem carve_fwd(float* smem, int ld) {
    FwdSmem S;
    S.bar = reinterpret_cast<uint64_t*>(smem);
    S.As = smem + 4;
    S.Bb = reinterpret_cast<uint32_t*>(S.As + MF_ROWS * ld);
    S.Bs = S.Bb + MF_COLS * ld;
    S.xch = reinterpret_cast<float(*)[32][4]>(S.Bs + MF_COLS * ld);
    S.red = reinterpret_cast<float(*)[MF_COLS]>(S.xch + MF_THREADS / 32);
    return S;
}
constexpr int MF_AUX_FLOATS = 4 + (MF_THREADS / 32) * (32 * 4 + MF_COLS);  // barrier + exchange + red

__global__ void __launch_bounds__(MF_THREADS, 1) mlp_hidden_fwd_kernel(const HidFwd a) {
    extern __shared__ __align__(16) float smem[];
    const int k = round8(a.Din), ld = pitch4(k), k8 = k >> 3;
    const FwdSmem S = carve_fwd(smem, ld);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int wr = warp & 7, h = warp >> 3;
    const int c0 = blockIdx.x * MF_COLS;
    const int row_base = blockIdx.y * MF_ROWS;  // (several row tiles only without BatchNorm: rows are independent)
    const int n_rows = min(MF_ROWS, a.B - row_base);
    init_bar(S.bar);
    stage_rows(S.As, ld, a.Din, k, n_rows, [&](int r) { return a.in + (long long)(row_base + r) * a.ld_in; }, S.bar,
               false, a.bulk != 0);
    stage_split(S.Bb, S.Bs, ld, a.Din, k,
                [&](int r) { return c0 + r < a.H ? a.W + (long long)(c0 + r) * a.Din : nullptr; });
    // per-column parameters of my 4 columns (requested before the product is waited for)
    bool cok[4];
    int col[4];
    float bias[4], gam[4], bet[4], rmean[4], rvar[4];
    const bool stat_writer = warp == 0 && g == 0;  // lanes 0 .. 3 of warp 0: the 16 columns
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        col[j] = c0 + 8 * (j >> 1) + 2 * t + (j & 1);
        cok[j] = col[j] < a.H;
        bias[j] = cok[j] ? a.bias[col[j]] : 0.f;
        gam[j] = (cok[j] && a.gamma) ? a.gamma[col[j]] : 0.f;
        bet[j] = (cok[j] && a.gamma) ? a.beta[col[j]] : 0.f;
        rmean[j] = (stat_writer && cok[j] && a.running_mean) ? a.running_mean[col[j]] : 0.f;
        rvar[j] = (stat_writer && cok[j] && a.running_var) ? a.running_var[col[j]] : 0.f;
    }
    const bool drop = a.p > 0.f && a.rng != nullptr;
    unsigned long long seed = 0, call = 0;
    if (drop) seed = a.rng[0], call = a.rng[1];
    tc::mbar_wait(S.bar, 0);
    __syncthreads();
    float acc[2][4];
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[nt][e] = 0.f;
    const int s_mid = (k8 + 1) >> 1;
    gemm_mma(S.As, S.Bb, S.Bs, ld, h ? s_mid : 0, h ? k8 : s_mid, wr, lane, acc);
    float zv[4], y[4];
    combine_halves(acc, S.xch, warp, lane, zv);
    const int row = row_base + 16 * wr + g + 8 * h;
    const bool rok = row < a.B;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        zv[j] += bias[j];
        y[j] = leaky(zv[j], a.slope);
    }
    if (a.gamma != nullptr) {  // training-mode BatchNorm1d: two-pass batch statistics
        const float invB = 1.f / (float)a.B;
        float s[4], q[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) s[j] = rok ? y[j] : 0.f;
        col_reduce4(s, S.red, warp, lane);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            s[j] *= invB;  // mean
            const float d = rok ? y[j] - s[j] : 0.f;
            q[j] = d * d;
        }
        col_reduce4(q, S.red, warp, lane);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float var = q[j] * invB;
            const float rstd = rsqrtf(var + a.eps);
            if (stat_writer && cok[j]) {  // (the old running statistics were requested before the product)
                if (a.save_mean) a.save_mean[col[j]] = s[j], a.save_rstd[col[j]] = rstd;
                if (a.running_mean) a.running_mean[col[j]] = fmaf(a.momentum, s[j] - rmean[j], rmean[j]);
                if (a.running_var) {
                    const float unbiased = a.B > 1 ? var * ((float)a.B / (float)(a.B - 1)) : var;
                    a.running_var[col[j]] = fmaf(a.momentum, unbiased - rvar[j], rvar[j]);
                }
            }
            y[j] = fmaf(y[j] - s[j], gam[j] * rstd, bet[j]);
        }
    }
    const uint32_t threshold = drop ? (uint32_t)fminf(a.p * 4294967296.f, 4294967295.f) : 0u;
    const float keep_scale = drop ? 1.f / (1.f - a.p) : 1.f;
    if (rok) {
#pragma unroll
        for (int pr = 0; pr < 2; ++pr) {  // column pairs (2 t, 2 t + 1) of the two 8-column tiles
            bool kf[2] = {true, true};
            if (drop) keep_flag2(seed, call, row, col[2 * pr], threshold, kf[0], kf[1]);
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int j = 2 * pr + e;
                if (!cok[j]) continue;
                const long long off = (long long)row * a.H + col[j];
                if (drop && a.keep) a.keep[off] = kf[e] ? 1 : 0;
                if (a.z) a.z[off] = zv[j];
                a.out[off] = kf[e] ? y[j] * keep_scale : 0.f;
            }
        }
    }
    // the last block to finish advances the call counter (every block has read it by then)
    if (tid == 0) {
        if (blockIdx.x == 0 && blockIdx.y == 0 && a.num_batches) a.num_batches[0] += 1;
        if (drop) {
            __threadfence();
            const unsigned long long tk = atomicAdd(a.rng + 2, 1ull);
            if (tk == (unsigned long long)gridDim.x * gridDim.y - 1) a.rng[2] = 0ull, a.rng[1] = call + 1ull;
        }
    }
}

struct HeadFwd {
    const float* in;  // [B, H]
    long long ld_in;
    const float *Wloc, *bloc, *Wstd, *bstd;  // [O, H], [O]
    float* loc;                              // [B, O]
    float* std;                              // [B, O]  softplus(.) + EPS
    float* spre;                             // [B, O]  pre-activation of the std head (NULL: not kept)
    int B, H, O;
    int bulk;
    // Wstd == NULL: a single plain head (the discriminator's `pred`)
};

__global__ void __launch_bounds__(MF_THREADS, 1) mlp_heads_fwd_kernel(const HeadFwd a) {
    extern __shared__ __align__(16) float smem[];
    const int k = round8(a.H), ld = pitch4(k), k8 = k >> 3;
    const FwdSmem S = carve_fwd(smem, ld);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int wr = warp & 7, h = warp >> 3;
    const int c0 = blockIdx.x * MF_COLS;
    const int row_base = blockIdx.y * MF_ROWS;
    const int n_rows = min(MF_ROWS, a.B - row_base);
    const int n_cols = a.Wstd ? 2 * a.O : a.O;
    init_bar(S.bar);
    stage_rows(S.As, ld, a.H, k, n_rows, [&](int r) { return a.in + (long long)(row_base + r) * a.ld_in; }, S.bar,
               false, a.bulk != 0);
    stage_split(S.Bb, S.Bs, ld, a.H, k, [&](int r) -> const float* {
        const int c = c0 + r;
        if (c < a.O) return a.Wloc + (long long)c * a.H;
        if (c < n_cols) return a.Wstd + (long long)(c - a.O) * a.H;
        return nullptr;
    });
    tc::mbar_wait(S.bar, 0);
    __syncthreads();
    float acc[2][4];
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[nt][e] = 0.f;
    const int s_mid = (k8 + 1) >> 1;
    gemm_mma(S.As, S.Bb, S.Bs, ld, h ? s_mid : 0, h ? k8 : s_mid, wr, lane, acc);
    float v[4];
    combine_halves(acc, S.xch, warp, lane, v);
    const int row = row_base + 16 * wr + g + 8 * h;
    if (row >= a.B) return;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int c = c0 + 8 * (j >> 1) + 2 * t + (j & 1);
        if (c >= n_cols) continue;
        const bool is_loc = c < a.O;
        const int oc = is_loc ? c : c - a.O;
        const float val = v[j] + (is_loc ? a.bloc[oc] : a.bstd[oc]);
        const long long off = (long long)row * a.O + oc;
        if (is_loc) {
            a.loc[off] = val;
        } else {
            a.std[off] = softplusf(val) + GLUE_EPS;
            if (a.spre) a.spre[off] = val;
        }
    }
}

// ------------------------------------------------------------------------------------ backward
// weight-gradient tile of this CTA: rows c0 + {g, g + 8} of dW, columns of the n-tiles this warp owns
template <typename RowDst>
__device__ __forceinline__ void store_wgrad(const float (&c)[2][4], int n_cols, int warp, int lane, RowDst row_dst) {
    const int g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        const int n = 8 * (warp + 16 * q) + 2 * t;
        if (n >= n_cols) continue;  // (n_cols is even: both columns of the pair exist)
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            float* dst = row_dst(g + 8 * i);
            if (dst) *reinterpret_cast<float2*>(dst + n) = make_float2(c[q][2 * i], c[q][2 * i + 1]);
        }
    }
}

struct HeadBwd {
    const float* h;  // [B, H] input of the heads
    long long ld_h;
    const float* d_loc;  // [B, O]
    const float* d_std;  // [B, O]  (NULL: a single plain head)
    const float* spre;   // [B, O]
    float* G;            // [B, n_heads O] = [d loc | d std sigmoid(spre)]
    float *dWloc, *dbloc, *dWstd, *dbstd;
    int B, H, O;
    int bulk;
};

__global__ void __launch_bounds__(MF_THREADS, 1) mlp_heads_bwd_kernel(const HeadBwd a) {
    extern __shared__ __align__(16) float smem[];
    const int k = round8(a.H), ld = pitch8(k);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem);
    float* In = smem + 4;
    float* Gs = In + MF_ROWS * ld;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int c0 = blockIdx.x * MF_COLS;
    const int n_cols = a.d_std ? 2 * a.O : a.O;
    init_bar(bar);
    float acc[2][4];
#pragma unroll
    for (int q = 0; q < 2; ++q)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[q][e] = 0.f;
    float db = 0.f;
    const int n_tiles = (a.B + MF_ROWS - 1) / MF_ROWS;
    for (int rt = 0; rt < n_tiles; ++rt) {  // (row tiles: the weight gradients accumulate in registers)
        const int row_base = rt * MF_ROWS;
        // rows left, NOT clamped to the tile: with min(MF_ROWS, .) inside this loop nvcc 12.9 drops the whole
        // `r < n_rows` branch of stage_rows, bulk copies included (tests/test_abi.py checks the SASS for them)
        const int n_rows = a.B - row_base;
        if (rt) __syncthreads();  // (the previous tile has been consumed)
        stage_rows(In, ld, a.H, k, n_rows, [&](int r) { return a.h + (long long)(row_base + r) * a.ld_h; }, bar, rt > 0,
                   a.bulk != 0);
#pragma unroll
        for (int m = 0; m < MF_ROWS * MF_COLS / MF_THREADS; ++m) {
            const int idx = tid + m * MF_THREADS;
            const int r = idx >> 4, cl = idx & 15, c = c0 + cl, row = row_base + r;
            float v = 0.f;
            if (r < n_rows && c < n_cols) {
                if (c < a.O) {
                    v = a.d_loc[(long long)row * a.O + c];
                } else {
                    const long long off = (long long)row * a.O + c - a.O;
                    const float pre = a.spre[off];
                    v = a.d_std[off] * (pre > 20.f ? 1.f : sigmoidf(pre));  // (torch's softplus threshold)
                }
                a.G[(long long)row * n_cols + c] = v;
            }
            Gs[r * MF_GS_LD + cl] = v;
        }
        tc::mbar_wait(bar, rt & 1);
        __syncthreads();
        wgrad_mma(Gs, In, ld, k >> 3, warp, lane, acc);
        if (tid < MF_COLS) {
            float s = 0.f;
            for (int r = 0; r < MF_ROWS; ++r) s += Gs[r * MF_GS_LD + tid];
            db += s;
        }
    }
    store_wgrad(acc, a.H, warp, lane, [&](int o) -> float* {
        const int c = c0 + o;
        if (c < a.O) return a.dWloc + (long long)c * a.H;
        if (c < n_cols) return a.dWstd + (long long)(c - a.O) * a.H;
        return nullptr;
    });
    if (tid < MF_COLS && c0 + tid < n_cols) {
        const int c = c0 + tid;
        if (c < a.O) a.dbloc[c] = db; else a.dbstd[c - a.O] = db;
    }
}

struct HidBwd {
    const float* Gab;  // [B, Kab] gradient reaching the output of the layer above (pitch Kab)
    int Kab;
    const float* Wa0;  // weights of the layer above: rows [0, n0) of [Kab, H] ...
    const float* Wa1;  // ... and rows [n0, Kab) (the two heads; NULL when n0 == Kab)
    int n0;
    const float* z;  // [B, H] Linear output of this block
    const float* gamma;  // NULL: no BatchNorm
    const float* save_mean;
    const float* save_rstd;
    const unsigned char* keep;  // NULL: no dropout
    const float* in;            // [B, Din] input of this block's Linear
    long long ld_in;
    int Din;
    float* dz;  // [B, H] (NULL: nobody below needs it)
    float* dW;  // [H, Din]
    float* db;  // [H]
    float* dgamma;
    float* dbeta;
    int B, H;
    float slope, p;
    int bulk_g, bulk_in;  // rows of Gab / of `in` can be fetched with bulk copies
};

__global__ void __launch_bounds__(MF_THREADS, 1) mlp_hidden_bwd_kernel(const HidBwd a) {
    extern __shared__ __align__(16) float smem[];
    const int ka = round8(a.Kab), lda = pitch4(ka), ki = round8(a.Din), ldi = pitch8(ki), k8 = ka >> 3;
    const int a_floats = MF_ROWS * (lda > ldi ? lda : ldi);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem);
    float* As = smem + 4;  // G_above, later this block's input
    uint32_t* Bb = reinterpret_cast<uint32_t*>(As + a_floats);  // W_above[:, columns of this CTA], transposed, split
    uint32_t* Bs = Bb + MF_COLS * lda;
    float* Gs = reinterpret_cast<float*>(Bs + MF_COLS * lda);  // dz slab [128][MF_GS_LD]
    float(*xch)[32][4] = reinterpret_cast<float(*)[32][4]>(Gs + MF_ROWS * MF_GS_LD);
    float(*red)[MF_COLS] = reinterpret_cast<float(*)[MF_COLS]>(xch + MF_THREADS / 32);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int wr = warp & 7, h = warp >> 3;
    const int c0 = blockIdx.x * MF_COLS;
    init_bar(bar);
    // Bt[o][j] = W_above[j][c0 + o] (64-byte runs along o), zero beyond Kab; loads first, then split + store
    {
        constexpr int PER = MF_KMAX * MF_COLS / MF_THREADS;  // 8
        float w[PER];
#pragma unroll
        for (int m = 0; m < PER; ++m) {
            const int idx = tid + m * MF_THREADS, j = idx >> 4, o = idx & 15;
            w[m] = 0.f;
            if (j < a.Kab && c0 + o < a.H) {
                const float* wrow = j < a.n0 ? a.Wa0 + (long long)j * a.H : a.Wa1 + (long long)(j - a.n0) * a.H;
                w[m] = __ldg(wrow + c0 + o);
            }
        }
#pragma unroll
        for (int m = 0; m < PER; ++m) {
            const int idx = tid + m * MF_THREADS, j = idx >> 4, o = idx & 15;
            if (j >= ka) continue;
            uint32_t b, s;
            split_tf32(w[m], b, s);
            Bb[o * lda + j] = b, Bs[o * lda + j] = s;
        }
    }
    bool cok[4];
    int col[4];
    float mean[4], rstd[4], gam[4], db_acc[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        col[j] = c0 + 8 * (j >> 1) + 2 * t + (j & 1);
        cok[j] = col[j] < a.H;
        const bool bn = cok[j] && a.gamma != nullptr;
        mean[j] = bn ? a.save_mean[col[j]] : 0.f;
        rstd[j] = bn ? a.save_rstd[col[j]] : 0.f;
        gam[j] = bn ? a.gamma[col[j]] : 0.f;
        db_acc[j] = 0.f;
    }
    float wacc[2][4];
#pragma unroll
    for (int q = 0; q < 2; ++q)
#pragma unroll
        for (int e = 0; e < 4; ++e) wacc[q][e] = 0.f;
    const float keep_scale = a.keep ? 1.f / (1.f - a.p) : 1.f;
    const int s_mid = (k8 + 1) >> 1;
    const int n_tiles = (a.B + MF_ROWS - 1) / MF_ROWS;  // (more than one only without BatchNorm)
    for (int rt = 0; rt < n_tiles; ++rt) {
        const int row_base = rt * MF_ROWS;
        // rows left, NOT clamped to the tile: with min(MF_ROWS, .) inside this loop nvcc 12.9 drops the whole
        // `r < n_rows` branch of stage_rows, bulk copies included (tests/test_abi.py checks the SASS for them)
        const int n_rows = a.B - row_base;
        if (rt) __syncthreads();  // (the previous tile's weight-gradient product is done with As and Gs)
        stage_rows(As, lda, a.Kab, ka, n_rows, [&](int r) { return a.Gab + (long long)(row_base + r) * a.Kab; }, bar,
                   rt > 0, a.bulk_g != 0);
        // saved activations of my 4 elements (requested before the product is waited for)
        const int row = row_base + 16 * wr + g + 8 * h;
        const bool rok = row < a.B;
        float zv[4];
        unsigned char kp[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const bool ok = rok && cok[j];
            const long long off = (long long)row * a.H + col[j];
            zv[j] = ok ? a.z[off] : 0.f;
            kp[j] = (ok && a.keep) ? a.keep[off] : (unsigned char)1;
        }
        tc::mbar_wait(bar, 0);  // (two phases per tile: parities 0, 1)
        __syncthreads();
        float acc[2][4];
#pragma unroll
        for (int nt = 0; nt < 2; ++nt)
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[nt][e] = 0.f;
        gemm_mma(As, Bb, Bs, lda, h ? s_mid : 0, h ? k8 : s_mid, wr, lane, acc);  // d L / d (block output)
        float gout[4], dzv[4];
        combine_halves(acc, xch, warp, lane, gout);
#pragma unroll
        for (int j = 0; j < 4; ++j) gout[j] = (rok && cok[j] && kp[j]) ? gout[j] * keep_scale : 0.f;
        if (a.gamma != nullptr) {
            float sg[4], sgx[4], xhat[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                xhat[j] = (leaky(zv[j], a.slope) - mean[j]) * rstd[j];
                sg[j] = gout[j];
                sgx[j] = gout[j] * xhat[j];
            }
            col_reduce4(sg, red, warp, lane);
            col_reduce4(sgx, red, warp, lane);
            const float invB = 1.f / (float)a.B;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (warp == 0 && g == 0 && cok[j]) {
                    if (a.dgamma) a.dgamma[col[j]] = sgx[j];
                    if (a.dbeta) a.dbeta[col[j]] = sg[j];
                }
                const float ga = gam[j] * rstd[j] * (gout[j] - invB * (sg[j] + xhat[j] * sgx[j]));
                dzv[j] = (rok && cok[j]) ? (zv[j] > 0.f ? ga : a.slope * ga) : 0.f;
            }
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) dzv[j] = zv[j] > 0.f ? gout[j] : a.slope * gout[j];
        }
        float sdz[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            sdz[j] = dzv[j];
            Gs[(16 * wr + g + 8 * h) * MF_GS_LD + 8 * (j >> 1) + 2 * t + (j & 1)] = dzv[j];
            if (a.dz && rok && cok[j]) a.dz[(long long)row * a.H + col[j]] = dzv[j];
        }
        col_reduce4(sdz, red, warp, lane);  // (its barriers also separate the product above from the restaging below)
#pragma unroll
        for (int j = 0; j < 4; ++j) db_acc[j] += sdz[j];
        // weight gradient: dW[c0 + o][i] += sum_r dz[r][o] in[r][i]
        stage_rows(As, ldi, a.Din, ki, n_rows, [&](int r) { return a.in + (long long)(row_base + r) * a.ld_in; }, bar, true,
                   a.bulk_in != 0);
        tc::mbar_wait(bar, 1);
        __syncthreads();
        wgrad_mma(Gs, As, ldi, ki >> 3, warp, lane, wacc);
    }
    if (warp == 0 && g == 0) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (cok[j] && a.db) a.db[col[j]] = db_acc[j];
    }
    store_wgrad(wacc, a.Din, warp, lane,
                [&](int o) -> float* { return c0 + o < a.H ? a.dW + (long long)(c0 + o) * a.Din : nullptr; });
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
inline bool dim_ok(int k) { return k > 0 && k <= MF_KMAX && (k & 1) == 0; }
// rows of a [*, k] fp32 matrix with pitch ld can be fetched with cp.async.bulk
inline int bulk_ok(const void* p, int64_t ld, int k) { return (aligned16(p) && (ld & 3) == 0 && (k & 3) == 0) ? 1 : 0; }
constexpr int64_t MF_MAX_ROWS = 8 * MF_ROWS;  // row tiles (only without BatchNorm)

}  // namespace
}  // namespace glue

using namespace glue;

extern "C" int glue_mlp_hidden_fwd(const float* in, int64_t ld_in, const float* W, const float* bias,
                                   const float* gamma, const float* beta, float* running_mean, float* running_var,
                                   int64_t* num_batches_tracked, uint64_t* rng_state, int64_t B, int32_t Din, int32_t H,
                                   float negative_slope, float eps, float momentum, float p_drop, float* z, float* out,
                                   float* save_mean, float* save_rstd, uint8_t* keep_mask, void* stream) {
    if (!in || !W || !bias || !out || B <= 0 || H <= 0) return GLUE_E_BADARG;
    if (gamma && (!beta)) return GLUE_E_BADARG;
    if (p_drop < 0.f || p_drop >= 1.f || (p_drop > 0.f && !rng_state)) return GLUE_E_BADARG;
    if (B > (gamma ? (int64_t)MF_ROWS : MF_MAX_ROWS) || !dim_ok(Din)) return GLUE_E_UNSUPPORTED;
    HidFwd a;
    a.in = in, a.ld_in = ld_in, a.W = W, a.bias = bias, a.gamma = gamma, a.beta = beta;
    a.running_mean = running_mean, a.running_var = running_var, a.num_batches = (long long*)num_batches_tracked;
    a.rng = p_drop > 0.f ? (unsigned long long*)rng_state : nullptr;
    a.z = z, a.out = out, a.save_mean = save_mean, a.save_rstd = save_rstd, a.keep = p_drop > 0.f ? keep_mask : nullptr;
    a.B = (int)B, a.Din = Din, a.H = H, a.slope = negative_slope, a.eps = eps, a.momentum = momentum, a.p = p_drop;
    a.bulk = bulk_ok(in, ld_in, Din);
    const int ld = pitch4(round8(Din));
    const size_t smem = ((size_t)(MF_ROWS + 2 * MF_COLS) * ld + MF_AUX_FLOATS) * sizeof(float);
    GLUE_SMEM_ONCE(mlp_hidden_fwd_kernel, 227 * 1024);
    const dim3 grid((H + MF_COLS - 1) / MF_COLS, (unsigned)((B + MF_ROWS - 1) / MF_ROWS));
    mlp_hidden_fwd_kernel<<<grid, MF_THREADS, smem, (cudaStream_t)stream>>>(a);
    GLUE_CHECK_LAUNCH();
    return 0;
}

extern "C" int glue_mlp_heads_fwd(const float* in, int64_t ld_in, const float* Wloc, const float* bloc,
                                  const float* Wstd, const float* bstd, int64_t B, int32_t H, int32_t O, float* loc,
                                  float* std, float* spre, void* stream) {
    if (!in || !Wloc || !bloc || !loc || B <= 0 || O <= 0) return GLUE_E_BADARG;
    if (Wstd && (!bstd || !std)) return GLUE_E_BADARG;
    if (B > MF_MAX_ROWS || !dim_ok(H)) return GLUE_E_UNSUPPORTED;
    HeadFwd a;
    a.in = in, a.ld_in = ld_in, a.Wloc = Wloc, a.bloc = bloc, a.Wstd = Wstd, a.bstd = bstd;
    a.loc = loc, a.std = std, a.spre = spre, a.B = (int)B, a.H = H, a.O = O;
    a.bulk = bulk_ok(in, ld_in, H);
    const int ld = pitch4(round8(H));
    const size_t smem = ((size_t)(MF_ROWS + 2 * MF_COLS) * ld + MF_AUX_FLOATS) * sizeof(float);
    GLUE_SMEM_ONCE(mlp_heads_fwd_kernel, 227 * 1024);
    const int n_cols = Wstd ? 2 * O : O;
    const dim3 grid((n_cols + MF_COLS - 1) / MF_COLS, (unsigned)((B + MF_ROWS - 1) / MF_ROWS));
    mlp_heads_fwd_kernel<<<grid, MF_THREADS, smem, (cudaStream_t)stream>>>(a);
    GLUE_CHECK_LAUNCH();
    return 0;
}

extern "C" int glue_mlp_heads_bwd(const float* h, int64_t ld_h, const float* d_loc, const float* d_std,
                                  const float* spre, int64_t B, int32_t H, int32_t O, float* G, float* dWloc,
                                  float* dbloc, float* dWstd, float* dbstd, void* stream) {
    if (!h || !d_loc || !G || !dWloc || !dbloc || B <= 0 || O <= 0) return GLUE_E_BADARG;
    if (d_std && (!spre || !dWstd || !dbstd)) return GLUE_E_BADARG;
    if (B > MF_MAX_ROWS || !dim_ok(H) || ((uintptr_t)dWloc & 7) || ((uintptr_t)dWstd & 7)) return GLUE_E_UNSUPPORTED;
    HeadBwd a;
    a.h = h, a.ld_h = ld_h, a.d_loc = d_loc, a.d_std = d_std, a.spre = spre, a.G = G;
    a.dWloc = dWloc, a.dbloc = dbloc, a.dWstd = dWstd, a.dbstd = dbstd, a.B = (int)B, a.H = H, a.O = O;
    a.bulk = bulk_ok(h, ld_h, H);
    const size_t smem = (4 + (size_t)MF_ROWS * pitch8(round8(H)) + (size_t)MF_ROWS * MF_GS_LD) * sizeof(float);
    GLUE_SMEM_ONCE(mlp_heads_bwd_kernel, 227 * 1024);
    const int n_cols = d_std ? 2 * O : O;
    mlp_heads_bwd_kernel<<<(n_cols + MF_COLS - 1) / MF_COLS, MF_THREADS, smem, (cudaStream_t)stream>>>(a);
    GLUE_CHECK_LAUNCH();
    return 0;
}

extern "C" int glue_mlp_hidden_bwd(const float* g_above, int32_t k_above, const float* w_above0,
                                   const float* w_above1, int32_t n0, const float* z, const float* gamma,
                                   const float* save_mean, const float* save_rstd, const uint8_t* keep_mask,
                                   const float* in, int64_t ld_in, int32_t Din, int64_t B, int32_t H,
                                   float negative_slope, float p_drop, float* dz, float* dW, float* db, float* dgamma,
                                   float* dbeta, void* stream) {
    if (!g_above || !w_above0 || !z || !in || !dW || B <= 0 || H <= 0 || n0 <= 0 || n0 > k_above) return GLUE_E_BADARG;
    if (n0 < k_above && !w_above1) return GLUE_E_BADARG;
    if (gamma && (!save_mean || !save_rstd)) return GLUE_E_BADARG;
    if (p_drop < 0.f || p_drop >= 1.f || (p_drop > 0.f && !keep_mask)) return GLUE_E_BADARG;
    if (B > (gamma ? (int64_t)MF_ROWS : MF_MAX_ROWS) || !dim_ok(Din) || k_above <= 0 || k_above > MF_KMAX ||
        ((uintptr_t)dW & 7))
        return GLUE_E_UNSUPPORTED;
    HidBwd a;
    a.Gab = g_above, a.Kab = k_above, a.Wa0 = w_above0, a.Wa1 = w_above1, a.n0 = n0, a.z = z, a.gamma = gamma;
    a.save_mean = save_mean, a.save_rstd = save_rstd, a.keep = p_drop > 0.f ? keep_mask : nullptr;
    a.in = in, a.ld_in = ld_in, a.Din = Din, a.dz = dz, a.dW = dW, a.db = db, a.dgamma = dgamma, a.dbeta = dbeta;
    a.B = (int)B, a.H = H, a.slope = negative_slope, a.p = p_drop;
    a.bulk_g = bulk_ok(g_above, k_above, k_above), a.bulk_in = bulk_ok(in, ld_in, Din);
    const int lda = pitch4(round8(k_above)), ldi = pitch8(round8(Din));
    const size_t smem = ((size_t)MF_ROWS * (lda > ldi ? lda : ldi) + (size_t)2 * MF_COLS * lda +
                         (size_t)MF_ROWS * MF_GS_LD + MF_AUX_FLOATS) * sizeof(float);  // (the 4 barrier words included)
    if (smem > 227 * 1024) return GLUE_E_UNSUPPORTED;
    GLUE_SMEM_ONCE(mlp_hidden_bwd_kernel, 227 * 1024);
    mlp_hidden_bwd_kernel<<<(H + MF_COLS - 1) / MF_COLS, MF_THREADS, smem, (cudaStream_t)stream>>>(a);
    GLUE_CHECK_LAUNCH();
    return 0;
}
