// GraphConv (CSR gather SpMM), device CSR build, and the graph decoder
// (gather-dot + fused BCE, segmented-reduction backward) for sm_100a.
//
// Reference behaviour: scglue/models/nn.py:26-57 (GraphConv),
// scglue/models/sc.py:54-59 (GraphDecoder), scglue/models/scglue.py:331-338
// (pos/neg balanced Bernoulli NLL).
#include <cub/device/device_radix_sort.cuh>

#include <stdlib.h>

#include "common.cuh"

namespace glue {

// ------------------------------------------------------------------ GraphConv
// One warp per CSR row.  The warp first loads up to 32 (col, val) pairs with one
// coalesced request, then broadcasts them by shuffle and gathers the source
// rows with VEC-wide loads (lanes < K/VEC active; 25 lanes x float2 at K = 50,
// whose 200-byte rows are only 8-byte aligned).  Products are rounded before
// the add and summed in CSR (= original edge) order, i.e. exactly the order of
// the reference's CPU scatter_add_, so the result does not depend on the launch
// geometry.
template <int VEC>
struct VecT;
template <>
struct VecT<1> {
    using type = float;
};
template <>
struct VecT<2> {
    using type = float2;
};
template <>
struct VecT<4> {
    using type = float4;
};

template <int VEC>
__device__ __forceinline__ void axpy_rn(float (&acc)[VEC], float w, const typename VecT<VEC>::type& x) {
    const float* xs = reinterpret_cast<const float*>(&x);
#pragma unroll
    for (int q = 0; q < VEC; ++q) acc[q] = __fadd_rn(acc[q], __fmul_rn(xs[q], w));
}

template <int VEC>
__global__ void __launch_bounds__(256)
graphconv_csr_kernel(const int32_t* __restrict__ indptr, const int32_t* __restrict__ col,
                     const float* __restrict__ val, const float* __restrict__ in,
                     float* __restrict__ out, int64_t n_rows, int K) {
    using V = typename VecT<VEC>::type;
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int nvec = K / VEC;
    for (int64_t row = warp0; row < n_rows; row += nwarps) {
        const int beg = indptr[row], end = indptr[row + 1];
        for (int c0 = 0; c0 < nvec; c0 += 32) {
            const int c = c0 + lane;
            const bool active = c < nvec;
            float acc[VEC];
#pragma unroll
            for (int q = 0; q < VEC; ++q) acc[q] = 0.f;
            for (int base = beg; base < end; base += 32) {
                const int n = min(32, end - base);
                int my_col = 0;
                float my_val = 0.f;
                if (lane < n) {
                    my_col = __ldg(col + base + lane);
                    my_val = __ldg(val + base + lane);
                }
                for (int t = 0; t < n; t += 4) {
                    V x[4];
                    float w[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int src = __shfl_sync(FULL, my_col, (t + q) & 31);
                        w[q] = __shfl_sync(FULL, my_val, (t + q) & 31);
                        if (active && t + q < n)
                            x[q] = __ldg(reinterpret_cast<const V*>(in + (int64_t)src * K) + c);
                    }
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        if (active && t + q < n) axpy_rn<VEC>(acc, w[q], x[q]);
                }
            }
            if (active) {
                V r;
                float* rs = reinterpret_cast<float*>(&r);
#pragma unroll
                for (int q = 0; q < VEC; ++q) rs[q] = acc[q];
                reinterpret_cast<V*>(out + row * K)[c] = r;
            }
        }
    }
}

// Sub-warp variant for rows of <= 32 float2 (K <= 64, even): 8 lanes per row, 4 rows per warp.
// The row-at-a-time kernel keeps one row's gathers in flight per warp and is bound by the
// pointer -> column -> source-row load latencies; here every warp has 4 independent rows (and two
// edges per row, up to four) in flight.  Each row is still summed by its own lanes in edge order with the
// products rounded before the add, so results are bit-identical to the warp-per-row kernel.
__global__ void __launch_bounds__(256)
graphconv_csr_sub8_kernel(const int32_t* __restrict__ indptr, const int32_t* __restrict__ col,
                          const float* __restrict__ val, const float* __restrict__ in,
                          float* __restrict__ out, int64_t n_rows, int K) {
    const int lane = threadIdx.x & 31, sub = lane & 7, grp = lane >> 3;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int nvec = K >> 1;  // float2 per row, <= 32
    bool act[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) act[q] = sub + 8 * q < nvec;
    for (int64_t row4 = warp0 * 4; row4 < n_rows; row4 += nwarps * 4) {
        const int64_t row = row4 + grp;
        const bool rok = row < n_rows;
        const int beg = rok ? __ldg(indptr + row) : 0, end = rok ? __ldg(indptr + row + 1) : 0;
        float2 acc[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[q] = make_float2(0.f, 0.f);
        for (int e = beg; e < end; e += 2) {  // two edges (8 gathers per lane) in flight: 4 was slower
            const bool two = e + 1 < end;
            const int c0 = __ldg(col + e), c1 = two ? __ldg(col + e + 1) : 0;
            const float w0 = __ldg(val + e), w1 = two ? __ldg(val + e + 1) : 0.f;
            const float2* s0 = reinterpret_cast<const float2*>(in + (int64_t)c0 * K);
            const float2* s1 = reinterpret_cast<const float2*>(in + (int64_t)c1 * K);
            float2 x0[4], x1[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                if (act[q]) x0[q] = __ldg(s0 + sub + 8 * q);
                if (act[q] && two) x1[q] = __ldg(s1 + sub + 8 * q);
            }
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if (act[q]) {
                    acc[q].x = __fadd_rn(acc[q].x, __fmul_rn(x0[q].x, w0));
                    acc[q].y = __fadd_rn(acc[q].y, __fmul_rn(x0[q].y, w0));
                    if (two) {
                        acc[q].x = __fadd_rn(acc[q].x, __fmul_rn(x1[q].x, w1));
                        acc[q].y = __fadd_rn(acc[q].y, __fmul_rn(x1[q].y, w1));
                    }
                }
        }
        if (rok) {
            float2* dst = reinterpret_cast<float2*>(out + row * K);
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if (act[q]) dst[sub + 8 * q] = acc[q];
        }
    }
}

// ------------------------------------------------------------------ CSR build
__global__ void csr_keys_kernel(const int64_t* __restrict__ key, int32_t* __restrict__ key32,
                                int32_t* __restrict__ ident, int64_t E) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < E) {
        key32[i] = (int32_t)key[i];
        ident[i] = (int32_t)i;
    }
}

__global__ void csr_fill_kernel(const int32_t* __restrict__ sorted_key,
                                const int32_t* __restrict__ perm,
                                const int64_t* __restrict__ other, const float* __restrict__ val,
                                int64_t E, int64_t n_rows, int32_t* __restrict__ indptr,
                                int32_t* __restrict__ col, float* __restrict__ out_val) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i > E) return;
    if (i < E) {
        const int32_t p = perm[i];
        col[i] = (int32_t)other[p];
        out_val[i] = val[p];
    }
    // rows whose first edge sits at sorted position i
    const int64_t prev = i == 0 ? -1 : (int64_t)sorted_key[i - 1];
    const int64_t cur = i == E ? n_rows : (int64_t)sorted_key[i];
    for (int64_t r = prev + 1; r <= cur; ++r) indptr[r] = (int32_t)i;
}

static int bits_for(int64_t n) {
    int bits = 1;
    while (((int64_t)1 << bits) < n && bits < 31) ++bits;
    return bits;
}

static size_t sort_temp_bytes(int64_t n) {
    size_t bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const int32_t*)nullptr, (int32_t*)nullptr,
                                    (const int32_t*)nullptr, (int32_t*)nullptr, (int)n, 0, 31);
    return bytes;
}

// ------------------------------------------------------------- graph decoder
// 8 lanes per edge (4 edges per warp): each lane strides over the float2 chunks
// of the two endpoint rows, 3 shuffles finish the dot product.
constexpr int GD_THREADS = 256;
constexpr int GD_LPE = 8;  // lanes per edge

__device__ __forceinline__ float group8_sum(float x) {
    x += __shfl_xor_sync(FULL, x, 4);
    x += __shfl_xor_sync(FULL, x, 2);
    x += __shfl_xor_sync(FULL, x, 1);
    return x;
}

__device__ __forceinline__ float pair_dot(const float* __restrict__ a, const float* __restrict__ b,
                                          int K, int sub, bool vec2) {
    float acc = 0.f;
    if (vec2) {
        const float2* a2 = reinterpret_cast<const float2*>(a);
        const float2* b2 = reinterpret_cast<const float2*>(b);
        for (int c = sub; c < (K >> 1); c += GD_LPE) {
            const float2 p = __ldg(a2 + c), q = __ldg(b2 + c);
            acc = fmaf(p.x, q.x, acc);
            acc = fmaf(p.y, q.y, acc);
        }
    } else {
        for (int c = sub; c < K; c += GD_LPE) acc = fmaf(__ldg(a + c), __ldg(b + c), acc);
    }
    return group8_sum(acc);
}

// partial layout per block: {sum_nll_neg, sum_nll_pos, n_neg, n_pos}
__global__ void __launch_bounds__(GD_THREADS)
graphdec_fwd_kernel(const float* __restrict__ v, const int64_t* __restrict__ eidx,
                    const float* __restrict__ esgn, const float* __restrict__ ewt, int64_t E, int K,
                    bool vec2, float* __restrict__ out_logits, float* __restrict__ out_nll,
                    float* __restrict__ out_coef, float* __restrict__ partial,
                    unsigned int* __restrict__ ticket, float* __restrict__ scales,
                    float* __restrict__ out_loss) {
    const int sub = threadIdx.x & (GD_LPE - 1);
    const int64_t grp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / GD_LPE;
    const int64_t ngrp = ((int64_t)gridDim.x * blockDim.x) / GD_LPE;
    float s_neg = 0.f, s_pos = 0.f, n_neg = 0.f, n_pos = 0.f;
    // every lane of a warp runs the same number of iterations (shuffles need the full warp)
    const int64_t iters = (E + ngrp - 1) / ngrp;
    for (int64_t it = 0; it < iters; ++it) {
        const int64_t e = grp0 + it * ngrp;
        const bool live = e < E;
        const int64_t s = live ? eidx[e] : 0, t = live ? eidx[E + e] : 0;
        const float dot = pair_dot(v + s * K, v + t * K, K, sub, vec2);
        if (live && sub == 0) {
            const float sg = esgn[e], w = ewt[e];
            const float logit = sg * dot;
            // Bernoulli(logits).log_prob(w) = -BCEWithLogits(logit, w)
            const float nll = fmaxf(logit, 0.f) - logit * w + log1pf(expf(-fabsf(logit)));
            if (out_logits) out_logits[e] = logit;
            if (out_nll) out_nll[e] = nll;
            if (out_coef) out_coef[e] = sigmoidf(logit) - w;  // scaled by class weight later
            if (w != 0.f) {
                s_pos += nll;
                n_pos += 1.f;
            } else {
                s_neg += nll;
                n_neg += 1.f;
            }
        }
    }
    __shared__ float red[GD_THREADS / 32][4];
    s_neg = warp_sum(s_neg), s_pos = warp_sum(s_pos);
    n_neg = warp_sum(n_neg), n_pos = warp_sum(n_pos);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) {
        red[wid][0] = s_neg, red[wid][1] = s_pos, red[wid][2] = n_neg, red[wid][3] = n_pos;
    }
    __syncthreads();
    if (threadIdx.x < 4) {
        float acc = 0.f;
        for (int w = 0; w < GD_THREADS / 32; ++w) acc += red[w][threadIdx.x];
        partial[(int64_t)blockIdx.x * 4 + threadIdx.x] = acc;
    }
    if (ticket == nullptr) return;
    // fused finalize: the last CTA to arrive reduces the per-block partials in a fixed order (thread <->
    // component t & 3 and one of 64 strided sub-sums, shuffle tree per component, then the 8 warps in
    // order) and publishes the balanced loss of scglue.py:332-338 and the two class weights
    __shared__ bool last;
    __shared__ double wsum[GD_THREADS / 32][4];
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!last) return;
    __threadfence();
    const int c = threadIdx.x & 3;
    double a = 0.0;
    for (int b = threadIdx.x >> 2; b < (int)gridDim.x; b += GD_THREADS / 4) a += (double)__ldcg(partial + (int64_t)b * 4 + c);
#pragma unroll
    for (int o = 16; o >= 4; o >>= 1) a += __shfl_xor_sync(FULL, a, o);  // lanes with equal (t & 3)
    if (lane < 4) wsum[wid][c] = a;
    __syncthreads();
    if (threadIdx.x == 0) {
        double acc[4] = {0.0, 0.0, 0.0, 0.0};
        for (int w = 0; w < GD_THREADS / 32; ++w)
            for (int q = 0; q < 4; ++q) acc[q] += wsum[w][q];
        const double n_neg = acc[2], n_pos = acc[3];
        const double avgc = (n_pos > 0) + (n_neg > 0);
        const double inv_neg = 1.0 / (fmax(n_neg, 1.0) * avgc), inv_pos = 1.0 / (fmax(n_pos, 1.0) * avgc);
        if (out_loss) out_loss[0] = (float)(acc[0] * inv_neg + acc[1] * inv_pos);
        scales[0] = (float)inv_neg;
        scales[1] = (float)inv_pos;
    }
}

// One block: fixed-order reduction of the per-block partials, the balanced loss
// of scglue.py:332-338 (without its .item() host sync), then class-weights the
// per-edge coefficients in place.
__global__ void __launch_bounds__(1024)
graphdec_finalize_kernel(const float* __restrict__ partial, int nblocks,
                         const float* __restrict__ ewt, int64_t E, float* __restrict__ out_loss,
                         float* __restrict__ out_coef) {
    // every block reduces the (few KB of) per-block partials itself, in a fixed order:
    // thread <-> (component c = t & 3, one of 256 strided sub-sums), then a shuffle tree per
    // component and a fixed-order sum over the 8 warps that hold it.
    __shared__ double wsum[32][4];
    __shared__ float scale[2];
    const int t = threadIdx.x, c = t & 3, sub = t >> 2;
    double a = 0.0;
    for (int b = sub; b < nblocks; b += 256) a += (double)partial[(int64_t)b * 4 + c];
#pragma unroll
    for (int o = 16; o >= 4; o >>= 1) a += __shfl_xor_sync(FULL, a, o);  // lanes with equal (t & 3)
    if ((t & 31) < 4) wsum[t >> 5][c] = a;
    __syncthreads();
    if (t == 0) {
        double acc[4] = {0.0, 0.0, 0.0, 0.0};
        for (int w = 0; w < 32; ++w)
            for (int q = 0; q < 4; ++q) acc[q] += wsum[w][q];
        const double n_neg = acc[2], n_pos = acc[3];
        const double avgc = (n_pos > 0) + (n_neg > 0);
        const double inv_neg = 1.0 / (fmax(n_neg, 1.0) * avgc), inv_pos = 1.0 / (fmax(n_pos, 1.0) * avgc);
        if (out_loss && blockIdx.x == 0) out_loss[0] = (float)(acc[0] * inv_neg + acc[1] * inv_pos);
        scale[0] = (float)inv_neg;
        scale[1] = (float)inv_pos;
    }
    __syncthreads();
    if (out_coef)
        for (int64_t e = (int64_t)blockIdx.x * blockDim.x + t; e < E; e += (int64_t)gridDim.x * blockDim.x)
            out_coef[e] *= scale[ewt[e] != 0.f ? 1 : 0];
}

__global__ void graphdec_incidence_kernel(const int64_t* __restrict__ eidx, int64_t E,
                                          int32_t* __restrict__ key, int32_t* __restrict__ entry) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < 2 * E) {
        key[i] = (int32_t)eidx[i];  // eidx is [2, E] row-major: i < E sources, i >= E targets
        entry[i] = (int32_t)i;
    }
}

// Segment heads walk their run of the vertex-sorted incidence list; 8 lanes per
// list position so that short segments (the common case) still fill the warp.
__global__ void __launch_bounds__(GD_THREADS)
graphdec_bwd_kernel(const float* __restrict__ v, const int64_t* __restrict__ eidx,
                    const float* __restrict__ esgn, const float* __restrict__ ewt,
                    const float* __restrict__ coef, const float* __restrict__ scales,
                    const float* __restrict__ upstream, const int32_t* __restrict__ key,
                    const int32_t* __restrict__ entry, int64_t E, int K,
                    float* __restrict__ grad_v) {
    const int sub = threadIdx.x & (GD_LPE - 1);
    const int64_t n = 2 * E;
    const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / GD_LPE;
    if (i >= n) return;
    const int32_t vert = key[i];
    if (i > 0 && key[i - 1] == vert) return;  // not a segment head
    const float up = upstream ? upstream[0] : 1.f;
    for (int c0 = 0; c0 < K; c0 += GD_LPE * 8) {
        float acc[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) acc[q] = 0.f;
        for (int64_t p = i; p < n && key[p] == vert; ++p) {
            const int32_t q_ent = entry[p];
            const int64_t e = q_ent < E ? q_ent : q_ent - E;
            const int64_t other = q_ent < E ? eidx[E + e] : eidx[e];
            float cf = coef[e];
            if (scales) cf *= scales[ewt[e] != 0.f ? 1 : 0];  // (class weights of the fused forward)
            const float w = cf * esgn[e] * up;
            const float* row = v + other * K;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int c = c0 + q * GD_LPE + sub;
                if (c < K) acc[q] = fmaf(w, __ldg(row + c), acc[q]);
            }
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int c = c0 + q * GD_LPE + sub;
            if (c < K) grad_v[(int64_t)vert * K + c] = acc[q];
        }
    }
}

// Second generation (K <= 64).  The sorted list is turned into a CSR over vertices right after the sort
// (graphdec_segments_kernel, still in the forward pass: it depends on `eidx` only): segptr[w] .. segptr[w + 1]
// are the list positions of vertex w, meta[p] = (other endpoint, edge) of position p.  The backward kernel
// then runs one 8-lane group per *vertex*: the lanes resolve up to 8 entries side by side (coefficient,
// sign, class weight by ewt[e] != 0 -- coefficients arrive unscaled from the fused forward) and broadcast
// them, the source rows of 2 entries are gathered at a time; every row of grad_v is written by its own
// group (zeros for untouched vertices: no memset).  Sums run in list order, as in the kernel above
// (bit-identical results).
__global__ void graphdec_segments_kernel(const int32_t* __restrict__ key, const int32_t* __restrict__ entry,
                                         const int64_t* __restrict__ eidx, int64_t E, int64_t V,
                                         int32_t* __restrict__ segptr, int2* __restrict__ meta) {
    const int64_t n = 2 * E;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n) return;
    if (i < n) {
        const int32_t ent = entry[i];
        const int64_t e = ent < E ? ent : ent - E;
        meta[i] = make_int2((int32_t)(ent < E ? eidx[E + e] : eidx[e]), (int32_t)e);
    }
    const int64_t prev = i == 0 ? -1 : (int64_t)key[i - 1];
    const int64_t cur = i == n ? V : (int64_t)key[i];
    for (int64_t r = prev + 1; r <= cur; ++r) segptr[r] = (int32_t)i;  // vertices whose run starts at i
}

template <bool VEC2>
__global__ void __launch_bounds__(GD_THREADS, 4)
graphdec_bwd3_kernel(const float* __restrict__ v, const float* __restrict__ esgn, const float* __restrict__ ewt,
                     const float* __restrict__ coef, const float* __restrict__ scales,
                     const float* __restrict__ upstream, const int32_t* __restrict__ segptr,
                     const int2* __restrict__ meta, int64_t V, int K, float* __restrict__ grad_v) {
    constexpr int Q = VEC2 ? 4 : 8;  // columns per lane: float2 q * 8 + sub, or float q * 8 + sub
    constexpr int NJ = 2;            // source rows gathered at a time
    using T = typename VecT<VEC2 ? 2 : 1>::type;
    const int lane = threadIdx.x & 31, sub = lane & 7;
    const unsigned gmask = 0xffu << (lane & 24);
    const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / GD_LPE;
    if (r >= V) return;
    const int beg = __ldg(segptr + r), end = __ldg(segptr + r + 1);
    const int ncol = VEC2 ? (K >> 1) : K;
    bool act[Q];
#pragma unroll
    for (int q = 0; q < Q; ++q) act[q] = q * 8 + sub < ncol;
    T acc[Q];
#pragma unroll
    for (int q = 0; q < Q; ++q) {
        if constexpr (VEC2) acc[q] = make_float2(0.f, 0.f); else acc[q] = 0.f;
    }
    if (beg < end) {
        const float up = upstream ? upstream[0] : 1.f;
        const float sc_neg = scales ? scales[0] : 1.f, sc_pos = scales ? scales[1] : 1.f;
        const T* vbase = reinterpret_cast<const T*>(v) + sub;
        for (int p = beg; p < end; p += 8) {
            const int m = min(8, end - p);
            int other = 0;
            float w = 0.f;
            if (sub < m) {
                const int2 mt = __ldg(meta + p + sub);
                other = mt.x;
                float cf = coef[mt.y];
                if (scales) cf *= (ewt[mt.y] != 0.f ? sc_pos : sc_neg);
                w = cf * esgn[mt.y] * up;
            }
#pragma unroll
            for (int jb = 0; jb < 8; jb += NJ) {
                if (jb < m) {
                    T x[NJ][Q];
                    float wj[NJ];
#pragma unroll
                    for (int j = 0; j < NJ; ++j) {
                        const int src = min(jb + j, m - 1);  // (beyond the run: a valid row, unused)
                        const int oj = __shfl_sync(gmask, other, src, 8);
                        wj[j] = __shfl_sync(gmask, w, src, 8);
                        const T* row = vbase + (int64_t)oj * ncol;
#pragma unroll
                        for (int q = 0; q < Q; ++q)
                            if (act[q]) x[j][q] = __ldg(row + q * 8);
                    }
#pragma unroll
                    for (int j = 0; j < NJ; ++j)
                        if (jb + j < m) {
#pragma unroll
                            for (int q = 0; q < Q; ++q)
                                if (act[q]) {
                                    if constexpr (VEC2) {
                                        acc[q].x = fmaf(wj[j], x[j][q].x, acc[q].x);
                                        acc[q].y = fmaf(wj[j], x[j][q].y, acc[q].y);
                                    } else {
                                        acc[q] = fmaf(wj[j], x[j][q], acc[q]);
                                    }
                                }
                        }
                }
            }
        }
    }
    T* dst = reinterpret_cast<T*>(grad_v + r * K) + sub;
#pragma unroll
    for (int q = 0; q < Q; ++q)
        if (act[q]) dst[q * 8] = acc[q];
}

}  // namespace glue

using namespace glue;

// =============================================================== C ABI
extern "C" int glue_graphconv_csr(const int32_t* indptr, const int32_t* col, const float* val,
                                  const float* in, float* out, int64_t n_rows, int32_t K,
                                  void* stream) {
    if (!indptr || !in || !out || n_rows <= 0 || K <= 0) return GLUE_E_BADARG;
    if (!col || !val) return GLUE_E_BADARG;
    cudaStream_t st = (cudaStream_t)stream;
    const int threads = 256;
    const int64_t want = (n_rows + 7) / 8;
    const int grid = (int)(want < (int64_t)sm_count() * 16 ? want : (int64_t)sm_count() * 16);
    const uintptr_t bits = (uintptr_t)in | (uintptr_t)out;
    static const bool use_sub8 = [] {
        const char* e = getenv("GLUE_GC_SUB8");
        return !(e && e[0] == '0');
    }();
    if (use_sub8 && K % 2 == 0 && K <= 64 && bits % 8 == 0) {
        const int64_t want4 = ((n_rows + 3) / 4 + 7) / 8;  // 4 rows per warp, 8 warps per block
        const int grid4 = (int)(want4 < 1 ? 1 : (want4 < (int64_t)sm_count() * 32 ? want4 : (int64_t)sm_count() * 32));
        graphconv_csr_sub8_kernel<<<grid4, threads, 0, st>>>(indptr, col, val, in, out, n_rows, K);
        GLUE_CHECK_LAUNCH();
        return 0;
    }
    if (K % 4 == 0 && bits % 16 == 0)
        graphconv_csr_kernel<4><<<grid, threads, 0, st>>>(indptr, col, val, in, out, n_rows, K);
    else if (K % 2 == 0 && bits % 8 == 0)
        graphconv_csr_kernel<2><<<grid, threads, 0, st>>>(indptr, col, val, in, out, n_rows, K);
    else
        graphconv_csr_kernel<1><<<grid, threads, 0, st>>>(indptr, col, val, in, out, n_rows, K);
    GLUE_CHECK_LAUNCH();
    return 0;
}

extern "C" size_t glue_csr_build_workspace_bytes(int64_t E, int64_t n_rows) {
    (void)n_rows;
    if (E <= 0) return 256;
    return 4 * align_up((size_t)E * 4, 256) + align_up(sort_temp_bytes(E), 256);
}

extern "C" int glue_csr_build(const int64_t* key, const int64_t* other, const float* val, int64_t E,
                              int64_t n_rows, int32_t* indptr, int32_t* col, float* out_val,
                              void* workspace, size_t workspace_bytes, void* stream) {
    if (!indptr || n_rows <= 0 || E < 0) return GLUE_E_BADARG;
    if (E > 0 && (!key || !other || !val || !col || !out_val || !workspace)) return GLUE_E_BADARG;
    if (E >= ((int64_t)1 << 31) || n_rows >= ((int64_t)1 << 31)) return GLUE_E_UNSUPPORTED;
    if (workspace_bytes < glue_csr_build_workspace_bytes(E, n_rows)) return GLUE_E_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t seg = align_up((size_t)E * 4, 256);
    char* ws = (char*)workspace;
    int32_t* k_in = (int32_t*)ws;
    int32_t* k_out = (int32_t*)(ws + seg);
    int32_t* p_in = (int32_t*)(ws + 2 * seg);
    int32_t* p_out = (int32_t*)(ws + 3 * seg);
    void* temp = ws + 4 * seg;
    size_t temp_bytes = E > 0 ? sort_temp_bytes(E) : 0;
    const int threads = 256;
    if (E > 0) {
        csr_keys_kernel<<<(unsigned)((E + threads - 1) / threads), threads, 0, st>>>(key, k_in, p_in, E);
        GLUE_CHECK_LAUNCH();
        GLUE_CUDA(cub::DeviceRadixSort::SortPairs(temp, temp_bytes, k_in, k_out, p_in, p_out, (int)E,
                                                  0, bits_for(n_rows), st));
    }
    csr_fill_kernel<<<(unsigned)((E + 1 + threads - 1) / threads), threads, 0, st>>>(
        k_out, p_out, other, val, E, n_rows, indptr, col, out_val);
    GLUE_CHECK_LAUNCH();
    return 0;
}

// workspace head: per-block partials of the forward kernel (4 floats each), then 8 words for the ticket
// and the class weights of the fused finalize
constexpr int GD_MAX_BLOCKS = 148 * 16 - 2;
constexpr int GD_TAIL_OFFSET = 148 * 16 * 4 - 8;  // in floats

static int graphdec_grid(int64_t E) {
    const int64_t per_block = GD_THREADS / GD_LPE;
    const int64_t want = (E + per_block - 1) / per_block;
    const int64_t cap = (int64_t)sm_count() * 8;
    return (int)(want < cap ? (want > 0 ? want : 1) : cap);
}

extern "C" size_t glue_graphdec_workspace_bytes(int64_t E, int64_t V) {
    if (E <= 0) return 256;
    if (V < 0) V = 0;
    const size_t partial = align_up((size_t)148 * 16 * 4 * sizeof(float), 256);  // <= 2368 blocks
    return partial + 4 * align_up((size_t)E * 8, 256) + align_up(sort_temp_bytes(2 * E), 256) +
           align_up((size_t)(V + 1) * 4, 256) + align_up((size_t)E * 16, 256);  // segptr, meta
}

extern "C" int glue_graphdec_bce_fwd(const float* v, const int64_t* eidx, const float* esgn,
                                     const float* ewt, int64_t E, int64_t V, int32_t K,
                                     float* out_loss, float* out_logits, float* out_nll,
                                     float* out_coef, void* workspace, size_t workspace_bytes,
                                     void* stream) {
    if (!v || !eidx || !esgn || !ewt || E <= 0 || V <= 0 || K <= 0 || !workspace) return GLUE_E_BADARG;
    if (workspace_bytes < glue_graphdec_workspace_bytes(E, V)) return GLUE_E_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = graphdec_grid(E);
    if (grid > GD_MAX_BLOCKS) return GLUE_E_UNSUPPORTED;
    float* partial = (float*)workspace;
    const bool vec2 = (K % 2 == 0) && ((uintptr_t)v % 8 == 0);
    graphdec_fwd_kernel<<<grid, GD_THREADS, 0, st>>>(v, eidx, esgn, ewt, E, K, vec2, out_logits,
                                                     out_nll, out_coef, partial, nullptr, nullptr, nullptr);
    GLUE_CHECK_LAUNCH();
    const int fin_blocks = out_coef ? (int)((E + 4095) / 4096 < 64 ? (E + 4095) / 4096 : 64) : 1;
    graphdec_finalize_kernel<<<fin_blocks < 1 ? 1 : fin_blocks, 1024, 0, st>>>(partial, grid, ewt, E, out_loss,
                                                                             out_coef);
    GLUE_CHECK_LAUNCH();
    return 0;
}

extern "C" int glue_graphdec_bwd(const float* v, const int64_t* eidx, const float* esgn,
                                 const float* coef, int64_t E, int64_t V, int32_t K, float* grad_v,
                                 void* workspace, size_t workspace_bytes, void* stream);

namespace {
struct GdSortBufs {
    int32_t *k_in, *k_out, *p_in, *p_out;
    void* temp;
    size_t temp_bytes;
    int32_t* segptr;  // [V + 1] list positions of every vertex's run
    int2* meta;       // [2 E] (other endpoint, edge) per list position
};
GdSortBufs gd_sort_bufs(void* workspace, int64_t E, int64_t V) {
    char* ws = (char*)workspace + align_up((size_t)148 * 16 * 4 * sizeof(float), 256);
    const size_t seg = align_up((size_t)E * 8, 256);
    GdSortBufs b;
    b.k_in = (int32_t*)ws, b.k_out = (int32_t*)(ws + seg);
    b.p_in = (int32_t*)(ws + 2 * seg), b.p_out = (int32_t*)(ws + 3 * seg);
    b.temp = ws + 4 * seg, b.temp_bytes = sort_temp_bytes(2 * E);
    char* tail = ws + 4 * seg + align_up(b.temp_bytes, 256);
    b.segptr = (int32_t*)tail;
    b.meta = (int2*)(tail + align_up((size_t)(V + 1) * 4, 256));
    return b;
}
}  // namespace

namespace {
// K <= 64: one group per vertex over the segment CSR; wider rows: the one-entry-at-a-time kernel + memset
int graphdec_bwd_launch(const float* v, const int64_t* eidx, const float* esgn, const float* ewt, const float* coef,
                        const float* scales, const float* upstream, const GdSortBufs& b, int64_t E, int64_t V, int K,
                        float* grad_v, cudaStream_t st) {
    if (K <= 64) {
        const bool vec2 = (K % 2 == 0) && (((uintptr_t)v | (uintptr_t)grad_v) % 8 == 0);
        const unsigned vblocks = (unsigned)((V * GD_LPE + GD_THREADS - 1) / GD_THREADS);
        if (vec2)
            graphdec_bwd3_kernel<true><<<vblocks, GD_THREADS, 0, st>>>(v, esgn, ewt, coef, scales, upstream, b.segptr,
                                                                       b.meta, V, K, grad_v);
        else
            graphdec_bwd3_kernel<false><<<vblocks, GD_THREADS, 0, st>>>(v, esgn, ewt, coef, scales, upstream, b.segptr,
                                                                        b.meta, V, K, grad_v);
        GLUE_CHECK_LAUNCH();
        return 0;
    }
    const int64_t blocks = (2 * E * GD_LPE + GD_THREADS - 1) / GD_THREADS;
    GLUE_CUDA(cudaMemsetAsync(grad_v, 0, (size_t)V * K * sizeof(float), st));
    graphdec_bwd_kernel<<<(unsigned)blocks, GD_THREADS, 0, st>>>(v, eidx, esgn, ewt, coef, scales, upstream, b.k_out,
                                                                 b.p_out, E, K, grad_v);
    GLUE_CHECK_LAUNCH();
    return 0;
}
}  // namespace

// The backward pass sums, per vertex, over its incident edges of the minibatch: (vertex, edge)
// incidences sorted by vertex.  That order depends on `eidx` only, so it can be prepared right after
// the forward kernel (off the critical path of the backward pass) in the same workspace, which then
// has to stay alive until glue_graphdec_bwd_sorted has run.
extern "C" int glue_graphdec_sort_incidence(const int64_t* eidx, int64_t E, int64_t V, void* workspace,
                                            size_t workspace_bytes, void* stream) {
    if (!eidx || E <= 0 || V <= 0 || !workspace) return GLUE_E_BADARG;
    if (2 * E >= ((int64_t)1 << 31) || V >= ((int64_t)1 << 31)) return GLUE_E_UNSUPPORTED;
    if (workspace_bytes < glue_graphdec_workspace_bytes(E, V)) return GLUE_E_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    const GdSortBufs b = gd_sort_bufs(workspace, E, V);
    const int threads = 256;
    graphdec_incidence_kernel<<<(unsigned)((2 * E + threads - 1) / threads), threads, 0, st>>>(eidx, E, b.k_in,
                                                                                               b.p_in);
    GLUE_CHECK_LAUNCH();
    size_t temp_bytes = b.temp_bytes;
    GLUE_CUDA(cub::DeviceRadixSort::SortPairs(b.temp, temp_bytes, b.k_in, b.k_out, b.p_in, b.p_out, (int)(2 * E), 0,
                                              bits_for(V), st));
    graphdec_segments_kernel<<<(unsigned)((2 * E + 1 + threads - 1) / threads), threads, 0, st>>>(
        b.k_out, b.p_out, eidx, E, V, b.segptr, b.meta);
    GLUE_CHECK_LAUNCH();
    return 0;
}

// `upstream` (device scalar, may be NULL) multiplies every coefficient: lets
// autograd hand over d L / d loss without an extra pass over grad_v.
extern "C" int glue_graphdec_bwd_sorted(const float* v, const int64_t* eidx, const float* esgn, const float* coef,
                                        const float* upstream, int64_t E, int64_t V, int32_t K, float* grad_v,
                                        void* workspace, size_t workspace_bytes, void* stream) {
    if (!v || !eidx || !esgn || !coef || !grad_v || E <= 0 || V <= 0 || K <= 0 || !workspace)
        return GLUE_E_BADARG;
    if (2 * E >= ((int64_t)1 << 31) || V >= ((int64_t)1 << 31)) return GLUE_E_UNSUPPORTED;
    if (workspace_bytes < glue_graphdec_workspace_bytes(E, V)) return GLUE_E_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    const GdSortBufs b = gd_sort_bufs(workspace, E, V);
    return graphdec_bwd_launch(v, eidx, esgn, nullptr, coef, nullptr, upstream, b, E, V, K, grad_v, st);
}

// ---- fused-finalize pair (what graph_nll of the training step calls) ---------------------------
// Forward: one launch -- the last CTA to arrive reduces the partials, writes the loss and leaves the two
// class weights in the workspace; out_coef stays *unscaled* (sigma(logit) - ewt).  With prepare_bwd the
// vertex-sorted incidence list follows on the same stream.  Backward: glue_graph_nll_bwd applies the
// class weights while it resolves the entries; it writes every row of grad_v itself.
extern "C" int glue_graph_nll_fwd(const float* v, const int64_t* eidx, const float* esgn, const float* ewt,
                                  int64_t E, int64_t V, int32_t K, float* out_loss, float* out_coef,
                                  int32_t prepare_bwd, void* workspace, size_t workspace_bytes, void* stream) {
    if (!v || !eidx || !esgn || !ewt || !out_loss || E <= 0 || V <= 0 || K <= 0 || !workspace) return GLUE_E_BADARG;
    if (workspace_bytes < glue_graphdec_workspace_bytes(E, V)) return GLUE_E_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = graphdec_grid(E);
    if (grid > GD_MAX_BLOCKS) return GLUE_E_UNSUPPORTED;
    float* partial = (float*)workspace;
    float* tail = partial + GD_TAIL_OFFSET;  // [0]: ticket, [1..2]: class weights
    GLUE_CUDA(cudaMemsetAsync(tail, 0, 4, st));
    const bool vec2 = (K % 2 == 0) && ((uintptr_t)v % 8 == 0);
    graphdec_fwd_kernel<<<grid, GD_THREADS, 0, st>>>(v, eidx, esgn, ewt, E, K, vec2, nullptr, nullptr, out_coef,
                                                     partial, (unsigned int*)tail, tail + 1, out_loss);
    GLUE_CHECK_LAUNCH();
    if (prepare_bwd) return glue_graphdec_sort_incidence(eidx, E, V, workspace, workspace_bytes, stream);
    return 0;
}

extern "C" int glue_graph_nll_bwd(const float* v, const int64_t* eidx, const float* esgn, const float* ewt,
                                  const float* coef, const float* upstream, int64_t E, int64_t V, int32_t K,
                                  float* grad_v, void* workspace, size_t workspace_bytes, void* stream) {
    if (!v || !eidx || !esgn || !ewt || !coef || !grad_v || E <= 0 || V <= 0 || K <= 0 || !workspace)
        return GLUE_E_BADARG;
    if (2 * E >= ((int64_t)1 << 31) || V >= ((int64_t)1 << 31)) return GLUE_E_UNSUPPORTED;
    if (workspace_bytes < glue_graphdec_workspace_bytes(E, V)) return GLUE_E_WORKSPACE;
    const GdSortBufs b = gd_sort_bufs(workspace, E, V);
    const float* scales = (const float*)workspace + GD_TAIL_OFFSET + 1;
    return graphdec_bwd_launch(v, eidx, esgn, ewt, coef, scales, upstream, b, E, V, K, grad_v, (cudaStream_t)stream);
}

extern "C" int glue_graphdec_bwd_scaled(const float* v, const int64_t* eidx, const float* esgn,
                                        const float* coef, const float* upstream, int64_t E,
                                        int64_t V, int32_t K, float* grad_v, void* workspace,
                                        size_t workspace_bytes, void* stream) {
    if (!v || !eidx || !esgn || !coef || !grad_v || E <= 0 || V <= 0 || K <= 0 || !workspace)
        return GLUE_E_BADARG;
    const int rc = glue_graphdec_sort_incidence(eidx, E, V, workspace, workspace_bytes, stream);
    if (rc) return rc;
    return glue_graphdec_bwd_sorted(v, eidx, esgn, coef, upstream, E, V, K, grad_v, workspace, workspace_bytes,
                                    stream);
}

extern "C" int glue_graphdec_bwd(const float* v, const int64_t* eidx, const float* esgn,
                                 const float* coef, int64_t E, int64_t V, int32_t K, float* grad_v,
                                 void* workspace, size_t workspace_bytes, void* stream) {
    return glue_graphdec_bwd_scaled(v, eidx, esgn, coef, nullptr, E, V, K, grad_v, workspace,
                                    workspace_bytes, stream);
}
