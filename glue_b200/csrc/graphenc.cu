// GraphEncoder head, fused (sm_100a).
//
// Replaces, for the whole vertex table in one forward and one backward launch sequence,
//   loc / std_lin Linear layers + softplus        scglue/models/sc.py:44-46
//   the reparameterised sample  v = mu + eps*sigma scglue/models/scglue.py:329
//   KL(N(mu, sigma) || N(0, 1)) summed             scglue/models/scglue.py:339-340
// i.e. ~30 ATen kernels over [V, K] tensors, among them two weight-gradient GEMMs with a
// K x K output and a V-long reduction that cuBLAS runs on a single CTA (154 us each at
// V = 52 000, profiles/r01c_launches_summary.md).
//
//   mu = ptr W_loc^T + b_loc        s = ptr W_std^T + b_std       sigma = softplus(s) + 1e-7
//   v  = mu + noise * sigma         kl = sum_ik ( -log sigma + (sigma^2 + mu^2)/2 - 1/2 )
//
// Tiles of 64 vertices per CTA iteration, 256 threads; thread (row r = t >> 2, lane group
// ig = t & 3) owns the float4 column blocks b = ig, ig + 4, ... of its row.  Weight matrices
// live in shared memory ([out][in], row stride chosen so that the 8 rows of a warp hit
// distinct banks).  Weight gradients are accumulated in registers over all tiles of a CTA
// (thread <-> (out = t >> 2, in = ig + 4 ii)), written as per-CTA partials and reduced in a fixed
// order => bit-reproducible.  Bound: FFMA / shared-memory issue (10 000 FMA per vertex fwd + bwd
// against 1 KB of HBM traffic per vertex).
#include <algorithm>
#include <cstdlib>
#include "common.cuh"

namespace glue {
namespace {

constexpr int TR = 64;          // vertex rows per tile
constexpr int ENC_THREADS = 256;

template <int KP>
struct EncCfg {
    static constexpr int NB4 = KP / 4;                   // float4 blocks per row
    static constexpr int KS = (KP % 32 == 20) ? KP : KP + 4;  // row stride (words): 4-bank skew per row
    static constexpr int NJ = (NB4 + 3) / 4;             // blocks per thread
    // weight rows: stride a multiple of 32 words plus a skew of 4 ((o >> 2) & 3) words, so that
    // the four lane groups of a warp (output rows o, o + 4, o + 8, o + 12) read distinct banks
    static constexpr int WS = (KP + 12 + 31) / 32 * 32;
};

__device__ __forceinline__ int wskew(int o) { return 4 * ((o >> 2) & 3); }

template <int KP>
struct EncSmem {
    float x[TR][EncCfg<KP>::KS];
    float wl[KP][EncCfg<KP>::WS];
    float ws[KP][EncCfg<KP>::WS];
    float d1[TR][EncCfg<KP>::KS];  // backward: d L / d mu
    float d2[TR][EncCfg<KP>::KS];  // backward: d L / d s
    float bl[KP], bs[KP];
    double red[ENC_THREADS / 32];
    unsigned int ticket;
};

struct EncArgs {
    const float *ptr, *w_loc, *b_loc, *w_std, *b_std, *noise;
    float *mu, *sigma, *vsamp, *kl_sum;
    // backward
    const float *grad_vsamp, *kl_coef;
    float *grad_ptr, *grad_w_loc, *grad_b_loc, *grad_w_std, *grad_b_std;
    // workspace
    double* klpart;         // [ncta]
    float* gwpart;          // [ncta][2][KP][KP + 1]  (last column: bias gradient)
    unsigned int* counter;  // [1]
    int64_t V;
    int K;
};

template <int KP>
__device__ __forceinline__ void load_weights(EncSmem<KP>& S, const EncArgs& a) {
    for (int idx = threadIdx.x; idx < KP * KP; idx += ENC_THREADS) {
        const int o = idx / KP, i = idx - o * KP;
        const bool ok = o < a.K && i < a.K;
        S.wl[o][wskew(o) + i] = ok ? a.w_loc[o * a.K + i] : 0.f;
        S.ws[o][wskew(o) + i] = ok ? a.w_std[o * a.K + i] : 0.f;
    }
    for (int o = threadIdx.x; o < KP; o += ENC_THREADS) {
        S.bl[o] = o < a.K ? a.b_loc[o] : 0.f;
        S.bs[o] = o < a.K ? a.b_std[o] : 0.f;
    }
}

template <int KP>
__device__ __forceinline__ void load_x_tile(EncSmem<KP>& S, const EncArgs& a, int64_t row0) {
    constexpr int KS = EncCfg<KP>::KS;
    const int K = a.K;
    for (int idx = threadIdx.x; idx < TR * K; idx += ENC_THREADS) {
        const int r = idx / K, k = idx - r * K;
        const int64_t row = row0 + r;
        S.x[r][k] = row < a.V ? __ldg(a.ptr + row * K + k) : 0.f;
    }
    if (K < KS)
        for (int idx = threadIdx.x; idx < TR * (KS - K); idx += ENC_THREADS) {
            const int r = idx / (KS - K), k = K + idx % (KS - K);
            S.x[r][k] = 0.f;
        }
}

__device__ __forceinline__ float dot4(const float4& x, const float4& w, float acc) {
    acc = fmaf(x.x, w.x, acc);
    acc = fmaf(x.y, w.y, acc);
    acc = fmaf(x.z, w.z, acc);
    return fmaf(x.w, w.w, acc);
}

// ------------------------------------------------------------------------------------ forward
template <int KP>
__global__ void __launch_bounds__(ENC_THREADS) graphenc_fwd_kernel(EncArgs a) {
    extern __shared__ float4 enc_smem_raw[];
    EncSmem<KP>& S = *reinterpret_cast<EncSmem<KP>*>(enc_smem_raw);
    constexpr int NB4 = EncCfg<KP>::NB4, NJ = EncCfg<KP>::NJ;
    const int t = threadIdx.x, r = t >> 2, ig = t & 3;
    const int K = a.K;
    load_weights<KP>(S, a);
    double kl_acc = 0.0;
    const int64_t ntiles = (a.V + TR - 1) / TR;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        __syncthreads();
        load_x_tile<KP>(S, a, tile * TR);
        __syncthreads();
        const int64_t row = tile * TR + r;
        const bool rowok = row < a.V;
        float4 xr[NB4];
#pragma unroll
        for (int i4 = 0; i4 < NB4; ++i4) xr[i4] = *reinterpret_cast<const float4*>(&S.x[r][4 * i4]);
        float kl_row = 0.f;
#pragma unroll
        for (int jj = 0; jj < NJ; ++jj) {
            const int b = ig + 4 * jj;
            if (b >= NB4) continue;
            float mu4[4], s4[4];
#pragma unroll
            for (int oo = 0; oo < 4; ++oo) {
                const int o = 4 * b + oo;
                float accl = S.bl[o], accs = S.bs[o];
#pragma unroll
                for (int i4 = 0; i4 < NB4; ++i4) {
                    accl = dot4(xr[i4], *reinterpret_cast<const float4*>(&S.wl[o][4 * ig + 4 * i4]), accl);
                    accs = dot4(xr[i4], *reinterpret_cast<const float4*>(&S.ws[o][4 * ig + 4 * i4]), accs);
                }
                mu4[oo] = accl, s4[oo] = accs;
            }
#pragma unroll
            for (int o2 = 0; o2 < 2; ++o2) {  // float2 granularity: rows are 8-byte aligned for even K
                const int o = 4 * b + 2 * o2;
                if (!rowok || o >= K) continue;
                const int64_t off = row * K + o;
                float2 nz = make_float2(0.f, 0.f);
                if (a.noise) nz = __ldg(reinterpret_cast<const float2*>(a.noise + off));
                const float m0 = mu4[2 * o2], m1 = mu4[2 * o2 + 1];
                const float g0 = softplusf(s4[2 * o2]) + GLUE_EPS, g1 = softplusf(s4[2 * o2 + 1]) + GLUE_EPS;
                *reinterpret_cast<float2*>(a.mu + off) = make_float2(m0, m1);
                *reinterpret_cast<float2*>(a.sigma + off) = make_float2(g0, g1);
                *reinterpret_cast<float2*>(a.vsamp + off) = make_float2(fmaf(nz.x, g0, m0), fmaf(nz.y, g1, m1));
                kl_row += (-logf(g0) + 0.5f * (g0 * g0 + m0 * m0) - 0.5f) +
                          (-logf(g1) + 0.5f * (g1 * g1 + m1 * m1) - 0.5f);
            }
        }
        kl_acc += (double)kl_row;
    }
    // KL: deterministic block reduction, per-CTA partial, last CTA sums the partials in order
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) kl_acc += __shfl_xor_sync(FULL, kl_acc, o);
    if ((t & 31) == 0) S.red[t >> 5] = kl_acc;
    __syncthreads();
    if (t == 0) {
        double s = 0.0;
        for (int w = 0; w < ENC_THREADS / 32; ++w) s += S.red[w];
        a.klpart[blockIdx.x] = s;
        __threadfence();
        S.ticket = atomicAdd(a.counter, 1u);
    }
    __syncthreads();
    if (S.ticket == gridDim.x - 1 && t == 0) {
        __threadfence();
        double s = 0.0;
        for (unsigned q = 0; q < gridDim.x; ++q) s += __ldcg(a.klpart + q);
        a.kl_sum[0] = (float)s;
        *a.counter = 0u;  // ready for the next call on this workspace
    }
}

// ----------------------------------------------------------------------------------- backward
template <int KP>
__global__ void __launch_bounds__(ENC_THREADS) graphenc_bwd_kernel(EncArgs a) {
    extern __shared__ float4 enc_smem_raw[];
    EncSmem<KP>& S = *reinterpret_cast<EncSmem<KP>*>(enc_smem_raw);
    constexpr int NB4 = EncCfg<KP>::NB4, NJ = EncCfg<KP>::NJ, KS = EncCfg<KP>::KS;
    constexpr int NI = KP / 4;  // inputs per thread in the weight-gradient mapping (i = ig + 4 ii)
    const int t = threadIdx.x, r = t >> 2, ig = t & 3;
    const int K = a.K;
    load_weights<KP>(S, a);
    const float ckl = a.kl_coef ? __ldg(a.kl_coef) : 0.f;
    float gwl[NI], gws[NI], gbl = 0.f, gbs = 0.f;
#pragma unroll
    for (int ii = 0; ii < NI; ++ii) gwl[ii] = gws[ii] = 0.f;
    const int wo = t >> 2;  // output feature owned in the weight-gradient mapping (TR == 64 == max KP)
    const int64_t ntiles = (a.V + TR - 1) / TR;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        __syncthreads();
        load_x_tile<KP>(S, a, tile * TR);
        const int64_t row = tile * TR + r;
        const bool rowok = row < a.V;
        // d L / d mu, d L / d s of this thread's column blocks
#pragma unroll
        for (int jj = 0; jj < NJ; ++jj) {
            const int b = ig + 4 * jj;
            if (b >= NB4) continue;
            float dm[4] = {0.f, 0.f, 0.f, 0.f}, ds[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int o2 = 0; o2 < 2; ++o2) {
                const int o = 4 * b + 2 * o2;
                if (!rowok || o >= K) continue;
                const int64_t off = row * K + o;
                const float2 gv = a.grad_vsamp ? __ldg(reinterpret_cast<const float2*>(a.grad_vsamp + off))
                                               : make_float2(0.f, 0.f);
                const float2 nz = a.noise ? __ldg(reinterpret_cast<const float2*>(a.noise + off))
                                          : make_float2(0.f, 0.f);
                const float2 m = __ldg(reinterpret_cast<const float2*>(a.mu + off));
                const float2 g = __ldg(reinterpret_cast<const float2*>(a.sigma + off));
                dm[2 * o2] = fmaf(ckl, m.x, gv.x);
                dm[2 * o2 + 1] = fmaf(ckl, m.y, gv.y);
                // d sigma / d s = sigmoid(s) = 1 - exp(-softplus(s)),  softplus(s) = sigma - 1e-7
                const float dsg0 = fmaf(gv.x, nz.x, ckl * (g.x - 1.f / g.x));
                const float dsg1 = fmaf(gv.y, nz.y, ckl * (g.y - 1.f / g.y));
                ds[2 * o2] = dsg0 * (1.f - expf(-(g.x - GLUE_EPS)));
                ds[2 * o2 + 1] = dsg1 * (1.f - expf(-(g.y - GLUE_EPS)));
            }
            *reinterpret_cast<float4*>(&S.d1[r][4 * b]) = make_float4(dm[0], dm[1], dm[2], dm[3]);
            *reinterpret_cast<float4*>(&S.d2[r][4 * b]) = make_float4(ds[0], ds[1], ds[2], ds[3]);
        }
        __syncthreads();
        // (1) d L / d ptr[r, i] = sum_o dmu[r, o] W_loc[o, i] + ds[r, o] W_std[o, i]
        {
            float4 acc[NJ];
#pragma unroll
            for (int jj = 0; jj < NJ; ++jj) acc[jj] = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int o = 0; o < K; ++o) {
                const float d1 = S.d1[r][o], d2 = S.d2[r][o];
                const int sk = wskew(o);
#pragma unroll
                for (int jj = 0; jj < NJ; ++jj) {
                    const int b = ig + 4 * jj;
                    if (b >= NB4) continue;
                    const float4 wl = *reinterpret_cast<const float4*>(&S.wl[o][sk + 4 * b]);
                    const float4 ws = *reinterpret_cast<const float4*>(&S.ws[o][sk + 4 * b]);
                    acc[jj].x = fmaf(d1, wl.x, fmaf(d2, ws.x, acc[jj].x));
                    acc[jj].y = fmaf(d1, wl.y, fmaf(d2, ws.y, acc[jj].y));
                    acc[jj].z = fmaf(d1, wl.z, fmaf(d2, ws.z, acc[jj].z));
                    acc[jj].w = fmaf(d1, wl.w, fmaf(d2, ws.w, acc[jj].w));
                }
            }
            if (rowok && a.grad_ptr) {
#pragma unroll
                for (int jj = 0; jj < NJ; ++jj) {
                    const int b = ig + 4 * jj;
                    if (b >= NB4) continue;
                    float* dst = a.grad_ptr + row * K + 4 * b;
                    if (4 * b < K) *reinterpret_cast<float2*>(dst) = make_float2(acc[jj].x, acc[jj].y);
                    if (4 * b + 2 < K) *reinterpret_cast<float2*>(dst + 2) = make_float2(acc[jj].z, acc[jj].w);
                }
            }
        }
        // (2) weight gradients: thread <-> (out = wo, in = ig + 4 ii), summed over the tile's rows
        if (wo < K) {
            for (int rr = 0; rr < TR; ++rr) {
                const float d1 = S.d1[rr][wo], d2 = S.d2[rr][wo];
#pragma unroll
                for (int ii = 0; ii < NI; ++ii) {
                    const float xi = S.x[rr][ig + 4 * ii];
                    gwl[ii] = fmaf(d1, xi, gwl[ii]);
                    gws[ii] = fmaf(d2, xi, gws[ii]);
                }
                if (ig == 0) gbl += d1, gbs += d2;
            }
        }
        (void)KS;
    }
    // per-CTA partials [2][KP][KP + 1]
    if (wo < K) {
        float* base = a.gwpart + (size_t)blockIdx.x * 2 * KP * (KP + 1);
#pragma unroll
        for (int ii = 0; ii < NI; ++ii) {
            const int i = ig + 4 * ii;
            if (i < K) {
                base[wo * (KP + 1) + i] = gwl[ii];
                base[(KP + wo) * (KP + 1) + i] = gws[ii];
            }
        }
        if (ig == 0) {
            base[wo * (KP + 1) + KP] = gbl;
            base[(KP + wo) * (KP + 1) + KP] = gbs;
        }
    }
}

// ---------------------------------------------------------------------- backward, tensor cores
// Same outputs as graphenc_bwd_kernel with the four K x K products per tile on warp-level tensor
// cores: mma.sync m16n8k8 TF32, every fp32 product rebuilt from three TF32 products
// (big*big + big*small + small*big, ~2^-21 relative) so that the 1e-4 parity bar holds.
//   dgrad   g_ptr[64 x K]  = d1[64 x K] W_loc[K x K] + d2[64 x K] W_std[K x K]
//           warp w: rows 16 (w & 3) .., output columns of half (w >> 2)
//   wgrad   gW_m[K x K]   += d_m^T[K x 64] x[64 x K]      (m = loc, std)
//           warp w: matrix (w >> 2), output rows 16 (w & 3) .., all column tiles; accumulators
//           stay in registers over all tiles of the CTA
// Fragment layouts (PTX ISA, m16n8k8 .tf32): g = lane >> 2, t = lane & 3;
//   A[16 x 8] row-major: a0 (g, t) a1 (g + 8, t) a2 (g, t + 4) a3 (g + 8, t + 4)
//   B[8 x 8]  col-major: b0 (t, g) b1 (t + 4, g)
//   C[16 x 8]:           c0 (g, 2t) c1 (g, 2t + 1) c2 (g + 8, 2t) c3 (g + 8, 2t + 1)
// Shared-memory strides: A-pattern reads (row g, col t) are conflict free for strides = 4 mod 32,
// B-pattern reads (row t, col g) for strides = 8 mod 32.
constexpr int MM_KP = 64;    // padded latent dim (K <= 64)
constexpr int MM_DS = 68;    // stride of d1 / d2 (A operand of dgrad; 2-way conflicts as A^T of wgrad)
constexpr int MM_BS = 72;    // stride of x and of the weight matrices (B operands)

constexpr int ENC_BWD_GROUPS = 2;                             // warp groups per CTA, one 64-row tile each per iteration
constexpr int ENC_BWD_THREADS = ENC_BWD_GROUPS * ENC_THREADS;  // 16 warps: with 8 the kernel waited on its loads
struct MmaSmem {
    float x[ENC_BWD_GROUPS][TR][MM_BS];
    float d1[ENC_BWD_GROUPS][TR][MM_DS];
    float d2[ENC_BWD_GROUPS][TR][MM_DS];
    float wl_big[MM_KP][MM_BS], wl_small[MM_KP][MM_BS];  // [o][i], TF32 split of the weights
    float ws_big[MM_KP][MM_BS], ws_small[MM_KP][MM_BS];
};

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
// c += A B with A, B given as (big, small) TF32 pairs: small terms first
__device__ __forceinline__ void mma_3x(float (&c)[4], const uint32_t (&ab)[4], const uint32_t (&as)[4], uint32_t bb0,
                                       uint32_t bb1, uint32_t bs0, uint32_t bs1) {
    mma_tf32(c, as, bb0, bb1);
    mma_tf32(c, ab, bs0, bs1);
    mma_tf32(c, ab, bb0, bb1);
}

__global__ void __launch_bounds__(ENC_BWD_THREADS) graphenc_bwd_mma_kernel(EncArgs a, int KPW /* 52 or 64: partial layout */) {
    extern __shared__ float4 enc_smem_raw[];
    MmaSmem& S = *reinterpret_cast<MmaSmem*>(enc_smem_raw);
    // two groups of 8 warps; a group works through its own tile exactly as the 8 warps of the one-group kernel did
    // (same fragments, same summation order per tile), the weights in shared memory serve both
    const int grp = threadIdx.x / ENC_THREADS, t = threadIdx.x % ENC_THREADS;
    const int warp = t >> 5, lane = t & 31, g = lane >> 2, tq = lane & 3;
    float(&X)[TR][MM_BS] = S.x[grp];
    float(&D1)[TR][MM_DS] = S.d1[grp];
    float(&D2)[TR][MM_DS] = S.d2[grp];
    const int K = a.K;
    const int nk8 = (K + 7) >> 3;  // 8-wide tiles covering K
    // weights -> shared memory, split once
    for (int idx = threadIdx.x; idx < MM_KP * MM_KP; idx += ENC_BWD_THREADS) {
        const int o = idx >> 6, i = idx & 63;
        const bool ok = o < K && i < K;
        uint32_t b, sm;
        split_tf32(ok ? a.w_loc[o * K + i] : 0.f, b, sm);
        S.wl_big[o][i] = __uint_as_float(b), S.wl_small[o][i] = __uint_as_float(sm);
        split_tf32(ok ? a.w_std[o * K + i] : 0.f, b, sm);
        S.ws_big[o][i] = __uint_as_float(b), S.ws_small[o][i] = __uint_as_float(sm);
    }
    const float ckl = a.kl_coef ? __ldg(a.kl_coef) : 0.f;
    // wgrad accumulators of this warp: matrix wm, output rows 16 mt .., column tiles 0 .. nk8 - 1
    const int wm = warp >> 2, mt = warp & 3;
    float gw[8][4];
#pragma unroll
    for (int n = 0; n < 8; ++n) gw[n][0] = gw[n][1] = gw[n][2] = gw[n][3] = 0.f;
    float gbias = 0.f;  // threads 0..63: column t of d1; 64..127: column t - 64 of d2
    const int64_t ntiles = (a.V + TR - 1) / TR;
    for (int64_t first = (int64_t)ENC_BWD_GROUPS * blockIdx.x; first < ntiles;
         first += (int64_t)ENC_BWD_GROUPS * gridDim.x) {
        __syncthreads();  // previous tile's reads are done
        const int64_t row0 = (first + grp) * TR;  // (past the last tile: every row is out of range, zeros flow through)
        // x tile and the two upstream-gradient tiles, zero padded to 64 columns
        for (int idx = t; idx < TR * (MM_KP / 2); idx += ENC_THREADS) {
            const int r = idx >> 5, o = 2 * (idx & 31);
            const int64_t row = row0 + r;
            float2 xv = make_float2(0.f, 0.f), dm = xv, ds = xv;
            if (row < a.V && o < K) {
                const int64_t off = row * K + o;
                xv = __ldg(reinterpret_cast<const float2*>(a.ptr + off));
                const float2 gv = a.grad_vsamp ? __ldg(reinterpret_cast<const float2*>(a.grad_vsamp + off))
                                               : make_float2(0.f, 0.f);
                const float2 nz = a.noise ? __ldg(reinterpret_cast<const float2*>(a.noise + off)) : make_float2(0.f, 0.f);
                const float2 m = __ldg(reinterpret_cast<const float2*>(a.mu + off));
                const float2 sg = __ldg(reinterpret_cast<const float2*>(a.sigma + off));
                dm = make_float2(fmaf(ckl, m.x, gv.x), fmaf(ckl, m.y, gv.y));
                const float dsg0 = fmaf(gv.x, nz.x, ckl * (sg.x - 1.f / sg.x));
                const float dsg1 = fmaf(gv.y, nz.y, ckl * (sg.y - 1.f / sg.y));
                ds = make_float2(dsg0 * (1.f - expf(-(sg.x - GLUE_EPS))), dsg1 * (1.f - expf(-(sg.y - GLUE_EPS))));
            }
            *reinterpret_cast<float2*>(&X[r][o]) = xv;
            *reinterpret_cast<float2*>(&D1[r][o]) = dm;
            *reinterpret_cast<float2*>(&D2[r][o]) = ds;
        }
        __syncthreads();
        // ---- bias gradients: column sums of d1 / d2
        if (t < 128) {
            const float(*d)[MM_DS] = t < 64 ? D1 : D2;
            const int col = t & 63;
            float s = 0.f;
#pragma unroll 8
            for (int r = 0; r < TR; ++r) s += d[r][col];
            gbias += s;
        }
        // ---- dgrad: rows 16 mt .. of this tile, column tiles [n_lo, n_hi)
        {
            const int n_lo = wm == 0 ? 0 : (nk8 + 1) / 2, n_hi = wm == 0 ? (nk8 + 1) / 2 : nk8;
            float acc[4][4];
#pragma unroll
            for (int n = 0; n < 4; ++n) acc[n][0] = acc[n][1] = acc[n][2] = acc[n][3] = 0.f;
            for (int mat = 0; mat < 2; ++mat) {
                const float(*d)[MM_DS] = mat == 0 ? D1 : D2;
                const float(*wb)[MM_BS] = mat == 0 ? S.wl_big : S.ws_big;
                const float(*wsm)[MM_BS] = mat == 0 ? S.wl_small : S.ws_small;
                for (int ks = 0; ks < nk8; ++ks) {
                    const int k0 = 8 * ks, r0 = 16 * mt;
                    uint32_t ab[4], as[4];
                    split_tf32(d[r0 + g][k0 + tq], ab[0], as[0]);
                    split_tf32(d[r0 + g + 8][k0 + tq], ab[1], as[1]);
                    split_tf32(d[r0 + g][k0 + tq + 4], ab[2], as[2]);
                    split_tf32(d[r0 + g + 8][k0 + tq + 4], ab[3], as[3]);
#pragma unroll
                    for (int n = 0; n < 4; ++n) {
                        const int nt = n_lo + n;
                        if (nt < n_hi) {
                            const int n0 = 8 * nt;
                            mma_3x(acc[n], ab, as, __float_as_uint(wb[k0 + tq][n0 + g]),
                                   __float_as_uint(wb[k0 + tq + 4][n0 + g]), __float_as_uint(wsm[k0 + tq][n0 + g]),
                                   __float_as_uint(wsm[k0 + tq + 4][n0 + g]));
                        }
                    }
                }
            }
            if (a.grad_ptr) {
#pragma unroll
                for (int n = 0; n < 4; ++n) {
                    const int nt = n_lo + n;
                    if (nt >= n_hi) continue;
                    const int col = 8 * nt + 2 * tq;
                    if (col >= K) continue;  // (K even: col + 1 < K as well)
                    const int64_t r_lo = row0 + 16 * mt + g, r_hi = r_lo + 8;
                    if (r_lo < a.V) *reinterpret_cast<float2*>(a.grad_ptr + r_lo * K + col) = make_float2(acc[n][0], acc[n][1]);
                    if (r_hi < a.V) *reinterpret_cast<float2*>(a.grad_ptr + r_hi * K + col) = make_float2(acc[n][2], acc[n][3]);
                }
            }
        }
        // ---- wgrad: gW_wm[16 mt + ., :] += d_wm^T x over the 64 rows of the tile
        {
            const float(*d)[MM_DS] = wm == 0 ? D1 : D2;
            const int m0 = 16 * mt;
            for (int ks = 0; ks < TR / 8; ++ks) {
                const int k0 = 8 * ks;  // rows of the tile
                uint32_t ab[4], as[4];  // A[m = o][k = r] = d[r][o]
                split_tf32(d[k0 + tq][m0 + g], ab[0], as[0]);
                split_tf32(d[k0 + tq][m0 + g + 8], ab[1], as[1]);
                split_tf32(d[k0 + tq + 4][m0 + g], ab[2], as[2]);
                split_tf32(d[k0 + tq + 4][m0 + g + 8], ab[3], as[3]);
#pragma unroll
                for (int n = 0; n < 8; ++n) {
                    if (n < nk8) {
                        const int n0 = 8 * n;
                        uint32_t bb0, bs0, bb1, bs1;  // B[k = r][n = i] = x[r][i]
                        split_tf32(X[k0 + tq][n0 + g], bb0, bs0);
                        split_tf32(X[k0 + tq + 4][n0 + g], bb1, bs1);
                        mma_3x(gw[n], ab, as, bb0, bb1, bs0, bs1);
                    }
                }
            }
        }
    }
    // per-group partials in the layout of graphenc_wgrad_reduce_kernel<KPW>: [2][KPW][KPW + 1]
    {
        const size_t part = (size_t)blockIdx.x * ENC_BWD_GROUPS + grp;
        float* base = a.gwpart + part * 2 * KPW * (KPW + 1) + (size_t)wm * KPW * (KPW + 1);
#pragma unroll
        for (int n = 0; n < 8; ++n) {
            if (n >= nk8) continue;
            const int i0 = 8 * n + 2 * tq;
            const int o_lo = 16 * mt + g, o_hi = o_lo + 8;
            if (i0 < K) {
                if (o_lo < K) base[o_lo * (KPW + 1) + i0] = gw[n][0], base[o_lo * (KPW + 1) + i0 + 1] = gw[n][1];
                if (o_hi < K) base[o_hi * (KPW + 1) + i0] = gw[n][2], base[o_hi * (KPW + 1) + i0 + 1] = gw[n][3];
            }
        }
        if (t < 128) {
            const int col = t & 63, m = t >> 6;
            if (col < K) a.gwpart[part * 2 * KPW * (KPW + 1) + (size_t)m * KPW * (KPW + 1) + col * (KPW + 1) + KPW] = gbias;
        }
    }
}

// ----------------------------------------------------------------------- forward, tensor cores
// mu = x W_loc^T + b_loc, s = x W_std^T + b_std per 64-vertex tile on mma.sync m16n8k8 TF32 x 3 (as the
// backward kernel above), then the epilogue of graphenc_fwd_kernel from the accumulator fragments.
//   warp w: rows 16 (w & 3) .., output-column tiles of half (w >> 2), both matrices (the epilogue needs
//   mu and sigma of the same element)
// A = x tile [64 x K] (stride 68 = 4 mod 32), B[k = i][n = o] = W[o][i]: the weights are kept
// transposed, [i][o], with stride 72 = 8 mod 32 (conflict-free B-pattern reads).
struct MmaFwdSmem {
    float x[TR][MM_DS];
    float wlT_big[MM_KP][MM_BS], wlT_small[MM_KP][MM_BS];  // [i][o]
    float wsT_big[MM_KP][MM_BS], wsT_small[MM_KP][MM_BS];
    float bl[MM_KP], bs[MM_KP];
    double red[ENC_THREADS / 32];
    unsigned int ticket;
};

__global__ void __launch_bounds__(ENC_THREADS) graphenc_fwd_mma_kernel(EncArgs a) {
    extern __shared__ float4 enc_smem_raw[];
    MmaFwdSmem& S = *reinterpret_cast<MmaFwdSmem*>(enc_smem_raw);
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31, g = lane >> 2, tq = lane & 3;
    const int K = a.K;
    const int nk8 = (K + 7) >> 3;
    for (int idx = t; idx < MM_KP * MM_KP; idx += ENC_THREADS) {
        const int o = idx >> 6, i = idx & 63;
        const bool ok = o < K && i < K;
        uint32_t b, sm;
        split_tf32(ok ? a.w_loc[o * K + i] : 0.f, b, sm);
        S.wlT_big[i][o] = __uint_as_float(b), S.wlT_small[i][o] = __uint_as_float(sm);
        split_tf32(ok ? a.w_std[o * K + i] : 0.f, b, sm);
        S.wsT_big[i][o] = __uint_as_float(b), S.wsT_small[i][o] = __uint_as_float(sm);
    }
    for (int o = t; o < MM_KP; o += ENC_THREADS) {
        S.bl[o] = o < K ? a.b_loc[o] : 0.f;
        S.bs[o] = o < K ? a.b_std[o] : 0.f;
    }
    const int mt = warp & 3, ch = warp >> 2;
    const int n_lo = ch == 0 ? 0 : (nk8 + 1) / 2, n_hi = ch == 0 ? (nk8 + 1) / 2 : nk8;
    double kl_acc = 0.0;
    const int64_t ntiles = (a.V + TR - 1) / TR;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        __syncthreads();  // previous tile's reads are done (first pass: weights are in place after the next one)
        const int64_t row0 = tile * TR;
        for (int idx = t; idx < TR * (MM_KP / 2); idx += ENC_THREADS) {
            const int r = idx >> 5, o = 2 * (idx & 31);
            const int64_t row = row0 + r;
            float2 xv = make_float2(0.f, 0.f);
            if (row < a.V && o < K) xv = __ldg(reinterpret_cast<const float2*>(a.ptr + row * K + o));
            *reinterpret_cast<float2*>(&S.x[r][o]) = xv;
        }
        __syncthreads();
        float am[4][4], as_[4][4];
#pragma unroll
        for (int n = 0; n < 4; ++n)
#pragma unroll
            for (int q = 0; q < 4; ++q) am[n][q] = as_[n][q] = 0.f;
        const int r0 = 16 * mt;
        for (int ks = 0; ks < nk8; ++ks) {
            const int k0 = 8 * ks;
            uint32_t ab[4], asm_[4];
            split_tf32(S.x[r0 + g][k0 + tq], ab[0], asm_[0]);
            split_tf32(S.x[r0 + g + 8][k0 + tq], ab[1], asm_[1]);
            split_tf32(S.x[r0 + g][k0 + tq + 4], ab[2], asm_[2]);
            split_tf32(S.x[r0 + g + 8][k0 + tq + 4], ab[3], asm_[3]);
#pragma unroll
            for (int n = 0; n < 4; ++n) {
                const int nt = n_lo + n;
                if (nt < n_hi) {
                    const int n0 = 8 * nt;
                    mma_3x(am[n], ab, asm_, __float_as_uint(S.wlT_big[k0 + tq][n0 + g]),
                           __float_as_uint(S.wlT_big[k0 + tq + 4][n0 + g]), __float_as_uint(S.wlT_small[k0 + tq][n0 + g]),
                           __float_as_uint(S.wlT_small[k0 + tq + 4][n0 + g]));
                    mma_3x(as_[n], ab, asm_, __float_as_uint(S.wsT_big[k0 + tq][n0 + g]),
                           __float_as_uint(S.wsT_big[k0 + tq + 4][n0 + g]), __float_as_uint(S.wsT_small[k0 + tq][n0 + g]),
                           __float_as_uint(S.wsT_small[k0 + tq + 4][n0 + g]));
                }
            }
        }
        float kl_row = 0.f;
#pragma unroll
        for (int n = 0; n < 4; ++n) {
            const int nt = n_lo + n;
            if (nt >= n_hi) continue;
            const int col = 8 * nt + 2 * tq;
            if (col >= K) continue;  // (K even: col + 1 < K as well)
            const float bl0 = S.bl[col], bl1 = S.bl[col + 1], bs0 = S.bs[col], bs1 = S.bs[col + 1];
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {  // rows r0 + g and r0 + g + 8
                const int64_t row = row0 + r0 + g + 8 * hh;
                if (row >= a.V) continue;
                const int64_t off = row * K + col;
                float2 nz = make_float2(0.f, 0.f);
                if (a.noise) nz = __ldg(reinterpret_cast<const float2*>(a.noise + off));
                const float m0 = am[n][2 * hh] + bl0, m1 = am[n][2 * hh + 1] + bl1;
                const float g0 = softplusf(as_[n][2 * hh] + bs0) + GLUE_EPS, g1 = softplusf(as_[n][2 * hh + 1] + bs1) + GLUE_EPS;
                *reinterpret_cast<float2*>(a.mu + off) = make_float2(m0, m1);
                *reinterpret_cast<float2*>(a.sigma + off) = make_float2(g0, g1);
                *reinterpret_cast<float2*>(a.vsamp + off) = make_float2(fmaf(nz.x, g0, m0), fmaf(nz.y, g1, m1));
                kl_row += (-logf(g0) + 0.5f * (g0 * g0 + m0 * m0) - 0.5f) +
                          (-logf(g1) + 0.5f * (g1 * g1 + m1 * m1) - 0.5f);
            }
        }
        kl_acc += (double)kl_row;
    }
    // KL: deterministic block reduction, per-CTA partial, last CTA sums the partials in order
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) kl_acc += __shfl_xor_sync(FULL, kl_acc, o);
    if (lane == 0) S.red[warp] = kl_acc;
    __syncthreads();
    if (t == 0) {
        double s = 0.0;
        for (int w = 0; w < ENC_THREADS / 32; ++w) s += S.red[w];
        a.klpart[blockIdx.x] = s;
        __threadfence();
        S.ticket = atomicAdd(a.counter, 1u);
    }
    __syncthreads();
    if (S.ticket == gridDim.x - 1 && t == 0) {
        __threadfence();
        double s = 0.0;
        for (unsigned q = 0; q < gridDim.x; ++q) s += __ldcg(a.klpart + q);
        a.kl_sum[0] = (float)s;
        *a.counter = 0u;  // ready for the next call on this workspace
    }
}

// fixed-order sum of the per-CTA weight-gradient partials; one thread per (matrix, out, in | bias)
template <int KP>
__global__ void graphenc_wgrad_reduce_kernel(EncArgs a, int ncta) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int per = KP * (KP + 1);
    if (idx >= 2 * per) return;
    const int mat = idx / per, rem = idx - mat * per;
    const int o = rem / (KP + 1), i = rem - o * (KP + 1);
    if (o >= a.K || (i >= a.K && i != KP)) return;
    float s = 0.f;
#pragma unroll 8
    for (int q = 0; q < ncta; ++q) s += a.gwpart[(size_t)q * 2 * per + idx];
    if (i == KP) {
        float* gb = mat == 0 ? a.grad_b_loc : a.grad_b_std;
        if (gb) gb[o] = s;
    } else {
        float* gw = mat == 0 ? a.grad_w_loc : a.grad_w_std;
        if (gw) gw[o * a.K + i] = s;
    }
}

int pick_kp(int K) { return K <= 52 ? 52 : (K <= 64 ? 64 : 0); }

// Persistent CTAs, tiles strided over them.  The kernels' shared memory admits one CTA per SM, so more CTAs than
// SMs only queue a second wave -- and while a kernel has blocks waiting to be dispatched, the blocks of every
// kernel launched after it (the objective tail, which runs beside the graph branch's backward pass) wait behind
// them.  The fewest CTAs with the makespan of a full grid (813 tiles: 6 per CTA with 148 CTAs and with 136) also
// leave a dozen SMs to the discriminator / encoder backward kernels, which need whole SMs, and halve the number
// of weight-gradient partials.
int enc_grid(int64_t V) {
    const int64_t ntiles = (V + TR - 1) / TR;
    const int64_t sms = sm_count();
    if (ntiles <= sms) return (int)ntiles;
    const int64_t per_cta = (ntiles + sms - 1) / sms;
    return (int)((ntiles + per_cta - 1) / per_cta);
}

// Backward pass: beside it run the discriminator's and the data encoders' backward kernels, 16 CTAs each that need a
// whole SM (mlp_fused.cu).  With the dozen SMs enc_grid leaves they take two waves; 30 free SMs hold two of them at
// once.  The head kernel then works through one more unit per CTA -- it has that slack when the graph is small
// (measured at the PBMC shape: 0.645 -> 0.620 ms per step; the same reserve on the forward kernel, which is the long
// pole of its phase, bought nothing).  `unit_tiles`: 64-row tiles a CTA takes per iteration (2 for the tensor-core
// kernel with its two warp groups).  Large graphs (200 k vertices): the head kernel is several times longer than
// the MLP kernels beside it and ends the step itself, so every CTA counts (1.31 -> 1.37 ms per step at the
// triple-omics shape with the reserve).
int enc_grid_bwd(int64_t V, int unit_tiles) {
    static const int full_grid = getenv("GLUE_ENC_BWD_FULL_GRID") ? atoi(getenv("GLUE_ENC_BWD_FULL_GRID")) : 0;
    const int64_t ntiles = (V + TR - 1) / TR;
    const int64_t units = (ntiles + unit_tiles - 1) / unit_tiles;
    const int64_t sms = sm_count();
    if (units <= sms) return (int)units;
    const int64_t full = (units + sms - 1) / sms;  // units per CTA on a full grid
    int64_t per_cta = full;
    if (!full_grid && sms > 60 && full * unit_tiles >= 4 && full * unit_tiles <= 8)
        per_cta = std::min((units + sms - 31) / (sms - 30), full + 1);
    return (int)((units + per_cta - 1) / per_cta);
}

size_t enc_ws_layout(int64_t V, int KP, char* base, EncArgs* a) {
    const int ncta = ENC_BWD_GROUPS * sm_count();  // (upper bound: one weight-gradient partial per warp group of a CTA)
    size_t off = 0;
    auto take = [&](size_t bytes) {
        char* p = base ? base + off : nullptr;
        off += align_up(bytes, 256);
        return p;
    };
    unsigned int* counter = (unsigned int*)take(16);
    double* klpart = (double*)take((size_t)ncta * 8);
    float* gwpart = (float*)take((size_t)ncta * 2 * KP * (KP + 1) * 4);
    if (a) a->counter = counter, a->klpart = klpart, a->gwpart = gwpart;
    return off;
}

template <int KP>
int run_fwd(EncArgs a, cudaStream_t st, int impl) {
    if (impl != 1) {  // default: warp-level tensor cores (TF32 x 3)
        const size_t smem = sizeof(MmaFwdSmem);
        GLUE_SMEM_ONCE(graphenc_fwd_mma_kernel, smem);
        graphenc_fwd_mma_kernel<<<enc_grid(a.V), ENC_THREADS, smem, st>>>(a);
    } else {
        const size_t smem = sizeof(EncSmem<KP>);
        auto k = graphenc_fwd_kernel<KP>;
        GLUE_SMEM_ONCE(k, smem);
        k<<<enc_grid(a.V), ENC_THREADS, smem, st>>>(a);
    }
    GLUE_CHECK_LAUNCH();
    return 0;
}

template <int KP>
int run_bwd(EncArgs a, cudaStream_t st, int impl) {
    // impl: 0 = default (warp-level tensor cores, TF32 x 3), 1 = FFMA kernel, 2 = tensor cores
    const bool use_mma = impl != 1;
    const int ncta = enc_grid_bwd(a.V, use_mma ? ENC_BWD_GROUPS : 1);
    const int nparts = use_mma ? ENC_BWD_GROUPS * ncta : ncta;
    if (use_mma) {
        const size_t smem = sizeof(MmaSmem);
        GLUE_SMEM_ONCE(graphenc_bwd_mma_kernel, smem);
        graphenc_bwd_mma_kernel<<<ncta, ENC_BWD_THREADS, smem, st>>>(a, KP);
    } else {
        const size_t smem = sizeof(EncSmem<KP>);
        auto k = graphenc_bwd_kernel<KP>;
        GLUE_SMEM_ONCE(k, smem);
        k<<<ncta, ENC_THREADS, smem, st>>>(a);
    }
    GLUE_CHECK_LAUNCH();
    const int total = 2 * KP * (KP + 1);
    graphenc_wgrad_reduce_kernel<KP><<<(total + 255) / 256, 256, 0, st>>>(a, nparts);
    GLUE_CHECK_LAUNCH();
    return 0;
}

}  // namespace
}  // namespace glue

extern "C" size_t glue_graphenc_workspace_bytes(int64_t V, int32_t K) {
    const int KP = glue::pick_kp(K);
    if (V <= 0 || KP == 0) return 0;
    return glue::enc_ws_layout(V, KP, nullptr, nullptr);
}

extern "C" int glue_graphenc_head_fwd_impl(const float* ptr, const float* w_loc, const float* b_loc, const float* w_std,
                                           const float* b_std, const float* noise, int64_t V, int32_t K, float* mu,
                                           float* sigma, float* vsamp, float* kl_sum, void* workspace,
                                           size_t workspace_bytes, int32_t impl, void* stream) {
    using namespace glue;
    if (!ptr || !w_loc || !b_loc || !w_std || !b_std || !mu || !sigma || !vsamp || !kl_sum || !workspace)
        return GLUE_E_BADARG;
    if (V <= 0 || K <= 0) return GLUE_E_BADARG;
    const int KP = pick_kp(K);
    if (KP == 0 || (K & 1)) return GLUE_E_UNSUPPORTED;
    EncArgs a = {};
    a.ptr = ptr, a.w_loc = w_loc, a.b_loc = b_loc, a.w_std = w_std, a.b_std = b_std, a.noise = noise;
    a.mu = mu, a.sigma = sigma, a.vsamp = vsamp, a.kl_sum = kl_sum, a.V = V, a.K = K;
    if (workspace_bytes < enc_ws_layout(V, KP, nullptr, nullptr)) return GLUE_E_WORKSPACE;
    enc_ws_layout(V, KP, (char*)workspace, &a);
    cudaStream_t st = (cudaStream_t)stream;
    GLUE_CUDA(cudaMemsetAsync(a.counter, 0, 16, st));
    return KP == 52 ? run_fwd<52>(a, st, impl) : run_fwd<64>(a, st, impl);
}

extern "C" int glue_graphenc_head_fwd(const float* ptr, const float* w_loc, const float* b_loc, const float* w_std,
                                      const float* b_std, const float* noise, int64_t V, int32_t K, float* mu,
                                      float* sigma, float* vsamp, float* kl_sum, void* workspace,
                                      size_t workspace_bytes, void* stream) {
    return glue_graphenc_head_fwd_impl(ptr, w_loc, b_loc, w_std, b_std, noise, V, K, mu, sigma, vsamp, kl_sum, workspace,
                                       workspace_bytes, 0, stream);
}

extern "C" int glue_graphenc_head_bwd_impl(const float* ptr, const float* w_loc, const float* w_std, const float* noise,
                                           const float* mu, const float* sigma, const float* grad_vsamp,
                                           const float* kl_coef, int64_t V, int32_t K, float* grad_ptr,
                                           float* grad_w_loc, float* grad_b_loc, float* grad_w_std, float* grad_b_std,
                                           void* workspace, size_t workspace_bytes, int32_t impl, void* stream) {
    using namespace glue;
    if (!ptr || !w_loc || !w_std || !mu || !sigma || !workspace) return GLUE_E_BADARG;
    if (V <= 0 || K <= 0) return GLUE_E_BADARG;
    const int KP = pick_kp(K);
    if (KP == 0 || (K & 1)) return GLUE_E_UNSUPPORTED;
    EncArgs a = {};
    a.ptr = ptr, a.w_loc = w_loc, a.w_std = w_std, a.noise = noise, a.mu = const_cast<float*>(mu);
    a.sigma = const_cast<float*>(sigma), a.grad_vsamp = grad_vsamp, a.kl_coef = kl_coef;
    // biases are not read by the backward pass; point them at the weights to keep load_weights simple
    a.b_loc = w_loc, a.b_std = w_std;
    a.grad_ptr = grad_ptr, a.grad_w_loc = grad_w_loc, a.grad_b_loc = grad_b_loc;
    a.grad_w_std = grad_w_std, a.grad_b_std = grad_b_std, a.V = V, a.K = K;
    if (workspace_bytes < enc_ws_layout(V, KP, nullptr, nullptr)) return GLUE_E_WORKSPACE;
    enc_ws_layout(V, KP, (char*)workspace, &a);
    cudaStream_t st = (cudaStream_t)stream;
    return KP == 52 ? run_bwd<52>(a, st, impl) : run_bwd<64>(a, st, impl);
}

extern "C" int glue_graphenc_head_bwd(const float* ptr, const float* w_loc, const float* w_std, const float* noise,
                                      const float* mu, const float* sigma, const float* grad_vsamp,
                                      const float* kl_coef, int64_t V, int32_t K, float* grad_ptr,
                                      float* grad_w_loc, float* grad_b_loc, float* grad_w_std, float* grad_b_std,
                                      void* workspace, size_t workspace_bytes, void* stream) {
    return glue_graphenc_head_bwd_impl(ptr, w_loc, w_std, noise, mu, sigma, grad_vsamp, kl_coef, V, K, grad_ptr,
                                       grad_w_loc, grad_b_loc, grad_w_std, grad_b_std, workspace, workspace_bytes, 0,
                                       stream);
}
