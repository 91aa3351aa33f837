// Fused data decoders, "v3": one launch sequence per modality for up to three loss terms that share
// the observations x, the feature table v and the per-feature parameters (the reconstruction,
// joint-cross and real-cross terms of PairedSCGLUETrainer, scglue/models/scglue.py:603-652), with
// TMA-staged operands.
//
// Same mathematics as decoder_tc.cu ("v2": three bf16 products per fp32 product on tcgen05, logits
// in TMEM only, passes A = soft-max row statistics, B = likelihood + all G terms, C = the -p c
// soft-max terms).  What changes:
//
//   * a pre-pass (presplit_kernel) writes the bf16 hi / lo images of v[vidx] ([F, 64], K padded with
//     zeros, column K = 1 / softplus(scale_lin) for the streamlined NB path) and of every term's u
//     block once per call; the three passes stage their operand tiles from those images with
//     cp.async.bulk.tensor (SWIZZLE_128B tensor maps, mbarrier expect_tx) -- no thread touches an
//     operand row any more (v2: one thread per row, 4.1 k of 4.3 k cycles per tile in pass A);
//   * a "unit" is (feature tile, term).  Units of one tile run back to back: the v tile is staged
//     once per tile, d/dv of the tile accumulates over the terms in one TMEM accumulator and leaves
//     once, the observations of the tile are read from HBM once (the other terms hit L1 / L2), and
//     every fixed cost (launch, set-up, tail) is paid once per pass instead of once per term;
//   * the epilogue warps write their coefficient tile back into the logits accumulator
//     (tcgen05.st) instead of into shared memory; four converter warps (one per TMEM lane quarter)
//     turn it into the bf16 hi / lo shared-memory operand of the two gradient products, so the
//     epilogue never waits for the tensor pipe; the same warps move finished d/dv tiles out.
//
// Warp roles (448 threads, 1 CTA / SM): warps 0-7 epilogue (warp w: TMEM lanes 32 (w & 3).., half
// w >> 2 of the cells), warps 8-11 converters / d/dv write-out, warp 12 MMA issuer, warp 13 TMA.
// TMEM: logits 2 x 128 columns, d/dv 64, d/du 64 per term (<= 3 terms) = 512.
#include <cuda.h>
#include <stddef.h>
#include <stdlib.h>

#include <type_traits>

#include "decoder_common.cuh"
#include "tc_common.cuh"
#include "tc_math.cuh"

namespace glue {
using namespace tc;
using namespace tcm;

namespace v3 {

constexpr int MAXSEG = GLUE_DECODER_MAX_TERMS;
constexpr int NTHREADS = 448;
constexpr int EPI_WARPS = 8;
constexpr int EPI_THREADS = 256;
constexpr int CONV_W0 = 8;  // warps 8-11
constexpr int CONV_THREADS = 128;
constexpr int MMA_WARP = 12;
constexpr int TMA_WARP = 13;
constexpr int NB_MAX = 24;

constexpr uint32_t COL_ACC = 0;   // logits: 2 x 128 columns
constexpr uint32_t COL_D1 = 256;  // d/dv tile: 64 columns (accumulates over the terms of a tile)
constexpr uint32_t COL_D2 = 320;  // d/du: 64 columns per term
constexpr uint32_t TMEM_COLS = 512;
constexpr size_t TRACE_BYTES = 32768;  // diagnostic stamps (impl bit 9): 768 + 960 + 1440 int64 values

constexpr uint64_t DESC_K = smem_desc_base(16, 1024);
constexpr uint64_t DESC_MN1 = smem_desc_base(1024, 1024);
constexpr uint64_t DESC_MN_X = smem_desc_base(16384, 1024);
constexpr uint32_t IDESC_FWD = instr_desc_bf16(128, 128, 0, 0);
constexpr uint32_t IDESC_D1 = instr_desc_bf16(128, 64, 0, 1);
constexpr uint32_t IDESC_D2 = instr_desc_bf16(128, 64, 1, 1);

struct Term {
    const float* u;           // [B, K]
    const float* row_weight;  // [B] or NULL
    float* loss;              // [1] or NULL
    float* grad_u;            // [B, K] or NULL
    float gscale;             // coefficient folded into the weights (1 when unset)
    float w0;                 // uniform weight -gscale / (B F) of log_prob
};

struct Ws3 {
    uint8_t* vimg;           // [2 Fpad][128 B]: hi rows, then lo rows
    uint8_t* uimg;           // [2 S 128][128 B]
    float2* ptab;            // [n_batches][Fpad] (softplus(scale_lin), bias) * log2(e); bias = -inf beyond F
    float* pmax;             // [2 ncta][S 128] (base 2)
    float* psum;             // [2 ncta][S 128]
    float* lse;              // [S 128] natural log
    float* cpart;            // [ncta][S 128]
    float* c;                // [S 128]
    float* apart;            // [ncta][S][128][64]
    float* losspart;         // [ncta][S]
    int* badpart;            // [ncta][S]
    float* rescale;          // [1]
    unsigned int* counters;  // [4]
    long long* trace;        // [3 kernels][16 units][16 events] clock64 stamps of CTA 0 (impl bit 9), else NULL
};

// diagnostic timeline (impl bit 9): kernel 0 = pass A, 1 = pass B, 2 = pass C; unit 15 holds kernel-wide events
#define T3(kern, unit, ev)                                                                       \
    do {                                                                                         \
        if (P.ws.trace != nullptr && blockIdx.x == 0 && (unit) < 16)                             \
            P.ws.trace[((kern) * 16 + (unit)) * 16 + (ev)] = clock64();                          \
    } while (0)

__device__ __forceinline__ long long globaltimer_ns() {
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// per-CTA lifetime (globaltimer, ns): [3 kernels][160 CTAs][entry, exit] behind the event table
#define T3CTA(kern, which)                                                                       \
    do {                                                                                         \
        if (P.ws.trace != nullptr && blockIdx.x < 160)                                           \
            P.ws.trace[768 + ((kern) * 160 + blockIdx.x) * 2 + (which)] = globaltimer_ns();      \
    } while (0)
// per-CTA epilogue accounting (warp 0 lane 0): [3 kernels][160 CTAs][smid, loop cycles, wait cycles] at 1728
#define T3EPI(kern, which, val)                                                                  \
    do {                                                                                         \
        if (P.ws.trace != nullptr && blockIdx.x < 160)                                           \
            P.ws.trace[1728 + ((kern) * 160 + blockIdx.x) * 3 + (which)] = (long long)(val);     \
    } while (0)
#define T3NS(kern, ev)                                                                           \
    do {                                                                                         \
        if (P.ws.trace != nullptr && blockIdx.x == 0) P.ws.trace[((kern) * 16 + 15) * 16 + (ev)] = globaltimer_ns(); \
    } while (0)

struct P3 {
    glue_decoder_args a;  // shared fields (u, row_weight, grad_scale, loss, grad_u are per term below)
    Term term[MAXSEG];
    Ws3 ws;
    int S;       // terms = units per feature tile
    int ntiles;  // feature tiles of 128
    int Fpad;    // ntiles * 128
    int ncell;   // cells (<= 128)
    int cinv;    // operand column K of v carries 1 / softplus(scale_lin)
};

struct Bars3 {
    uint64_t v_full[2], v_empty[2], u_full[2], u_empty[2], acc_full[2], acc_empty[2], coef_full[2];
    uint64_t x_full[4], x_empty, d1_full, d1_empty, d2_full;  // x_full: per block of 32 cells
};

struct __align__(1024) SmemMain3 {
    uint8_t u[2][2][OP_BYTES];  // [ring slot][hi, lo]
    uint8_t v[2][2][OP_BYTES];  // [stage][hi, lo]
    uint8_t x_hi[2 * OP_BYTES], x_lo[2 * OP_BYTES];  // [cell block of 64][128 feature rows][128 B]
    float stage[4][32][33];     // d/dv rows on their way out, one block per converter warp
    int rowidx[4][32];          // row of d/dv (= row of v) of every feature row of the tile being written out (-1: none)
    float pgrad[4][128];        // parameter-table gradient partials of the upper-half epilogue threads
    alignas(16) uint32_t xoff32[136];  // element offset of every cell's row of x (streamlined path); 8 pad entries
    alignas(16) int xrow[128];         // row of x of every cell
    alignas(16) int bidx[128];
    alignas(16) float l[128];
    alignas(16) float lse[MAXSEG][128];   // -lse in base 2
    alignas(16) float c[MAXSEG][128];
    alignas(16) float wrow[MAXSEG][128];  // weight of log_prob per cell
    alignas(16) float zeros[128];         // weights seen by feature rows beyond F
    float cw[2][4][128];
    float c_acc[MAXSEG][128];
    float red_l[MAXSEG][EPI_WARPS];
    int red_b[MAXSEG][EPI_WARPS];
    Bars3 bars;
    uint32_t tmem_base;
    uint32_t ticket;
};

struct __align__(1024) SmemStats3 {
    uint8_t u[2][2][OP_BYTES];
    uint8_t v[2][2][OP_BYTES];
    float2 ptab[2][NB_MAX][128];  // per (stage, parameter row, feature of the tile)
    int bidx[128];
    Bars3 bars;
    uint32_t tmem_base;
};

static_assert(sizeof(SmemMain3) <= 227 * 1024, "shared memory per CTA");
static_assert(sizeof(SmemStats3) <= 227 * 1024, "shared memory per CTA");
static_assert(offsetof(SmemMain3, wrow) % 16 == 0 && offsetof(SmemMain3, lse) % 16 == 0 &&
                  offsetof(SmemMain3, xoff32) % 16 == 0 && offsetof(SmemMain3, zeros) % 16 == 0 &&
                  offsetof(SmemMain3, c) % 16 == 0 && offsetof(SmemMain3, l) % 16 == 0 &&
                  offsetof(SmemMain3, bidx) % 16 == 0,
              "vector loads of the per-cell tables");
static_assert(offsetof(SmemStats3, ptab) % 16 == 0, "bulk copy destination");

// predicated global store (no branch: a run of these keeps its loads and address arithmetic batched)
__device__ __forceinline__ void st_global_pred(float* ptr, float val, bool pred) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %2, 0;\n\t"
        "@p st.global.f32 [%0], %1;\n\t}" ::"l"(ptr),
        "f"(val), "r"((int)pred)
        : "memory");
}

__device__ __forceinline__ int orig_row(const glue_decoder_args& a, int pos) {
    return a.row_order ? a.row_order[pos] : pos;
}

// ---------------------------------------------------------------------------------------------
// pre-pass: bf16 hi / lo operand images (plain row-major [rows][64]; TMA applies the swizzle),
// the parameter table of pass A, zeroed tickets.  16 threads per row, 4 columns each.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void write_part(uint8_t* img, int64_t row, int64_t nrows, int part, const float* src, int K,
                                           int colx, float extra) {
    float vals[8];
#pragma unroll
    for (int q = 0; q < 4; ++q) {  // (rows are 8-byte aligned: K is even)
        const int k = 8 * part + 2 * q;
        float2 pr = make_float2(0.f, 0.f);
        if (src != nullptr && k < K) pr = __ldg(reinterpret_cast<const float2*>(src + k));
        vals[2 * q] = pr.x, vals[2 * q + 1] = pr.y;
    }
#pragma unroll
    for (int q = 0; q < 8; ++q)
        if (8 * part + q == colx) vals[q] = extra;
    uint32_t h[4], l[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) split2(vals[2 * q], vals[2 * q + 1], h[q], l[q]);
    *reinterpret_cast<uint4*>(img + row * ROW_BYTES + part * 16) = make_uint4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<uint4*>(img + (nrows + row) * ROW_BYTES + part * 16) = make_uint4(l[0], l[1], l[2], l[3]);
}

__global__ void __launch_bounds__(256) presplit_kernel(P3 P) {
    const glue_decoder_args& a = P.a;
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int part = (int)(gid & 7);
    const int64_t row = gid >> 3;
    const int64_t nv = P.Fpad, nu = (int64_t)P.S * 128;
    if (gid < 4) P.ws.counters[gid] = 0u;
    if (row < nv) {
        const int64_t j = row;
        const float* src = nullptr;
        float extra = 0.f;
        if (j < a.F) {
            const int64_t vr = a.vidx ? a.vidx[j] : j;
            src = a.v + vr * a.K;
            if (P.cinv && part == (a.K >> 3)) {
                const float sp = softplusf(a.scale_lin[j]);
                extra = sp > 1e-18f ? 1.f / sp : 0.f;
            }
        }
        write_part(P.ws.vimg, row, nv, part, src, a.K, P.cinv ? a.K : -1, extra);
        if (P.ws.ptab != nullptr) {
            for (int b = part; b < a.n_batches; b += 8) {
                float2 pr = make_float2(0.f, -INFINITY);
                if (j < a.F) {
                    const int64_t off = (int64_t)b * a.F + j;
                    pr = make_float2(softplusf(a.scale_lin[off]) * LOG2EF, a.bias[off] * LOG2EF);
                }
                P.ws.ptab[(int64_t)b * P.Fpad + j] = pr;
            }
        }
    } else if (row < nv + nu) {
        const int64_t r = row - nv;
        const int seg = (int)(r >> 7), cell = (int)(r & 127);
        const float* src = nullptr;
        if (cell < P.ncell) src = P.term[seg].u + (int64_t)orig_row(a, cell) * a.K;
        write_part(P.ws.uimg, r, nu, part, src, a.K, -1, 0.f);
    }
}

// logits tile: acc = A B^T with both operands K-major [128][64], three bf16 products (small terms first)
__device__ __forceinline__ void issue_logits(uint32_t tmem_acc, const uint8_t* a_hi, const uint8_t* a_lo,
                                             const uint8_t* b_hi, const uint8_t* b_lo) {
    const uint32_t ah = smem_u32(a_hi), al = smem_u32(a_lo), bh = smem_u32(b_hi), bl = smem_u32(b_lo);
    const uint32_t as[3] = {al, ah, ah}, bs[3] = {bh, bl, bh};
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
        for (uint32_t ks = 0; ks < 4; ++ks)
            umma_bf16(tmem_acc, smem_desc(DESC_K, as[c] + ks * 32), smem_desc(DESC_K, bs[c] + ks * 32), IDESC_FWD,
                      (c | ks) ? 1u : 0u);
}

// operand loads of one CTA, in consumption order: per feature tile the v tile, then the u block of each
// of its units (two-slot ring).  Loads [op_begin, op_end) of that sequence; the first ones wait for
// nothing and are issued before the per-cell tables are complete (tma_prologue_ops).
template <typename SM>
__device__ __forceinline__ void tma_ops(SM& S, const CUtensorMap* mapV, const CUtensorMap* mapU, const P3& P, int cta,
                                        int ncta, bool with_ptab, int kern, int op_begin, int op_end) {
    const int nS = P.S, per = 1 + nS;
    for (int op = op_begin; op < op_end; ++op) {
        const int it = op / per, k = op - it * per;
        if (k == 0) {
            const int vs = it & 1, f0 = (cta + it * ncta) * 128;
            mbar_wait(&S.bars.v_empty[vs], ((it >> 1) & 1) ^ 1);
            const uint32_t pbytes = with_ptab ? (uint32_t)P.a.n_batches * 1024u : 0u;
            mbar_arrive_expect_tx(&S.bars.v_full[vs], 2 * OP_BYTES + pbytes);
            tma_load_2d(S.v[vs][0], mapV, 0, f0, &S.bars.v_full[vs]);
            tma_load_2d(S.v[vs][1], mapV, 0, P.Fpad + f0, &S.bars.v_full[vs]);
            T3(kern, it * nS, 0);
            if constexpr (std::is_same<SM, SmemStats3>::value) {
                if (with_ptab)
                    for (int b = 0; b < P.a.n_batches; ++b)
                        bulk_load(&S.ptab[vs][b][0], P.ws.ptab + (int64_t)b * P.Fpad + f0, 1024u, &S.bars.v_full[vs]);
            }
        } else {
            const int seg = k - 1, n = it * nS + seg, us = n & 1;
            mbar_wait(&S.bars.u_empty[us], ((n >> 1) & 1) ^ 1);
            mbar_arrive_expect_tx(&S.bars.u_full[us], 2 * OP_BYTES);
            tma_load_2d(S.u[us][0], mapU, 0, seg * 128, &S.bars.u_full[us]);
            tma_load_2d(S.u[us][1], mapU, 0, (nS + seg) * 128, &S.bars.u_full[us]);
            T3(kern, n, 1);
        }
    }
}
__device__ __forceinline__ int tma_prologue_ops(int nS, int ntl) {
    const int total = ntl * (1 + nS), free_ops = nS == 1 ? 4 : 3;  // v0 u0 v1 u1 | v0 u0 u1: fresh buffers
    return total < free_ops ? total : free_ops;
}

template <typename SM>
__device__ __forceinline__ void init_bars(SM& S, bool grad, int v_empty_count) {
    for (int s = 0; s < 2; ++s) {
        mbar_init(&S.bars.v_full[s], 1);
        mbar_init(&S.bars.v_empty[s], v_empty_count);
        mbar_init(&S.bars.u_full[s], 1);
        mbar_init(&S.bars.u_empty[s], 1);
        mbar_init(&S.bars.acc_full[s], 1);
        mbar_init(&S.bars.acc_empty[s], grad ? CONV_THREADS / 32 : EPI_WARPS);
        mbar_init(&S.bars.coef_full[s], EPI_WARPS);
    }
    for (int cb = 0; cb < 4; ++cb) mbar_init(&S.bars.x_full[cb], CONV_THREADS / 32);
    mbar_init(&S.bars.x_empty, 1);
    mbar_init(&S.bars.d1_full, 1);
    mbar_init(&S.bars.d1_empty, CONV_THREADS / 32);
    mbar_init(&S.bars.d2_full, 1);
    mbar_fence_init();
}

// ---------------------------------------------------------------------------------------------
// pass A: soft-max row statistics of every unit.  D[cell, feature] = U V^T; thread <-> cell;
// base-2 running (max, sum) per term in registers.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NTHREADS, 1)
    rowstats3_kernel(const __grid_constant__ CUtensorMap mapV, const __grid_constant__ CUtensorMap mapU, const P3 P) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    SmemStats3& S = *reinterpret_cast<SmemStats3*>(smem_raw);
    if ((smem_u32(smem_raw) & 1023u) != 0) __trap();
    const glue_decoder_args& a = P.a;
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const int cta = blockIdx.x, ncta = gridDim.x;
    const int ntl = (P.ntiles - cta + ncta - 1) / ncta;
    const int nS = P.S, nunits = ntl * nS;
    if (t == 0) T3(0, 15, 0);

    if (t == 0) init_bars(S, false, EPI_WARPS);
    if (warp == MMA_WARP) tmem_alloc(&S.tmem_base, 256);
    if (t < 128) {
        const bool ok = t < P.ncell;
        const int orig = ok ? orig_row(a, t) : 0;
        S.bidx[t] = (ok && a.b) ? (int)a.b[orig] : 0;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = S.tmem_base;
    if (t == 0) T3(0, 15, 1);

    if (warp == TMA_WARP) {
        if (lane == 0) tma_ops(S, &mapV, &mapU, P, cta, ncta, true, 0, 0, ntl * (1 + nS));
    } else if (warp == MMA_WARP) {
        if (lane == 0) {
            for (int n = 0; n < nunits; ++n) {
                const int it = n / nS, seg = n - it * nS, vs = it & 1, us = n & 1, as = n & 1;
                if (seg == 0) mbar_wait(&S.bars.v_full[vs], (it >> 1) & 1);
                mbar_wait(&S.bars.u_full[us], (n >> 1) & 1);
                mbar_wait(&S.bars.acc_empty[as], ((n >> 1) & 1) ^ 1);
                tc_fence_after();
                issue_logits(tmem_base + COL_ACC + as * 128, S.u[us][0], S.u[us][1], S.v[vs][0], S.v[vs][1]);
                umma_commit(&S.bars.acc_full[as]);
                umma_commit(&S.bars.u_empty[us]);
                T3(0, n, 2);
                // (the parameter table of the stage is read by the epilogue: released there)
            }
        }
    } else if (warp < EPI_WARPS) {
        const int q = warp & 3, h = warp >> 2;
        const int cell = 32 * q + lane;
        const int my_b = S.bidx[cell];
        float m[MAXSEG], ssum[MAXSEG];
#pragma unroll
        for (int s = 0; s < MAXSEG; ++s) m[s] = -INFINITY, ssum[s] = 0.f;
        for (int n = 0; n < nunits; ++n) {
            const int it = n / nS, seg = n - it * nS, vs = it & 1, as = n & 1;
            const int f0 = (cta + it * ncta) * 128;
            const int nvalid = min(128, a.F - f0);
            if (t == 0) T3(0, n, 5);
            mbar_wait(&S.bars.acc_full[as], (n >> 1) & 1);
            tc_fence_after();
            if (t == 0) T3(0, n, 6);
            float mm = -INFINITY, ss = 0.f;
#pragma unroll
            for (int s = 0; s < MAXSEG; ++s)
                if (s == seg) mm = m[s], ss = ssum[s];
#pragma unroll 1
            for (int cc = 0; cc < 2; ++cc) {
                const int c0 = 64 * h + 32 * cc;
                if (c0 < nvalid) {
                    uint32_t d[32];
                    tmem_ld32(tmem_base + ((uint32_t)(32 * q) << 16) + COL_ACC + as * 128 + c0, d);
                    tmem_ld_wait();
                    const float2* tab = &S.ptab[vs][my_b][c0];
                    float av[32];
                    float cm = -INFINITY;
#pragma unroll
                    for (int r = 0; r < 32; ++r) {
                        const float2 pr = tab[r];
                        av[r] = fmaf(pr.x, __uint_as_float(d[r]), pr.y);  // (bias = -inf beyond F)
                        cm = fmaxf(cm, av[r]);
                    }
                    const float nm = fmaxf(mm, cm);
                    if (nm > -INFINITY) {
                        float part = 0.f;
#pragma unroll
                        for (int r = 0; r < 32; ++r) part += ex2f(av[r] - nm);
                        ss = ss * ex2f(mm - nm) + part;
                        mm = nm;
                    }
                }
            }
#pragma unroll
            for (int s = 0; s < MAXSEG; ++s)
                if (s == seg) m[s] = mm, ssum[s] = ss;
            if (t == 0) T3(0, n, 7);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&S.bars.acc_empty[as]);
                if (seg == nS - 1) mbar_arrive(&S.bars.v_empty[vs]);
            }
        }
        if (cell < P.ncell) {
#pragma unroll
            for (int s = 0; s < MAXSEG; ++s)
                if (s < nS) {
                    const int64_t slot = (int64_t)(cta * 2 + h) * (nS * 128) + s * 128 + cell;
                    P.ws.pmax[slot] = m[s];
                    P.ws.psum[slot] = ssum[s];
                }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == MMA_WARP) tmem_dealloc(tmem_base, 256);
    if (t == 0) T3(0, 15, 5);
}

// lse (natural log) from the base-2 per-(CTA, column half) partials; block = 32 rows x 32 strided
// sub-ranges of the partials, sub-results meet in a fixed order.  rows = S * 128 (term, cell).
__global__ void __launch_bounds__(1024) lse3_kernel(Ws3 ws, int nparts, int nrows) {
    __shared__ float red[2][32][33];
    const int lane = threadIdx.x & 31, sub = threadIdx.x >> 5;
    const int row = blockIdx.x * 32 + lane;
    const bool act = row < nrows;
    float pm[10], ps[10];
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        const int c = sub + 32 * i;
        const bool ok = act && c < nparts;
        pm[i] = ok ? ws.pmax[(int64_t)c * nrows + row] : -INFINITY;
        ps[i] = ok ? ws.psum[(int64_t)c * nrows + row] : 0.f;
    }
    float mx = -INFINITY;
#pragma unroll
    for (int i = 0; i < 10; ++i) mx = fmaxf(mx, pm[i]);
    red[0][sub][lane] = mx;
    __syncthreads();
#pragma unroll 8
    for (int q = 0; q < 32; ++q) mx = fmaxf(mx, red[0][q][lane]);
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < 10; ++i) sum += pm[i] > -INFINITY ? ps[i] * exp2f(pm[i] - mx) : 0.f;
    red[1][sub][lane] = sum;
    __syncthreads();
    if (act && sub == 0) {
        float tot = 0.f;
#pragma unroll 8
        for (int q = 0; q < 32; ++q) tot += red[1][q][lane];
        ws.lse[row] = (mx + log2f(tot)) * LN2F;
    }
}

// ---------------------------------------------------------------------------------------------
// pass B (PHASE 0: likelihood + all G terms) and pass C (PHASE 1: the -p_ij c_i soft-max terms).
// D[feature, cell] = V U^T; thread <-> feature.  NB1: a single parameter row; FAST: streamlined NB
// pass B (see decoder_tc.cu).  ACC = 1: outputs on the feature side add to what a previous call
// wrote (not used yet: B <= 128).
// ---------------------------------------------------------------------------------------------
template <int FAM, int PHASE, bool GRAD, bool NB1, bool FAST>
__global__ void __launch_bounds__(NTHREADS, 1)
    main3_kernel(const __grid_constant__ CUtensorMap mapV, const __grid_constant__ CUtensorMap mapU, const P3 P) {
    static_assert(!FAST || (FAM == GLUE_NB && NB1 && PHASE == 0), "streamlined path: NB pass B");
    constexpr int EW = EPI_WARPS, EPI_T = EPI_THREADS;
    constexpr int NH = 2;   // threads sharing a feature row (splits of the cell columns)
    constexpr int DC = 32;  // columns of a 64-wide d/du accumulator per epilogue thread
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    SmemMain3& S = *reinterpret_cast<SmemMain3*>(smem_raw);
    if ((smem_u32(smem_raw) & 1023u) != 0) __trap();
    const glue_decoder_args& a = P.a;
    constexpr bool NBF = (FAM == GLUE_NB || FAM == GLUE_ZINB);
    constexpr bool ZI = (FAM == GLUE_ZINB || FAM == GLUE_ZILN);
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const int cta = blockIdx.x, ncta = gridDim.x;
    const int ntl = (P.ntiles - cta + ncta - 1) / ncta;
    const int ncell = P.ncell;
    const int nS = P.S, nunits = ntl * nS;
    const int ngrp = (ncell + 7) >> 3;
    constexpr int KRN = 1 + PHASE;
    if (t == 0) T3(KRN, 15, 0);
    if (t == 0) T3NS(KRN, 13);
    if (t == 0) T3CTA(KRN, 0);

    if (t == 0) init_bars(S, GRAD, 1);
    if (warp == TMA_WARP && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapV) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapU) : "memory");
    }
    __syncthreads();  // barriers are live: the operand loads start while the per-cell tables are filled
    if (warp == MMA_WARP) tmem_alloc(&S.tmem_base, TMEM_COLS);
    if (t < 128) {
        const bool ok = t < ncell;
        const int orig = ok ? orig_row(a, t) : 0;
        // every load first
        const int bi = (ok && a.b) ? (int)a.b[orig] : 0;
        const float li = (ok && a.l) ? a.l[orig] : 1.f;
        float lsev[MAXSEG], cv[MAXSEG], rwv[MAXSEG];
#pragma unroll
        for (int s = 0; s < MAXSEG; ++s) {
            const bool on = ok && s < nS;
            lsev[s] = (on && NBF) ? P.ws.lse[s * 128 + t] : 0.f;
            cv[s] = (on && PHASE == 1) ? P.ws.c[s * 128 + t] : 0.f;
            rwv[s] = (on && P.term[s].row_weight) ? P.term[s].row_weight[orig] : 0.f;
        }
        S.bidx[t] = bi;
        S.xrow[t] = orig;
        S.xoff32[t] = (uint32_t)((long long)orig * a.F);  // (FAST: B F < 2^31 checked by the host)
        if (t < 8) S.xoff32[128 + t] = 0u;
        S.zeros[t] = 0.f;
        S.l[t] = li;
#pragma unroll
        for (int s = 0; s < MAXSEG; ++s)
            if (s < nS) {
                S.lse[s][t] = -(lsev[s] * LOG2EF);  // -lse in base 2
                S.c[s][t] = cv[s];
                S.c_acc[s][t] = 0.f;
                const Term& T = P.term[s];
                S.wrow[s][t] = ok ? (T.row_weight ? -rwv[s] * T.gscale : T.w0) : 0.f;
            }
    }
    if (warp == TMA_WARP && lane == 0) tma_ops(S, &mapV, &mapU, P, cta, ncta, false, KRN, 0, tma_prologue_ops(nS, ntl));
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = S.tmem_base;
    if (t == 0) T3(KRN, 15, 1);

    if (warp == TMA_WARP) {
        // ===== operand loads =====
        if (lane == 0) tma_ops(S, &mapV, &mapU, P, cta, ncta, false, KRN, tma_prologue_ops(nS, ntl), ntl * (1 + nS));
    } else if (warp == MMA_WARP) {
        // ===== MMA issuer =====
        if (lane == 0) {
            auto issue_fwd = [&](int n) {
                const int it = n / nS, seg = n - it * nS, vs = it & 1, us = n & 1, as = n & 1;
                if (seg == 0) mbar_wait(&S.bars.v_full[vs], (it >> 1) & 1);
                mbar_wait(&S.bars.u_full[us], (n >> 1) & 1);
                mbar_wait(&S.bars.acc_empty[as], ((n >> 1) & 1) ^ 1);
                tc_fence_after();
                issue_logits(tmem_base + COL_ACC + as * 128, S.v[vs][0], S.v[vs][1], S.u[us][0], S.u[us][1]);
                umma_commit(&S.bars.acc_full[as]);
                T3(KRN, n, 2);
                if (!GRAD) {
                    umma_commit(&S.bars.u_empty[us]);
                    if (seg == nS - 1) umma_commit(&S.bars.v_empty[vs]);
                }
            };
            issue_fwd(0);
            const uint32_t nks1 = (uint32_t)(ncell + 15) >> 4;  // K steps over cells
            for (int n = 0; n < nunits; ++n) {
                if (n + 1 < nunits) issue_fwd(n + 1);
                if (GRAD) {
                    const int it = n / nS, seg = n - it * nS, vs = it & 1, us = n & 1;
                    const uint32_t xh = smem_u32(S.x_hi), xl = smem_u32(S.x_lo);
                    const uint32_t uh = smem_u32(S.u[us][0]), ul = smem_u32(S.u[us][1]);
                    const uint32_t vh = smem_u32(S.v[vs][0]), vl = smem_u32(S.v[vs][1]);
                    const uint32_t xs[3] = {xl, xh, xh};
                    // D1[feature, k] (+)= sum_cell X[feature, cell] U[cell, k]   (all terms of the tile);
                    // issued per block of 32 cells as the converters finish it
                    {
                        const uint32_t bs[3] = {uh, ul, uh};
                        const uint32_t d1 = tmem_base + COL_D1;
                        for (uint32_t cb = 0; cb < 4; ++cb) {
                            mbar_wait(&S.bars.x_full[cb], n & 1);
                            if (cb == 0) {
                                T3(KRN, n, 3);
                                if (seg == 0) mbar_wait(&S.bars.d1_empty, (it & 1) ^ 1);
                            }
                            tc_fence_after();
#pragma unroll
                            for (int c = 0; c < 3; ++c)
#pragma unroll
                                for (uint32_t kq = 0; kq < 2; ++kq) {
                                    const uint32_t ks = 2 * cb + kq;
                                    if (ks < nks1)
                                        umma_bf16(d1, smem_desc(DESC_K, xs[c] + (ks >> 2) * OP_BYTES + (ks & 3) * 32),
                                                  smem_desc(DESC_MN1, bs[c] + ks * 2048), IDESC_D1,
                                                  (seg | cb | c | kq) ? 1u : 0u);
                                }
                        }
                        umma_commit(&S.bars.u_empty[us]);  // (logits of this unit and the next were issued before)
                        if (seg == nS - 1) umma_commit(&S.bars.d1_full);
                    }
                    // D2[term][cell, k] += sum_feature X[feature, cell] V[feature, k]
                    {
                        const uint32_t bs[3] = {vh, vl, vh};
                        const uint32_t d2 = tmem_base + COL_D2 + 64 * seg;
#pragma unroll
                        for (int c = 0; c < 3; ++c)
#pragma unroll
                            for (uint32_t ks = 0; ks < 8; ++ks)
                                umma_bf16(d2, smem_desc(DESC_MN_X, xs[c] + ks * 2048),
                                          smem_desc(DESC_MN1, bs[c] + ks * 2048), IDESC_D2, (it | c | ks) ? 1u : 0u);
                        umma_commit(&S.bars.x_empty);
                        if (seg == nS - 1) umma_commit(&S.bars.v_empty[vs]);
                        T3(KRN, n, 4);
                    }
                }
            }
            if (GRAD) umma_commit(&S.bars.d2_full);
        }
    } else if (warp >= CONV_W0 && warp < CONV_W0 + 4) {
        // ===== converters: coefficient tile TMEM -> bf16 hi / lo shared-memory operand; d/dv write-out;
        // L2 prefetch of the observations one feature tile ahead =====
        const int cq = warp - CONV_W0;  // TMEM lane quarter
        const int ct = t - CONV_W0 * 32;
        const int frow = 32 * cq + lane;
        const uint32_t lane_addr = (uint32_t)(32 * cq) << 16;
        const bool add_v = (PHASE == 1);
        auto prefetch_x = [&](int itn) {
            if (itn >= ntl) return;
            const int f0 = (cta + itn * ncta) * 128;
            const int nvalid = min(128, a.F - f0);
            if (NB1 && ct < 16) {  // the tile's parameter rows (512 B per table)
                const float* tabs[4] = {a.scale_lin, a.bias, a.disp, a.zi_logits};
                const float* tp = tabs[ct >> 2];
                const int o = 32 * (ct & 3);
                if (tp != nullptr && o < nvalid) asm volatile("prefetch.global.L2 [%0];" ::"l"(tp + f0 + o));
            }
            if (PHASE != 0) return;
            for (int cell = ct; cell < ncell; cell += CONV_THREADS) {
                const float* px = a.x + (int64_t)S.xrow[cell] * a.F + f0;
                for (int o = 0; o < nvalid; o += 32) asm volatile("prefetch.global.L2 [%0];" ::"l"(px + o));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(px + nvalid - 1));
            }
        };
        auto readout_d1 = [&](int itp) {
            // my row of d/dv; rows leave with consecutive lanes on consecutive addresses (transposed
            // through the warp's staging block, 32 columns at a time).  Shared-memory reads are
            // collected in registers before the global stores (the compiler keeps generic-pointer
            // stores and shared loads in program order).  Pass C adds to what pass B wrote: the old
            // values of the first 32 columns are requested before the tile is waited for.
            const bool tr0 = ct == 0 && itp == 0;
            if (tr0) T3(KRN, 15, 6);
            const int j = (cta + itp * ncta) * 128 + frow;
            int myrow = -1;
            if (j < a.F && a.grad_v) myrow = (int)(a.vidx ? a.vidx[j] : (long long)j);
            S.rowidx[cq][lane] = myrow;
            __syncwarp();
            const int* rowidx = S.rowidx[cq];
            float(*st)[33] = S.stage[cq];
            float old[32];
            // (row indices are re-read from shared memory where they are needed: at most 64 values per
            // thread are live at any point, nothing is demoted to local memory)
            auto load_old = [&](int kk) {
#pragma unroll
                for (int r = 0; r < 32; ++r) {
                    const int row = rowidx[r];
                    old[r] = (add_v && kk < a.K && row >= 0) ? __ldcg(a.grad_v + (long long)row * a.K + kk) : 0.f;
                }
            };
            auto store_rows = [&](int kk) {
                float vals[32];
                int rows[32];
#pragma unroll
                for (int r = 0; r < 32; ++r) {
                    rows[r] = rowidx[r];
                    vals[r] = st[r][lane] + old[r];
                }
                asm volatile("" ::: "memory");  // every shared-memory read before the first global store
                const bool kok = kk < a.K;
#pragma unroll
                for (int r = 0; r < 32; ++r)
                    st_global_pred(a.grad_v + (long long)rows[r] * a.K + kk, vals[r], kok && rows[r] >= 0);
            };
            load_old(lane);
            if (tr0) T3(KRN, 15, 7);
            mbar_wait(&S.bars.d1_full, itp & 1);
            tc_fence_after();
            if (tr0) T3(KRN, 15, 8);
            if (ct == 0) T3(KRN, itp * nS + nS - 1, 13);
            {
                uint32_t r0[32];
                tmem_ld32(tmem_base + lane_addr + COL_D1, r0);
                tmem_ld_wait();
                if (tr0) T3(KRN, 15, 9);
#pragma unroll
                for (int k = 0; k < 32; ++k) st[lane][k] = __uint_as_float(r0[k]);
            }
            __syncwarp();
            if (tr0) T3(KRN, 15, 10);
            store_rows(lane);
            __syncwarp();
            if (tr0) T3(KRN, 15, 11);
            if (a.K > 32) {
                load_old(32 + lane);
                uint32_t r1[32];
                tmem_ld32(tmem_base + lane_addr + COL_D1 + 32, r1);
                tmem_ld_wait();
#pragma unroll
                for (int k = 0; k < 32; ++k) st[lane][k] = __uint_as_float(r1[k]);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&S.bars.d1_empty);
            if (tr0) T3(KRN, 15, 12);
            if (a.K > 32) {
                store_rows(32 + lane);
                __syncwarp();
            }
            if (ct == 0) T3(KRN, itp * nS + nS - 1, 14);
        };
        if (GRAD) {
            prefetch_x(0);
            for (int n = 0; n < nunits; ++n) {
                const int it = n / nS, seg = n - it * nS, as = n & 1;
                if (seg == 0) prefetch_x(it + 1);
                mbar_wait(&S.bars.coef_full[as], (n >> 1) & 1);
                tc_fence_after();
                if (ct == 0) T3(KRN, n, 10);
                mbar_wait(&S.bars.x_empty, (n & 1) ^ 1);  // gradient products of the previous unit are done with X
                if (ct == 0) T3(KRN, n, 11);
#pragma unroll 1
                for (int cb = 0; cb < 4; ++cb) {
                    uint32_t d[32];
                    tmem_ld32(tmem_base + lane_addr + COL_ACC + as * 128 + 32 * cb, d);
                    tmem_ld_wait();
#pragma unroll
                    for (int g8 = 0; g8 < 4; ++g8) {
                        const int c0 = 32 * cb + 8 * g8;
                        uint32_t hh[4], ll[4];
                        if (c0 < 8 * ngrp) {
#pragma unroll
                            for (int e2 = 0; e2 < 4; ++e2)
                                split2(__uint_as_float(d[8 * g8 + 2 * e2]), __uint_as_float(d[8 * g8 + 2 * e2 + 1]),
                                       hh[e2], ll[e2]);
                        } else {  // cell groups no epilogue thread visits: exact zeros
#pragma unroll
                            for (int e2 = 0; e2 < 4; ++e2) hh[e2] = 0u, ll[e2] = 0u;
                        }
                        const uint32_t off =
                            (uint32_t)(c0 >> 6) * OP_BYTES + swz((uint32_t)frow, (uint32_t)((c0 & 63) >> 3));
                        st_shared_v4(smem_u32(S.x_hi) + off, hh[0], hh[1], hh[2], hh[3]);
                        st_shared_v4(smem_u32(S.x_lo) + off, ll[0], ll[1], ll[2], ll[3]);
                    }
                    fence_proxy_async();
                    if (cb == 3) tc_fence_before();
                    __syncwarp();
                    if (lane == 0) {
                        if (cb == 3) mbar_arrive(&S.bars.acc_empty[as]);
                        mbar_arrive(&S.bars.x_full[cb]);
                    }
                }
                if (ct == 0) T3(KRN, n, 12);
                if (seg == 0 && it > 0) readout_d1(it - 1);  // (its products were committed a unit ago)
            }
            readout_d1(ntl - 1);
        } else {
            for (int it = 0; it < ntl; ++it) prefetch_x(it);
        }
    } else if (warp < EW) {
        // ===== epilogue =====
        const int q = warp & 3, h = warp >> 2;
        const int frow = 32 * q + lane;
        const uint32_t lane_addr = (uint32_t)(32 * q) << 16;
        // groups of 8 cells; the NH threads that share a feature split the groups in contiguous runs
        const int g_begin = (ngrp * h + NH - 1) / NH;
        const int g_end = (ngrp * (h + 1) + NH - 1) / NH;
        const bool skip_nan = a.reduce == 0 && P.term[0].row_weight == nullptr;  // nanmean semantics (single term)
        long long t_loop = 0, t_wait = 0;
        float loss_seg[MAXSEG];
        int bad_seg[MAXSEG];
#pragma unroll
        for (int s = 0; s < MAXSEG; ++s) loss_seg[s] = 0.f, bad_seg[s] = 0;

        // per-tile state (NB1: parameters and their gradient sums live across the units of a tile)
        int j = 0;
        bool jok = false;
        int cur_b = -1;
        FeatParams fp = {};
        float gb = 0.f, gs = 0.f, gd = 0.f, gz = 0.f;
        bool pend = false;
        int64_t pend_off = 0;
        float pend_b = 0.f, pend_s = 0.f, pend_d = 0.f, pend_z = 0.f;
        float s2 = 0.f, b2 = 0.f, lg1 = 0.f, dg1 = 0.f, lg2 = 0.f, dg2 = 0.f, lg3 = 0.f, dg3 = 0.f;
        f2 ngb2 = pk(0.f, 0.f), ngs2 = pk(0.f, 0.f), gd2 = pk(0.f, 0.f);

        auto write_param = [&](int64_t off, float vb, float vs, float vd, float vz, bool add) {
            const bool wd = PHASE == 0 && a.grad_disp, wz = PHASE == 0 && ZI && a.grad_zi_logits;
            float ob = 0.f, os = 0.f, od = 0.f, oz = 0.f;
            if (add) {  // every load before the first store
                if (a.grad_bias) ob = a.grad_bias[off];
                if (a.grad_scale_lin) os = a.grad_scale_lin[off];
                if (wd) od = a.grad_disp[off];
                if (wz) oz = a.grad_zi_logits[off];
            }
            if (a.grad_bias) a.grad_bias[off] = ob + vb;
            if (a.grad_scale_lin) a.grad_scale_lin[off] = os + vs;
            if (wd) a.grad_disp[off] = od + vd;
            if (wz) a.grad_zi_logits[off] = oz + vz;
        };
        // several parameter rows: read-modify-write per run of cells with one batch index; the first run
        // of the upper-half thread is written after a barrier (fixed order of the two updates of an entry)
        auto flush = [&]() {
            if (GRAD && cur_b >= 0 && jok) {
                const int64_t off = (int64_t)cur_b * a.F + j;
                const float vs = gs * fp.sg;
                if (h == 1 && !pend) {
                    pend = true, pend_off = off, pend_b = gb, pend_s = vs, pend_d = gd, pend_z = gz;
                } else {
                    write_param(off, gb, vs, gd, gz, true);
                }
            }
            gb = gs = gd = gz = 0.f;
        };
        // single parameter row: the two threads of a feature meet in shared memory, the lower-half thread
        // writes (pass C adds to what pass B wrote; those values were requested when the tile started)
        float old_b = 0.f, old_s = 0.f;
        auto flush_nb1 = [&]() {
            const float vs = gs * fp.sg;
            if (h == 1) {
                S.pgrad[0][frow] = gb, S.pgrad[1][frow] = vs, S.pgrad[2][frow] = gd, S.pgrad[3][frow] = gz;
            }
            named_bar_sync(4 + q, 64);  // the two warps of this TMEM lane quarter
            if (h == 0 && jok) {
                const int64_t off = j;
                if (a.grad_bias) a.grad_bias[off] = (gb + S.pgrad[0][frow]) + old_b;
                if (a.grad_scale_lin) a.grad_scale_lin[off] = (vs + S.pgrad[1][frow]) + old_s;
                if (PHASE == 0 && a.grad_disp) a.grad_disp[off] = gd + S.pgrad[2][frow];
                if (PHASE == 0 && ZI && a.grad_zi_logits) a.grad_zi_logits[off] = gz + S.pgrad[3][frow];
            }
            named_bar_sync(4 + q, 64);  // (pgrad is free again)
            gb = gs = gd = gz = 0.f;
        };

        for (int n = 0; n < nunits; ++n) {
            const int it = n / nS, seg = n - it * nS, as = n & 1;
            const int f0 = (cta + it * ncta) * 128;
            if (seg == 0 || !NB1) {
                j = f0 + frow;
                jok = j < a.F;
                cur_b = -1;
                pend = false;
                gb = gs = gd = gz = 0.f;
                ngb2 = pk(0.f, 0.f), ngs2 = pk(0.f, 0.f), gd2 = pk(0.f, 0.f);
                if (NB1) {
                    if (GRAD && PHASE == 1 && h == 0 && jok) {
                        old_b = a.grad_bias ? __ldcg(a.grad_bias + j) : 0.f;
                        old_s = a.grad_scale_lin ? __ldcg(a.grad_scale_lin + j) : 0.f;
                    }
                    fp = load_params<FAM>(a, 0, j, jok);
                    cur_b = 0;
                    s2 = fp.s * LOG2EF, b2 = fp.bias * LOG2EF;  // logits in base 2
                    if (PHASE == 1 && !jok) b2 = -INFINITY;     // features beyond F: p = 0
                    if (NBF && PHASE == 0) {  // log-gamma / digamma differences for counts 1 and 2
                        const float ti = 1.f / fp.th;
                        lg1 = fp.disp, dg1 = ti;
                        lg2 = logf(fp.th * (fp.th + 1.f)) - 0.69314718055994530942f;
                        dg2 = ti + 1.f / (fp.th + 1.f);
                        lg3 = lg2 + logf(fp.th + 2.f) - 1.09861228866810969140f;  // - log 3
                        dg3 = dg2 + 1.f / (fp.th + 2.f);
                    }
                }
            }
            float loss_acc = 0.f;
            int bad_acc = 0;
            f2 loss2 = pk(0.f, 0.f);

            // observations of the first group are requested before the logits are waited for
            float xn[8];
            const float* xj = a.x + (jok ? j : a.F - 1);
            if (FAST) asm volatile("" : "+l"(xj));  // keep the column base in a register pair
            auto load_x = [&](int g) {
                if constexpr (FAST) {
                    // every load in bounds without predicates: column clamped to F - 1, cells beyond
                    // ncell (and the group after the last one) read row offset 0; weight 0 either way
                    const uint4 o0 = *reinterpret_cast<const uint4*>(&S.xoff32[8 * g]);
                    const uint4 o1 = *reinterpret_cast<const uint4*>(&S.xoff32[8 * g + 4]);
                    xn[0] = __ldg(xj + o0.x), xn[1] = __ldg(xj + o0.y), xn[2] = __ldg(xj + o0.z), xn[3] = __ldg(xj + o0.w);
                    xn[4] = __ldg(xj + o1.x), xn[5] = __ldg(xj + o1.y), xn[6] = __ldg(xj + o1.z), xn[7] = __ldg(xj + o1.w);
                } else {
#pragma unroll
                    for (int e8 = 0; e8 < 8; ++e8) {
                        const int cell = 8 * g + e8;
                        xn[e8] = (PHASE == 0 && jok && g < g_end && cell < ncell)
                                     ? __ldg(a.x + (int64_t)S.xrow[cell] * a.F + j) : 0.f;
                    }
                }
            };
            const f2 S2 = pk(s2, s2), B2 = pk(b2, b2), NS2 = pk(-fp.s, -fp.s), TH2 = pk(fp.th, fp.th);
            const f2 EPS2 = pk(GLUE_EPS, GLUE_EPS), M1 = pk(-1.f, -1.f), LN2_2 = pk(LN2F, LN2F);
            const f2 NLN2 = pk(-LN2F, -LN2F), DISP2 = pk(fp.disp, fp.disp);
            // FAST: log-gamma / digamma differences of the counts 0 .. 3 as the cubic x (c1 + c2 x + c3 x^2)
            // through their exact values (sparse count data: 3 is still common, 4 and more is rare)
            const float ld1 = lg1, ld2 = 0.5f * lg2, ld3 = lg3 * (1.f / 3.f);
            const float lq_c3 = 0.5f * ((ld3 - ld2) - (ld2 - ld1)), lq_c2 = (ld2 - ld1) - 3.f * lq_c3;
            const float lq_c1 = ld1 - lq_c2 - lq_c3;
            const float dd1 = dg1, dd2 = 0.5f * dg2, dd3 = dg3 * (1.f / 3.f);
            const float dq_c3 = 0.5f * ((dd3 - dd2) - (dd2 - dd1)), dq_c2 = (dd2 - dd1) - 3.f * dq_c3;
            const float dq_c1 = dd1 - dq_c2 - dq_c3;
            const f2 LQ1 = pk(lq_c1, lq_c1), LQ2 = pk(lq_c2, lq_c2), LQ3 = pk(lq_c3, lq_c3);
            const f2 DQ1 = pk(dq_c1, dq_c1), DQ2 = pk(dq_c2, dq_c2), DQ3 = pk(dq_c3, dq_c3);
            const f2 M2 = pk(-2.f, -2.f), M3 = pk(-3.f, -3.f);
            const float* wsrc = (FAST && !jok) ? S.zeros : S.wrow[seg];
            const float* lsep = S.lse[seg];
            const float* auxp = PHASE == 0 ? S.l : S.c[seg];
            load_x(g_begin);
            if (t == 0) T3(KRN, n, 5);
            const long long tw0 = clock64();
            mbar_wait(&S.bars.acc_full[as], (n >> 1) & 1);
            tc_fence_after();
            const long long tw1 = clock64();
            t_wait += tw1 - tw0;
            if (t == 0) T3(KRN, n, 6);
#pragma unroll 1
            for (int g = g_begin; g < g_end; ++g) {
                const int c0 = g * 8;
                uint32_t d[8];
                tmem_ld8(tmem_base + lane_addr + COL_ACC + as * 128 + c0, d);
                float lsev[8], auxv[8], wv[8];  // aux = library size (pass B) or c_i (pass C)
                int bv[8];
                if (PHASE == 0) {
                    const float4 w0 = *reinterpret_cast<const float4*>(&wsrc[c0]);
                    const float4 w1 = *reinterpret_cast<const float4*>(&wsrc[c0 + 4]);
                    wv[0] = w0.x, wv[1] = w0.y, wv[2] = w0.z, wv[3] = w0.w;
                    wv[4] = w1.x, wv[5] = w1.y, wv[6] = w1.z, wv[7] = w1.w;
                }
                {
                    const float4 l0 = *reinterpret_cast<const float4*>(&lsep[c0]);
                    const float4 l1 = *reinterpret_cast<const float4*>(&lsep[c0 + 4]);
                    const float4 a0 = *reinterpret_cast<const float4*>(&auxp[c0]);
                    const float4 a1 = *reinterpret_cast<const float4*>(&auxp[c0 + 4]);
                    lsev[0] = l0.x, lsev[1] = l0.y, lsev[2] = l0.z, lsev[3] = l0.w;
                    lsev[4] = l1.x, lsev[5] = l1.y, lsev[6] = l1.z, lsev[7] = l1.w;
                    auxv[0] = a0.x, auxv[1] = a0.y, auxv[2] = a0.z, auxv[3] = a0.w;
                    auxv[4] = a1.x, auxv[5] = a1.y, auxv[6] = a1.z, auxv[7] = a1.w;
                    if (!NB1) {
                        const int4 b0 = *reinterpret_cast<const int4*>(&S.bidx[c0]);
                        const int4 b1 = *reinterpret_cast<const int4*>(&S.bidx[c0 + 4]);
                        bv[0] = b0.x, bv[1] = b0.y, bv[2] = b0.z, bv[3] = b0.w;
                        bv[4] = b1.x, bv[5] = b1.y, bv[6] = b1.z, bv[7] = b1.w;
                    }
                }
                float xv[8];
#pragma unroll
                for (int e8 = 0; e8 < 8; ++e8) xv[e8] = xn[e8];
                load_x(g + 1);
                tmem_ld_wait();
                float coef[8], gch[8];
                if constexpr (FAST) {
                    // ---- pass B, NB, streamlined: counts outside {0, 1, 2} and NaN are dealt with per
                    // thread first (their gamma terms go straight to the accumulators, their quadratic
                    // argument becomes 0), then packed pairs of cells without any selects
                    float xq[8];
                    uint32_t rb = 0u;
#pragma unroll
                    for (int pr_ = 0; pr_ < 4; ++pr_) {
                        const f2 x2 = pk(xv[2 * pr_], xv[2 * pr_ + 1]);
                        float p0, p1;
                        up(mul2(mul2(x2, add2(x2, M1)), mul2(add2(x2, M2), add2(x2, M3))), p0, p1);  // x (x-1) (x-2) (x-3)
                        rb |= __float_as_uint(p0) | __float_as_uint(p1);
                        xq[2 * pr_] = xv[2 * pr_], xq[2 * pr_ + 1] = xv[2 * pr_ + 1];
                    }
                    if ((rb & 0x7fffffffu) != 0u) {
#pragma unroll
                        for (int e8 = 0; e8 < 8; ++e8) {
                            const float x = xv[e8];
                            if (skip_nan && x != x) {
                                bad_acc += (jok && c0 + e8 < ncell) ? 1 : 0;
                                xv[e8] = 0.f, xq[e8] = 0.f, wv[e8] = 0.f;
                            } else if (!(x == 0.f || x == 1.f || x == 2.f || x == 3.f)) {
                                float lg_t, dg_t;
                                nb_gamma_terms_t<true>(x, fp.th, lg_t, dg_t);
                                xq[e8] = 0.f;
                                loss_acc = fmaf(wv[e8], lg_t, loss_acc);
                                if (GRAD) gd = fmaf(wv[e8] * fp.th, dg_t, gd);
                            }
                        }
                    }
#pragma unroll
                    for (int pr_ = 0; pr_ < 4; ++pr_) {
                        const int e0 = 2 * pr_, e1 = e0 + 1;
                        const f2 x2 = pk(xv[e0], xv[e1]), xq2 = pk(xq[e0], xq[e1]), W2 = pk(wv[e0], wv[e1]);
                        const f2 dot2 = pk(__uint_as_float(d[e0]), __uint_as_float(d[e1]));
                        const f2 t2 = add2(fma2(S2, dot2, B2), pk(lsev[e0], lsev[e1]));
                        float t0, t1;
                        up(t2, t0, t1);
                        const f2 mu2 = mul2(pk(ex2f(t0), ex2f(t1)), pk(auxv[e0], auxv[e1]));
                        const f2 m2 = add2(mu2, EPS2), mt2 = add2(m2, TH2);
                        float m0, m1, mt0, mt1, mm0, mm1;
                        up(m2, m0, m1);
                        up(mt2, mt0, mt1);
                        const f2 l2m = pk(lg2f(m0), lg2f(m1)), l2t = pk(lg2f(mt0), lg2f(mt1));
                        up(mul2(m2, mt2), mm0, mm1);
                        const f2 r2 = pk(rcpf(mm0), rcpf(mm1));  // 1/m = r mt, 1/(m + theta) = r m
                        const f2 sig2 = mul2(mul2(m2, m2), r2);
                        const f2 dlt2 = fma2(l2t, NLN2, DISP2);                        // log theta - log(m + theta)
                        const f2 xl2 = mul2(fma2(l2t, M1, l2m), LN2_2);                // log m - log(m + theta)
                        const f2 ndd2 = fma2(add2(x2, TH2), sig2, mul2(x2, M1));       // -(x - (x + theta) sig)
                        const f2 lgq2 = mul2(xq2, fma2(fma2(LQ3, xq2, LQ2), xq2, LQ1));
                        const f2 lp2 = fma2(TH2, dlt2, fma2(x2, xl2, lgq2));
                        loss2 = fma2(W2, lp2, loss2);
                        if (GRAD) {
                            const f2 rat2 = mul2(mu2, mul2(r2, mt2));
                            const f2 dgq2 = mul2(xq2, fma2(fma2(DQ3, xq2, DQ2), xq2, DQ1));
                            const f2 tl2 = fma2(TH2, add2(dlt2, dgq2), ndd2);
                            const f2 nG2 = mul2(W2, mul2(ndd2, rat2));
                            ngb2 = add2(ngb2, nG2);
                            ngs2 = fma2(nG2, dot2, ngs2);
                            gd2 = fma2(W2, tl2, gd2);
                            up(mul2(nG2, NS2), coef[e0], coef[e1]);
                        } else {
                            coef[e0] = coef[e1] = 0.f;
                        }
                        gch[e0] = gch[e1] = 0.f;
                    }
                } else if constexpr (NB1 && PHASE == 1) {
                    // ---- pass C: packed pairs of cells, -G = p c accumulated (signs fixed at the flush)
#pragma unroll
                    for (int pr_ = 0; pr_ < 4; ++pr_) {
                        const int e0 = 2 * pr_, e1 = e0 + 1;
                        const f2 dot2 = pk(__uint_as_float(d[e0]), __uint_as_float(d[e1]));
                        const f2 t2 = add2(fma2(S2, dot2, B2), pk(lsev[e0], lsev[e1]));  // b2 = -inf kills !jok
                        float t0, t1;
                        up(t2, t0, t1);
                        const f2 nG2 = mul2(pk(ex2f(t0), ex2f(t1)), pk(auxv[e0], auxv[e1]));  // c_i = 0 beyond ncell
                        ngb2 = add2(ngb2, nG2);
                        ngs2 = fma2(nG2, dot2, ngs2);
                        up(mul2(nG2, NS2), coef[e0], coef[e1]);
                        gch[e0] = gch[e1] = 0.f;
                    }
                } else if constexpr (NB1 && NBF) {
                    // ---- pass B, NB / ZINB, single parameter row, straight-line in three stages ------
                    float xs[8], dlt[8], xl[8], dd[8], ratio[8], lgv[8], dgv[8], Wv[8];
#pragma unroll
                    for (int e8 = 0; e8 < 8; ++e8) {
                        const float dot = __uint_as_float(d[e8]);
                        const float x = xv[e8];
                        const bool isnan_x = x != x;
                        bool valid = jok && (c0 + e8 < ncell);
                        if (skip_nan && isnan_x) {
                            bad_acc += valid ? 1 : 0;
                            valid = false;
                        }
                        xs[e8] = (skip_nan && isnan_x) ? 0.f : x;
                        Wv[e8] = valid ? wv[e8] : 0.f;
                        const float pr = ex2f(fmaf(s2, dot, b2) + lsev[e8]);
                        const float mu = pr * auxv[e8];
                        const float m = mu + GLUE_EPS, mt = m + fp.th;
                        const float Lm = lg2f(m) * LN2F, Lt = lg2f(mt) * LN2F;
                        const float r = rcpf(m * mt);  // 1/m = r mt, 1/(m + theta) = r m
                        const float sig = (m * m) * r;
                        ratio[e8] = mu * (r * mt);
                        dlt[e8] = fp.disp - Lt;
                        xl[e8] = Lm - Lt;
                        dd[e8] = xs[e8] - (xs[e8] + fp.th) * sig;
                        lgv[e8] = xs[e8] == 1.f ? lg1 : (xs[e8] == 2.f ? lg2 : 0.f);
                        dgv[e8] = xs[e8] == 1.f ? dg1 : (xs[e8] == 2.f ? dg2 : 0.f);
                    }
#pragma unroll
                    for (int e8 = 0; e8 < 8; ++e8)  // counts other than 0, 1, 2: rare
                        if (!(xs[e8] == 0.f || xs[e8] == 1.f || xs[e8] == 2.f)) {
                            float lg_t, dg_t;
                            nb_gamma_terms_t<true>(xs[e8], fp.th, lg_t, dg_t);
                            lgv[e8] = lg_t, dgv[e8] = dg_t;
                        }
#pragma unroll
                    for (int e8 = 0; e8 < 8; ++e8) {
                        const float dot = __uint_as_float(d[e8]);
                        float lp = fmaf(fp.th, dlt[e8], fmaf(xs[e8], xl[e8], lgv[e8]));
                        const float tl = fmaf(fp.th, dlt[e8] + dgv[e8], -dd[e8]);
                        float wnb = 1.f, gzv = 0.f;
                        if (FAM == GLUE_ZINB) {
                            const float e1 = ex2f(lp * LOG2EF);
                            const float E = e1 + fp.ezi + GLUE_EPS;
                            const float rE = rcpf(E);
                            const bool isz = fabsf(xs[e8]) < GLUE_EPS;
                            lp = isz ? lg2f(E) * LN2F - fp.spz : lp - fp.spz;
                            wnb = isz ? e1 * rE : 1.f;
                            gzv = isz ? fp.ezi * rE - fp.sgz : -fp.sgz;
                        }
                        const float W = Wv[e8];
                        loss_acc = fmaf(W, lp, loss_acc);
                        float G = 0.f;
                        if (GRAD) {
                            G = W * (wnb * dd[e8] * ratio[e8]);
                            gb += G;
                            gs = fmaf(G, dot, gs);
                            gd = fmaf(W, wnb * tl, gd);
                            gz = fmaf(W, gzv, gz);
                        }
                        coef[e8] = G * fp.s;
                        gch[e8] = G;
                    }
                } else {
                    // ---- generic path: other families, several parameter rows ---------------------
#pragma unroll
                    for (int e8 = 0; e8 < 8; ++e8) {
                        const int cell = c0 + e8;
                        const bool cok = cell < ncell;  // warp-uniform
                        float G = 0.f, cf = 0.f;
                        if (NB1 || cok) {
                            if (!NB1 && bv[e8] != cur_b) {
                                flush();
                                fp = load_params<FAM>(a, bv[e8], j, jok);
                                cur_b = bv[e8];
                            }
                            const float dot = __uint_as_float(d[e8]);
                            const float av = fmaf(fp.s, dot, fp.bias);
                            if (PHASE == 0) {
                                const float x = xv[e8];
                                const float lse_nat = NBF ? -(lsev[e8] * LN2F) : 0.f;
                                const ElemOut e = element<FAM, true>(fp, av, x, lse_nat, auxv[e8]);
                                bool valid = jok && cok;
                                if (skip_nan && x != x) {
                                    bad_acc += valid ? 1 : 0;
                                    valid = false;
                                }
                                const float W = valid ? wv[e8] : 0.f;
                                loss_acc += valid ? W * e.lp : 0.f;
                                if (GRAD) {
                                    G = valid ? W * e.g : 0.f;
                                    cf = G * fp.s;
                                    gb += G;
                                    gs += G * dot;
                                    gd += valid ? W * e.gd : 0.f;
                                    gz += valid ? W * e.gz : 0.f;
                                }
                            } else {
                                const float pr = (jok && cok) ? ex2f(fmaf(av, LOG2EF, lsev[e8])) : 0.f;
                                G = -(pr * auxv[e8]);
                                cf = G * fp.s;
                                gb += G;
                                gs += G * dot;
                            }
                        }
                        coef[e8] = cf;
                        gch[e8] = G;
                    }
                }
                if (GRAD) {
                    // coefficient tile X[feature, cell] = dL/da softplus(scale): back into the logits
                    // accumulator, same lanes and columns this thread just read
                    uint32_t cu[8];
#pragma unroll
                    for (int e8 = 0; e8 < 8; ++e8) cu[e8] = __float_as_uint(coef[e8]);
                    tmem_st8(tmem_base + lane_addr + COL_ACC + as * 128 + c0, cu);
                    if (NBF && PHASE == 0 && !(FAST || (NB1 && P.cinv))) {
                        // c_i partial: transpose-reduce the (32 feature lanes x 8 cells) block over lanes;
                        // afterwards lane L holds the warp total of cell c0 + 4 bit4(L) + 2 bit3(L) + bit2(L)
#pragma unroll
                        for (int off2 = 16, nn = 4; nn >= 1; off2 >>= 1, nn >>= 1) {
                            const bool upper = (lane & off2) != 0;
#pragma unroll
                            for (int k = 0; k < nn; ++k) {
                                const float send = upper ? gch[k] : gch[k + nn];
                                const float keep = upper ? gch[k + nn] : gch[k];
                                gch[k] = keep + __shfl_xor_sync(FULL, send, off2);
                            }
                        }
                        float tot = gch[0];
                        tot += __shfl_xor_sync(FULL, tot, 2);
                        tot += __shfl_xor_sync(FULL, tot, 1);
                        if ((lane & 3) == 0) S.cw[n & 1][q][c0 + (lane >> 2)] = tot;
                    }
                }
            }
            t_loop += clock64() - tw1;
            if (t == 0) T3(KRN, n, 7);
            {  // loss of the unit
                float a0, a1;
                up(loss2, a0, a1);
                loss_acc += a0 + a1;
#pragma unroll
                for (int s = 0; s < MAXSEG; ++s)
                    if (s == seg) loss_seg[s] += loss_acc, bad_seg[s] += bad_acc;
            }
            if (GRAD) {
                tmem_st_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&S.bars.coef_full[as]);
            } else {
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&S.bars.acc_empty[as]);
            }
            if (t == 0) T3(KRN, n, 8);
            const bool last_of_tile = seg == nS - 1;
            if (GRAD && NB1 && last_of_tile) {
                float a0, a1;  // packed accumulators of the fast paths (they hold -G sums)
                up(ngb2, a0, a1);
                gb -= a0 + a1;
                up(ngs2, a0, a1);
                gs -= a0 + a1;
                up(gd2, a0, a1);
                gd += a0 + a1;
                flush_nb1();
            }
            if (GRAD && !NB1) {
                flush();
                named_bar_sync(1, EPI_T);
                if (pend) write_param(pend_off, pend_b, pend_s, pend_d, pend_z, true);
                pend = false;
                // the next unit of this tile updates the same table entries
                if (nS > 1) named_bar_sync(1, EPI_T);
            }
            if (GRAD && NBF && PHASE == 0 && !(FAST || (NB1 && P.cinv))) {
                if (NB1) named_bar_sync(1, EPI_T);  // (cw of this unit is complete)
                if (t < ncell)
                    S.c_acc[seg][t] += (S.cw[n & 1][0][t] + S.cw[n & 1][1][t]) + (S.cw[n & 1][2][t] + S.cw[n & 1][3][t]);
            }
        }
        if (t == 0) T3(KRN, 15, 2);
        if (t == 0) {
            unsigned smid;
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            T3EPI(KRN, 0, smid);
            T3EPI(KRN, 1, t_loop);
            T3EPI(KRN, 2, t_wait);
        }
        if (GRAD) {
            // d/du partials of this CTA: rows = cells, one 64-column accumulator per term
            mbar_wait(&S.bars.d2_full, 0);
            tc_fence_after();
            if (t == 0) T3(KRN, 15, 3);
            for (int s = 0; s < nS; ++s) {
                uint32_t r[DC];
                tmem_ld32(tmem_base + lane_addr + COL_D2 + 64 * s + DC * h, r);
                tmem_ld_wait();
                float4* dst =
                    reinterpret_cast<float4*>(P.ws.apart + (((int64_t)cta * nS + s) * 128 + frow) * 64 + DC * h);
                float4 old[DC / 4];
#pragma unroll
                for (int k4 = 0; k4 < DC / 4; ++k4) old[k4] = PHASE == 1 ? dst[k4] : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int k4 = 0; k4 < DC / 4; ++k4)
                    dst[k4] = make_float4(__uint_as_float(r[4 * k4]) + old[k4].x, __uint_as_float(r[4 * k4 + 1]) + old[k4].y,
                                          __uint_as_float(r[4 * k4 + 2]) + old[k4].z,
                                          __uint_as_float(r[4 * k4 + 3]) + old[k4].w);
                if (NBF && PHASE == 0) {
                    if (FAST || (NB1 && P.cinv)) {
                        // c_i partial = column K of the d/du accumulator (v carries 1 / softplus(scale) there)
                        float cval = 0.f;
#pragma unroll
                        for (int k = 0; k < DC; ++k)
                            if (DC * h + k == a.K) cval = __uint_as_float(r[k]);
                        if (a.K / DC == h && frow < ncell) P.ws.cpart[(int64_t)cta * (nS * 128) + s * 128 + frow] = cval;
                    }
                }
            }
            if (NBF && PHASE == 0 && !(FAST || (NB1 && P.cinv))) {
                named_bar_sync(1, EPI_T);
                for (int s = 0; s < nS; ++s)
                    if (t < ncell) P.ws.cpart[(int64_t)cta * (nS * 128) + s * 128 + t] = S.c_acc[s][t];
            }
        }
        if (PHASE == 0) {
#pragma unroll
            for (int s = 0; s < MAXSEG; ++s) {
                const float wl = warp_sum(loss_seg[s]);
                int wb = bad_seg[s];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) wb += __shfl_xor_sync(FULL, wb, o);
                if (lane == 0) S.red_l[s][warp] = wl, S.red_b[s][warp] = wb;
            }
            named_bar_sync(2, EPI_T);
            if (t < nS) {
                float sl = 0.f;
                int sb = 0;
                for (int w = 0; w < EW; ++w) sl += S.red_l[t][w], sb += S.red_b[t][w];
                P.ws.losspart[cta * nS + t] = sl;
                P.ws.badpart[cta * nS + t] = sb;
            }
        }
    }
    tc_fence_before();
    if (PHASE == 0) __threadfence();
    __syncthreads();
    if (t == 0) T3(KRN, 15, 4);
    if (warp == MMA_WARP) tmem_dealloc(tmem_base, TMEM_COLS);
    if (PHASE == 0) {
        // ---- the last CTA to arrive reduces the per-CTA partials: c_i, losses, nanmean factor -----
        if (t == 0) S.ticket = atomicAdd(P.ws.counters + 1, 1u);
        __syncthreads();
        if (S.ticket == gridDim.x - 1) {
            __threadfence();
            const int nc = (int)gridDim.x;
            float* cpart3 = &S.stage[0][0][0];  // [3][128] (the write-out blocks are idle by now)
            for (int s = 0; s < nS; ++s) {
                if (GRAD && NBF && t < 384) {  // three strided sub-sums per cell, combined in a fixed order
                    const int cell = t & 127, sub = t >> 7;
                    float c = 0.f;
                    if (cell < ncell) {
#pragma unroll 10
                        for (int qq = sub; qq < nc; qq += 3)
                            c += __ldcg(P.ws.cpart + (int64_t)qq * (nS * 128) + s * 128 + cell);
                    }
                    cpart3[sub * 128 + cell] = c;
                }
                __syncthreads();
                if (GRAD && NBF && t < ncell) P.ws.c[s * 128 + t] = (cpart3[t] + cpart3[128 + t]) + cpart3[256 + t];
                __syncthreads();
            }
            if (warp == MMA_WARP) {
                for (int s = 0; s < nS; ++s) {
                    double loss = 0.0;
                    long long bad = 0;
                    for (int qq = lane; qq < nc; qq += 32) {
                        loss += (double)__ldcg(P.ws.losspart + qq * nS + s);
                        bad += __ldcg(P.ws.badpart + qq * nS + s);
                    }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        loss += __shfl_xor_sync(FULL, loss, o);
                        bad += __shfl_xor_sync(FULL, bad, o);
                    }
                    if (lane == 0) {
                        const double total = (double)a.B * (double)a.F;
                        const double ratio = bad > 0 ? total / (total - (double)bad) : 1.0;
                        if (s == 0) P.ws.rescale[0] = (float)ratio;
                        if (P.term[s].loss) P.term[s].loss[0] = (float)(loss * ratio / (double)P.term[s].gscale);
                    }
                }
            }
        }
    }
    if (t == 0) T3(KRN, 15, 5);
    if (t == 0) T3NS(KRN, 14);
    if (t == 0) T3CTA(KRN, 1);
}

// d/du = sum over CTAs of the D2 partials; grid = S * ncell blocks, 1024 threads = 64 latent columns x
// 16 strided sub-sums (<= 10 loads in flight per thread) combined in a fixed order
__global__ void __launch_bounds__(1024) grad_u3_kernel(P3 P, int ncta) {
    const glue_decoder_args& a = P.a;
    const int s = blockIdx.x / P.ncell, pos = blockIdx.x - s * P.ncell;
    const int k = threadIdx.x & 63, sub = threadIdx.x >> 6;
    __shared__ float part[16][64];
    float vals[10];
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        const int q = sub + 16 * i;
        vals[i] = q < ncta ? __ldcg(P.ws.apart + (((int64_t)q * P.S + s) * 128 + pos) * 64 + k) : 0.f;
    }
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < 10; ++i) acc += vals[i];
    part[sub][k] = acc;
    __syncthreads();
    float* gu = P.term[s].grad_u;
    if (sub == 0 && k < a.K && gu) {
        float tot = 0.f;
#pragma unroll
        for (int q = 0; q < 16; ++q) tot += part[q][k];
        gu[(int64_t)orig_row(a, pos) * a.K + k] = tot;
    }
}

// only does work when NaN observations were skipped (nanmean denominator; single term)
__global__ void rescale3_kernel(P3 P) {
    const float ratio = P.ws.rescale[0];
    if (ratio == 1.f) return;
    const glue_decoder_args& a = P.a;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nth = (int64_t)gridDim.x * blockDim.x;
    float* gu = P.term[0].grad_u;
    if (gu)
        for (int64_t i = tid; i < (int64_t)a.B * a.K; i += nth) gu[i] *= ratio;
    if (a.grad_v)
        for (int64_t i = tid; i < (int64_t)a.F * a.K; i += nth) {
            const int64_t jj = i / a.K, k = i - jj * a.K;
            const int64_t vr = a.vidx ? a.vidx[jj] : jj;
            a.grad_v[vr * a.K + k] *= ratio;
        }
    float* tabs[4] = {a.grad_scale_lin, a.grad_bias, a.grad_disp, a.grad_zi_logits};
    for (int q = 0; q < 4; ++q)
        if (tabs[q])
            for (int64_t i = tid; i < (int64_t)a.n_batches * a.F; i += nth) tabs[q][i] *= ratio;
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
size_t ws_layout(const glue_decoder_args& a, int S, int ncta, int Fpad, char* base, Ws3* ws) {
    size_t off = 0;
    auto take = [&](size_t bytes, size_t align) {
        off = align_up(off, align);
        char* ptr = base ? base + off : nullptr;
        off += bytes;
        return ptr;
    };
    Ws3 w;
    w.trace = (long long*)take(TRACE_BYTES, 1024);  // (first: tools find it at the aligned base)
    w.vimg = (uint8_t*)take((size_t)2 * Fpad * ROW_BYTES, 1024);
    w.uimg = (uint8_t*)take((size_t)2 * S * 128 * ROW_BYTES, 1024);
    w.ptab = (float2*)take((size_t)a.n_batches * Fpad * sizeof(float2), 256);
    const size_t rows = (size_t)S * 128;
    w.pmax = (float*)take(2 * (size_t)ncta * rows * 4, 256);
    w.psum = (float*)take(2 * (size_t)ncta * rows * 4, 256);
    w.lse = (float*)take(rows * 4, 256);
    w.cpart = (float*)take((size_t)ncta * rows * 4, 256);
    w.c = (float*)take(rows * 4, 256);
    w.apart = (float*)take((size_t)ncta * rows * 64 * 4, 256);
    w.losspart = (float*)take((size_t)ncta * S * 4, 256);
    w.badpart = (int*)take((size_t)ncta * S * 4, 256);
    w.rescale = (float*)take(4, 256);
    w.counters = (unsigned int*)take(16, 256);
    if (ws) *ws = w;
    return align_up(off, 1024) + 1024;  // (+ slack to align the base itself)
}

// Persistent CTAs, feature tiles strided over them.  The fewest CTAs that still give every CTA at most
// ceil(ntiles / #SM) tiles: the makespan is that of a full grid (391 tiles: 3 per CTA with 148 CTAs and
// with 131), and the SMs left over take the kernels that run beside this call on other streams -- the
// decoder call of a small modality (16 CTAs at 2 000 genes) would otherwise wait for, or delay, a full wave.
int grid_for(int F, int& ntiles) {
    ntiles = (F + 127) / 128;
    const int sms = sm_count();
    if (ntiles <= sms) return ntiles;
    const int per_cta = (ntiles + sms - 1) / sms;
    return (ntiles + per_cta - 1) / per_cta;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

// [rows][64 bf16] row-major image, boxes of 128 rows, 128-byte swizzle
int make_map(CUtensorMap* map, void* base, uint64_t rows) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return GLUE_E_UNSUPPORTED;
    const cuuint64_t dims[2] = {64, rows};
    const cuuint64_t strides[1] = {ROW_BYTES};
    const cuuint32_t box[2] = {64, 128};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult rc = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return rc == CUDA_SUCCESS ? 0 : GLUE_E_UNSUPPORTED;
}

template <int FAM, int PHASE, bool GRAD, bool NB1, bool FAST>
int launch_main(const CUtensorMap& mv, const CUtensorMap& mu, const P3& P, int ncta, cudaStream_t st) {
    const size_t smem = sizeof(SmemMain3);
    auto k = main3_kernel<FAM, PHASE, GRAD, NB1, FAST>;
    GLUE_SMEM_ONCE(k, smem);
    k<<<ncta, NTHREADS, smem, st>>>(mv, mu, P);
    GLUE_CHECK_LAUNCH();
    return 0;
}

template <int FAM, bool NB1, bool FAST>
int launch_phase(int phase, bool grad, const CUtensorMap& mv, const CUtensorMap& mu, const P3& P, int ncta,
                 cudaStream_t st) {
    if (phase == 1) return launch_main<FAM, 1, true, NB1, false>(mv, mu, P, ncta, st);
    if (grad) return launch_main<FAM, 0, true, NB1, FAST>(mv, mu, P, ncta, st);
    return launch_main<FAM, 0, false, NB1, FAST>(mv, mu, P, ncta, st);
}

template <bool NB1>
int launch_family(int family, int phase, bool grad, bool fast, const CUtensorMap& mv, const CUtensorMap& mu, const P3& P,
                  int ncta, cudaStream_t st) {
    if (phase == 1) return launch_phase<GLUE_NB, NB1, false>(1, true, mv, mu, P, ncta, st);  // family independent
    if constexpr (NB1) {
        if (family == GLUE_NB && fast) return launch_phase<GLUE_NB, true, true>(0, grad, mv, mu, P, ncta, st);
    }
    switch (family) {
        case GLUE_NB: return launch_phase<GLUE_NB, NB1, false>(0, grad, mv, mu, P, ncta, st);
        case GLUE_ZINB: return launch_phase<GLUE_ZINB, NB1, false>(0, grad, mv, mu, P, ncta, st);
        case GLUE_NORMAL: return launch_phase<GLUE_NORMAL, NB1, false>(0, grad, mv, mu, P, ncta, st);
        default: return launch_phase<GLUE_ZILN, NB1, false>(0, grad, mv, mu, P, ncta, st);
    }
}

}  // namespace v3

bool tc3_eligible(const glue_decoder_multi_args& m, int mode) {
    const glue_decoder_args& a = m.shared;
    if (mode != MODE_LOSS && mode != MODE_GRAD) return false;
    if (m.n_terms < 1 || m.n_terms > v3::MAXSEG) return false;
    if (a.wmat) return false;
    if (a.B > 128 || a.K > 64 || (a.K & 1) || a.n_batches > v3::NB_MAX) return false;
    if (a.n_batches > 1 && a.b && !a.row_order) return false;
    if ((reinterpret_cast<uintptr_t>(a.grad_v) & 3u) != 0) return false;
    if (m.n_terms > 1) {  // NaN skipping (nanmean) rescales by one factor: single term only
        for (int s = 0; s < m.n_terms; ++s)
            if (a.reduce == 0 && !m.row_weight[s]) return false;
    }
    return v3::encode_fn() != nullptr;
}

size_t tc3_workspace_bytes(const glue_decoder_args& a, int n_terms) {
    int ntiles;
    const int ncta = v3::grid_for(a.F, ntiles);
    return v3::ws_layout(a, n_terms, ncta, ntiles * 128, nullptr, nullptr);
}

int tc3_run(const glue_decoder_multi_args& m, int mode, cudaStream_t st) {
    using namespace v3;
    const glue_decoder_args& a = m.shared;
    P3 P = {};
    P.a = a;
    P.S = m.n_terms;
    const int ncta = grid_for(a.F, P.ntiles);
    P.Fpad = P.ntiles * 128;
    P.ncell = a.B;
    const size_t need = ws_layout(a, P.S, ncta, P.Fpad, nullptr, nullptr);
    if (a.workspace_bytes < need) return GLUE_E_WORKSPACE;
    char* base = (char*)(((uintptr_t)a.workspace + 1023) & ~(uintptr_t)1023);
    ws_layout(a, P.S, ncta, P.Fpad, base, &P.ws);
    const bool nbf = is_nb_family(a.family);
    const bool grad = mode == MODE_GRAD;
    const bool nb1 = a.n_batches == 1;
    for (int s = 0; s < P.S; ++s) {
        Term& T = P.term[s];
        T.u = m.u[s], T.row_weight = m.row_weight[s], T.loss = m.loss[s], T.grad_u = grad ? m.grad_u[s] : nullptr;
        T.gscale = m.grad_scale[s] != 0.f ? m.grad_scale[s] : 1.f;
        T.w0 = (float)(-1.0 / ((double)a.B * (double)a.F)) * T.gscale;
        if (!T.u) return GLUE_E_BADARG;
    }
    const bool small = (long long)a.B * (long long)a.F < (1ll << 31);
    const bool fast = a.family == GLUE_NB && nb1 && a.K <= 62 && small && !(a.impl & 0x100);
    P.cinv = (nbf && nb1 && a.K <= 62 && grad) ? 1 : 0;
    if (!nbf) P.ws.ptab = nullptr;
    if (a.impl & 0x200)
        GLUE_CUDA(cudaMemsetAsync(P.ws.trace, 0, TRACE_BYTES, st));
    else
        P.ws.trace = nullptr;

    CUtensorMap mapV, mapU;
    int rc = make_map(&mapV, P.ws.vimg, (uint64_t)2 * P.Fpad);
    if (rc) return rc;
    rc = make_map(&mapU, P.ws.uimg, (uint64_t)2 * P.S * 128);
    if (rc) return rc;

    if (grad && a.n_batches > 1) {
        const size_t tab = (size_t)a.n_batches * a.F * sizeof(float);
        if (a.grad_scale_lin) GLUE_CUDA(cudaMemsetAsync(a.grad_scale_lin, 0, tab, st));
        if (a.grad_bias) GLUE_CUDA(cudaMemsetAsync(a.grad_bias, 0, tab, st));
        if (a.grad_disp) GLUE_CUDA(cudaMemsetAsync(a.grad_disp, 0, tab, st));
        if (a.grad_zi_logits) GLUE_CUDA(cudaMemsetAsync(a.grad_zi_logits, 0, tab, st));
    }
    {
        const int64_t nthreads = ((int64_t)P.Fpad + (int64_t)P.S * 128) * 8;
        presplit_kernel<<<(unsigned)((nthreads + 255) / 256), 256, 0, st>>>(P);
        GLUE_CHECK_LAUNCH();
    }
    if (nbf) {
        const size_t smem = sizeof(SmemStats3);
        GLUE_SMEM_ONCE(rowstats3_kernel, smem);
        rowstats3_kernel<<<ncta, NTHREADS, smem, st>>>(mapV, mapU, P);
        GLUE_CHECK_LAUNCH();
        lse3_kernel<<<(P.S * 128 + 31) / 32, 1024, 0, st>>>(P.ws, 2 * ncta, P.S * 128);  // 2 ncta <= 320 partials
        GLUE_CHECK_LAUNCH();
    }
    rc = nb1 ? launch_family<true>(a.family, 0, grad, fast, mapV, mapU, P, ncta, st)
             : launch_family<false>(a.family, 0, grad, fast, mapV, mapU, P, ncta, st);
    if (rc) return rc;
    if (grad) {
        if (nbf) {
            rc = nb1 ? launch_family<true>(a.family, 1, true, fast, mapV, mapU, P, ncta, st)
                     : launch_family<false>(a.family, 1, true, fast, mapV, mapU, P, ncta, st);
            if (rc) return rc;
        }
        grad_u3_kernel<<<P.S * P.ncell, 1024, 0, st>>>(P, ncta);  // ncta <= 160
        GLUE_CHECK_LAUNCH();
        if (P.S == 1 && a.reduce == 0 && !m.row_weight[0]) {
            rescale3_kernel<<<sm_count(), 256, 0, st>>>(P);
            GLUE_CHECK_LAUNCH();
        }
    }
    return 0;
}

}  // namespace glue
