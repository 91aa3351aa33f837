// Fused data decoders, "v2": tcgen05 tensor-core contraction + TMEM-resident logits (sm_100a).
//
// Same mathematics and workspace protocol as decoder.cu (see the header there); what changes is
// where the K = latent_dim contractions run.  ncu on v1 showed the low-K contraction to be
// FFMA/latency bound (profiles/r01_decoder_v1_full.md), the condition under which the design
// moves it to the tensor cores:
//
//   * u and the v tile are split into bf16 (hi, lo) pairs and staged in shared memory in the
//     UMMA SWIZZLE_128B canonical layout ([rows][64 bf16]); every fp32 product is rebuilt from
//     three bf16 MMAs (hi*hi + hi*lo + lo*hi, fp32 accumulation, ~2^-17 relative error);
//   * logits tile  D[feature, cell] = V U^T  (M = 128 features = TMEM lanes, N = 128 cells)
//     lives in TMEM only; epilogue thread <-> feature keeps the per-feature parameters and the
//     per-feature gradient sums in registers exactly as in v1;
//   * the epilogue writes the coefficient tile  X[feature, cell] = dL/da * softplus(scale)
//     once (bf16 hi/lo) to shared memory; the SAME bytes are consumed by two MMAs, as a K-major
//     A operand (d/dv tile = X U, accumulator D1) and as an MN-major A operand (d/du partial =
//     X^T V, accumulator D2 kept in TMEM across all tiles of the CTA);
//   * the soft-max row statistics pass uses the transposed tile D[cell, feature] = U V^T so that
//     max / sum-exp are per-thread reductions over TMEM columns.
//
// Warp roles (416 threads, 1 CTA / SM): warps 0-7 epilogue (warp w owns TMEM lanes
// 32 (w & 3) .. +31 and one half of the cell columns), warps 8-11 stage the v tiles (double
// buffered) and prefetch the step's observations into L2, warp 12 issues
// tcgen05.mma (one lane).  Pipelines are mbarrier based.
#include <stddef.h>
#include <stdlib.h>

#include "decoder_common.cuh"
#include "tc_common.cuh"
#include "tc_math.cuh"

namespace glue {
using namespace tc;
using namespace tcm;

namespace {

constexpr int NTHREADS = 416;
constexpr int EPI_THREADS = 256;
constexpr int EPI_WARPS = 8;
constexpr int PROD_T0 = 256;  // warps 8-11: one producer thread per operand row
constexpr int PROD_THREADS = 128;
constexpr int MMA_WARP = 12;
constexpr int NB_MAX = 24;  // parameter rows the row-statistics table holds per v stage
constexpr int TC_MAX_CHUNKS = 32;  // cell blocks of 128 per call (B <= 4096)

constexpr uint32_t COL_ACC = 0;    // logits accumulators: 2 x 128 columns
constexpr uint32_t COL_D1 = 256;   // d/dv tile accumulators: 2 x 64 columns
constexpr uint32_t COL_D2 = 384;   // d/du accumulator: 64 columns
constexpr uint32_t TMEM_COLS = 512;

// UMMA descriptors (see tc_common.cuh)
constexpr uint64_t DESC_K = smem_desc_base(16, 1024);         // K-major, 8-row groups 1024 B apart
constexpr uint64_t DESC_MN1 = smem_desc_base(1024, 1024);     // MN-major, one 64-element atom along MN
constexpr uint64_t DESC_MN_X = smem_desc_base(16384, 1024);   // MN-major X: two atoms along MN, 16 KB apart
constexpr uint32_t IDESC_FWD = instr_desc_bf16(128, 128, 0, 0);
constexpr uint32_t IDESC_D1 = instr_desc_bf16(128, 64, 0, 1);
constexpr uint32_t IDESC_D2 = instr_desc_bf16(128, 64, 1, 1);

struct Bars {
    uint64_t v_full[2], v_empty[2], acc_full[2], acc_empty[2], x_full, x_empty, d1_full[2], d1_empty[2], d2_full;
};

struct __align__(1024) SmemMain {
    uint8_t u_hi[OP_BYTES], u_lo[OP_BYTES];
    uint8_t v_hi[2][OP_BYTES], v_lo[2][OP_BYTES];
    uint8_t x_hi[2 * OP_BYTES], x_lo[2 * OP_BYTES];  // [cell block of 64][128 feature rows][128 B]
    long long xoff[128];
    float lse[128], l[128], c[128], wrow[128];
    int bidx[128];
    float cw[2][4][128];
    float c_acc[128];
    float stage[128][65];  // d/dv tile on its way from TMEM (thread <-> row) to coalesced stores
    float red_l[EPI_WARPS];
    int red_b[EPI_WARPS];
    float cpart3[3][128];
    long long gvrow[3][128];           // d/dv element offset of every feature row of the tiles in flight (-1 beyond F)
    alignas(16) uint32_t xoff32[136];  // element offset of every cell's row of x (streamlined path); 8 pad entries
    alignas(16) float zeros[128];      // weights seen by feature rows beyond F
    Bars bars;
    uint32_t tmem_base;
    uint32_t ticket;
};

struct __align__(1024) SmemStats {
    uint8_t u_hi[OP_BYTES], u_lo[OP_BYTES];
    uint8_t v_hi[2][OP_BYTES], v_lo[2][OP_BYTES];
    float2 ptab[2][NB_MAX][130];  // (softplus(scale_lin), bias) per (stage, batch, feature-in-tile)
    int bidx[128];
    Bars bars;
    uint32_t tmem_base;
    uint32_t ticket;
};

static_assert(offsetof(SmemMain, wrow) % 16 == 0 && offsetof(SmemMain, lse) % 16 == 0 &&
                  offsetof(SmemMain, xoff32) % 16 == 0 && offsetof(SmemMain, zeros) % 16 == 0,
              "vector loads of the per-cell tables");
static_assert(sizeof(SmemMain) <= 227 * 1024, "shared memory per CTA");

struct TcParams {
    DecParams p;
    int ntiles;
    int cell0, ncell;  // cell block [cell0, cell0 + ncell) of the (sorted) minibatch, ncell <= 128
    int acc;           // outputs accumulate (a previous cell block already wrote them)
    int chunk, nchunks;  // this launch handles cell block `chunk` of `nchunks`
    long long* trace;  // optional: clock64 stamps of CTA 0, [kernel 0..2][tile < 8][event < 16]
    int pdl;           // launches carry the programmatic-dependent-launch attribute (GLUE_TC_PDL=0 switches it off)
};

#define TC_TRACE(kern, it, ev)                                                              \
    do {                                                                                    \
        if (P.trace != nullptr && blockIdx.x == 0 && (it) < 8)                              \
            P.trace[((kern) * 8 + (it)) * 16 + (ev)] = clock64();                           \
    } while (0)

__device__ __forceinline__ int orig_row(const glue_decoder_args& a, int pos) {
    return a.row_order ? a.row_order[pos] : pos;
}

// v tile producer: one thread per feature row.  CINV (streamlined pass B, single parameter row,
// K <= 62): operand column K carries 1 / softplus(scale_lin_j).  The logits and the d/dv product
// never see it (u has zeros there); in the d/du product  D2[cell, k] = sum_j X[j, cell] V[j, k]
// with X = G softplus(scale) it makes column K the soft-max correction  c_i = sum_j G_ij.
template <bool CINV = false, int NPROD = PROD_THREADS>
__device__ __forceinline__ void produce_v_rows(const glue_decoder_args& a, uint8_t* hi, uint8_t* lo, int f0, int pt,
                                               long long* gvrow = nullptr) {
    for (int r = pt; r < 128; r += NPROD) {
        const int j = f0 + r;
        const float* src = nullptr;
        float rs = 0.f;
        if (gvrow) gvrow[r] = -1;
        if (j < a.F) {
            const int64_t vr = a.vidx ? a.vidx[j] : (int64_t)j;
            src = a.v + vr * a.K;
            if (gvrow && a.grad_v) gvrow[r] = vr * a.K;
            if (CINV) {
                const float sp = softplusf(a.scale_lin[j]);
                rs = sp > 1e-18f ? 1.f / sp : 0.f;
            }
        }
        stage_row(smem_u32(hi), smem_u32(lo), (uint32_t)r, src, a.K);
        if (CINV) {  // (same thread: ordered after the zero fill of the chunk above)
            uint32_t h2, l2;
            split2(rs, 0.f, h2, l2);
            const uint32_t off = swz((uint32_t)r, (uint32_t)a.K >> 3) + (((uint32_t)a.K & 7u) << 1);
            asm volatile("st.shared.b32 [%0], %1;" ::"r"(smem_u32(hi) + off), "r"(h2) : "memory");
            asm volatile("st.shared.b32 [%0], %1;" ::"r"(smem_u32(lo) + off), "r"(l2) : "memory");
        }
    }
}

// logits tile: acc = A B^T with both operands K-major [128][64], three bf16 products
__device__ __forceinline__ void issue_logits(uint32_t tmem_acc, const uint8_t* a_hi, const uint8_t* a_lo,
                                             const uint8_t* b_hi, const uint8_t* b_lo) {
    const uint32_t ah = smem_u32(a_hi), al = smem_u32(a_lo), bh = smem_u32(b_hi), bl = smem_u32(b_lo);
    const uint32_t as[3] = {al, ah, ah}, bs[3] = {bh, bl, bh};  // small terms first
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
        for (uint32_t ks = 0; ks < 4; ++ks)
            umma_bf16(tmem_acc, smem_desc(DESC_K, as[c] + ks * 32), smem_desc(DESC_K, bs[c] + ks * 32), IDESC_FWD,
                      (c | ks) ? 1u : 0u);
}

// ---------------------------------------------------------------------------------------------
// pass A: soft-max row statistics.  D[cell, feature] = U V^T; thread <-> cell.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NTHREADS, 1) tc_rowstats_kernel(TcParams P) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    SmemStats& S = *reinterpret_cast<SmemStats*>(smem_raw);
    if ((smem_u32(smem_raw) & 1023u) != 0) __trap();  // swizzled operand tiles need 1024 B alignment
    const glue_decoder_args& a = P.p.a;
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const int cta = blockIdx.x, ncta = gridDim.x;
    const int ntl = (P.ntiles - cta + ncta - 1) / ncta;
    if (t == 0) TC_TRACE(0, 7, 2);
    griddep_launch();
    griddep_wait();

    if (t == 0) {
        for (int s = 0; s < 2; ++s) {
            mbar_init(&S.bars.v_full[s], PROD_THREADS);
            mbar_init(&S.bars.v_empty[s], EPI_WARPS);
            mbar_init(&S.bars.acc_full[s], 1);
            mbar_init(&S.bars.acc_empty[s], EPI_WARPS);
        }
        mbar_fence_init();
    }
    if (warp == MMA_WARP) tmem_alloc(&S.tmem_base, 256);
    if (t < 128) {
        const bool ok = t < P.ncell;
        const int orig = ok ? orig_row(a, P.cell0 + t) : 0;
        stage_row(smem_u32(S.u_hi), smem_u32(S.u_lo), (uint32_t)t, ok ? a.u + (int64_t)orig * a.K : nullptr, a.K);
        S.bidx[t] = (ok && a.b) ? (int)a.b[orig] : 0;
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = S.tmem_base;
    if (t == 0) TC_TRACE(0, 7, 0);

    if (warp >= PROD_T0 / 32 && warp < MMA_WARP) {
        const int pt = t - PROD_T0;
        for (int it = 0; it < ntl; ++it) {
            const int s = it & 1, f0 = (cta + it * ncta) * 128;
            mbar_wait(&S.bars.v_empty[s], ((it >> 1) & 1) ^ 1);
            if (pt == 0) TC_TRACE(0, it, 0);
            produce_v_rows(a, S.v_hi[s], S.v_lo[s], f0, pt);
            if (pt == 0) TC_TRACE(0, it, 12);
            for (int r = pt; r < 128; r += PROD_THREADS) {
                const int j = f0 + r;
                for (int b = 0; b < a.n_batches; ++b) {
                    float2 pr = make_float2(0.f, 0.f);
                    if (j < a.F) {
                        const int64_t off = (int64_t)b * a.F + j;
                        pr = make_float2(softplusf(a.scale_lin[off]), a.bias[off]);
                    }
                    S.ptab[s][b][r] = pr;
                }
            }
            fence_proxy_async();
            mbar_arrive(&S.bars.v_full[s]);
            if (pt == 0) TC_TRACE(0, it, 1);
        }
    } else if (warp == MMA_WARP) {
        if (lane == 0) {
            for (int it = 0; it < ntl; ++it) {
                const int s = it & 1;
                mbar_wait(&S.bars.v_full[s], (it >> 1) & 1);
                mbar_wait(&S.bars.acc_empty[s], ((it >> 1) & 1) ^ 1);
                tc_fence_after();
                issue_logits(tmem_base + COL_ACC + s * 128, S.u_hi, S.u_lo, S.v_hi[s], S.v_lo[s]);
                umma_commit(&S.bars.acc_full[s]);
            }
        }
    } else if (warp < EPI_WARPS) {
        const int q = warp & 3, h = warp >> 2;
        const int cell = 32 * q + lane;
        const int my_b = S.bidx[cell];
        float m = -INFINITY, ssum = 0.f;
        for (int it = 0; it < ntl; ++it) {
            const int s = it & 1, f0 = (cta + it * ncta) * 128;
            const int nvalid = min(128, a.F - f0);
            mbar_wait(&S.bars.acc_full[s], (it >> 1) & 1);
            tc_fence_after();
            if (t == 0) TC_TRACE(0, it, 6);
#pragma unroll 1
            for (int cc = 0; cc < 2; ++cc) {
                const int c0 = 64 * h + 32 * cc;
                if (c0 < nvalid) {
                    uint32_t d[32];
                    tmem_ld32(tmem_base + ((uint32_t)(32 * q) << 16) + COL_ACC + s * 128 + c0, d);
                    tmem_ld_wait();
                    const float2* tab = &S.ptab[s][my_b][c0];
                    float av[32];
                    float cm = -INFINITY;
#pragma unroll
                    for (int r = 0; r < 32; ++r) {
                        const float2 pr = tab[r];
                        const float val = fmaf(pr.x, __uint_as_float(d[r]), pr.y);
                        av[r] = (c0 + r < nvalid) ? val : -INFINITY;
                        cm = fmaxf(cm, av[r]);
                    }
                    const float nm = fmaxf(m, cm);
                    if (nm > -INFINITY) {
                        float part = 0.f;
#pragma unroll
                        for (int r = 0; r < 32; ++r) part += expf(av[r] - nm);
                        ssum = ssum * expf(m - nm) + part;
                        m = nm;
                    }
                }
            }
            if (t == 0) TC_TRACE(0, it, 8);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&S.bars.acc_empty[s]);
                mbar_arrive(&S.bars.v_empty[s]);
            }
        }
        if (cell < P.ncell) {
            const int64_t slot = (int64_t)(cta * 2 + h) * a.B + P.cell0 + cell;
            P.p.ws.pmax[slot] = m;
            P.p.ws.psum[slot] = ssum;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == MMA_WARP) tmem_dealloc(tmem_base, 256);
    if (t == 0) TC_TRACE(0, 7, 1);
}

// ---------------------------------------------------------------------------------------------
// pass B (PHASE 0: likelihood + all G terms) and pass C (PHASE 1: the -p_ij c_i soft-max terms).
// D[feature, cell] = V U^T; thread <-> feature.  NB1: a single parameter row (n_batches == 1).
// FAST: streamlined NB pass B (NB1, K <= 62, B F < 2^31) -- 32-bit unpredicated observation
// loads, gamma terms of the counts 0 / 1 / 2 as a quadratic, everything else (larger counts, NaN)
// handled per thread ahead of the packed arithmetic, c_i taken from the tensor cores
// (produce_v_rows<CINV>).
// (A variant with 16 epilogue warps -- four per TMEM lane quarter -- was parity-green and no
// faster: with the streamlined arithmetic the loop is bound by the FP32 pipe, not by latency.)
// ---------------------------------------------------------------------------------------------
template <int FAM, int PHASE, bool GRAD, bool NB1, bool FAST>
__global__ void __launch_bounds__(NTHREADS, 1) tc_main_kernel(TcParams P) {
    static_assert(!FAST || (FAM == GLUE_NB && NB1 && PHASE == 0), "streamlined path: NB pass B");
    constexpr int EW = EPI_WARPS, NPROD = PROD_THREADS, NT = NTHREADS, EPI_T = EPI_THREADS;
    constexpr int PROD0 = PROD_T0, MMAW = MMA_WARP;
    constexpr int NH = 2;    // threads sharing a feature row (splits of the cell columns)
    constexpr int DC = 32;   // columns of the 64-wide d/du accumulator per epilogue thread
    constexpr bool CINV = FAST && GRAD;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    SmemMain& S = *reinterpret_cast<SmemMain*>(smem_raw);
    if ((smem_u32(smem_raw) & 1023u) != 0) __trap();  // swizzled operand tiles need 1024 B alignment
    const glue_decoder_args& a = P.p.a;
    constexpr bool NBF = (FAM == GLUE_NB || FAM == GLUE_ZINB);
    constexpr bool ZI = (FAM == GLUE_ZINB || FAM == GLUE_ZILN);
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const int cta = blockIdx.x, ncta = gridDim.x;
    const int ntl = (P.ntiles - cta + ncta - 1) / ncta;
    const int ncell = P.ncell;
    if (t == 0) TC_TRACE(1 + PHASE, 7, 2);
    griddep_launch();
    griddep_wait();

    if (t == 0) {
        for (int s = 0; s < 2; ++s) {
            mbar_init(&S.bars.v_full[s], NPROD);
            mbar_init(&S.bars.v_empty[s], 1);
            mbar_init(&S.bars.acc_full[s], 1);
            mbar_init(&S.bars.acc_empty[s], EW);
            mbar_init(&S.bars.d1_full[s], 1);
            mbar_init(&S.bars.d1_empty[s], NPROD / 32);
        }
        mbar_init(&S.bars.x_full, EW);
        mbar_init(&S.bars.x_empty, 1);
        mbar_init(&S.bars.d2_full, 1);
        mbar_fence_init();
    }
    if (warp == MMAW) tmem_alloc(&S.tmem_base, TMEM_COLS);
    if (t < 128) {
        const bool ok = t < ncell;
        const int pos = P.cell0 + t;
        const int orig = ok ? orig_row(a, pos) : 0;
        stage_row(smem_u32(S.u_hi), smem_u32(S.u_lo), (uint32_t)t, ok ? a.u + (int64_t)orig * a.K : nullptr, a.K);
        S.bidx[t] = (ok && a.b) ? (int)a.b[orig] : 0;
        S.xoff[t] = (long long)orig * a.F;
        S.xoff32[t] = (uint32_t)((long long)orig * a.F);  // (FAST: B F < 2^31 checked by the host)
        S.zeros[t] = 0.f;
        if (t < 8) S.xoff32[128 + t] = 0u;
        S.l[t] = (ok && a.l) ? a.l[orig] : 1.f;
        S.lse[t] = (ok && NBF) ? -(P.p.ws.lse[pos] * LOG2EF) : 0.f;  // -lse in base 2
        S.c[t] = (ok && PHASE == 1) ? P.p.ws.c[pos] : 0.f;
        S.c_acc[t] = 0.f;
        S.wrow[t] = ok ? (a.row_weight ? -a.row_weight[orig] * P.p.gscale : P.p.w0) : 0.f;  // weight of log_prob
    }
    if (GRAD) {  // cells beyond ncell must read as exact zeros in the coefficient tile
        uint4* xz = reinterpret_cast<uint4*>(S.x_hi);
        for (int i = t; i < (int)(4 * OP_BYTES / 16); i += NT) xz[i] = make_uint4(0, 0, 0, 0);
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = S.tmem_base;
    if (t == 0) TC_TRACE(1 + PHASE, 7, 0);

    if (warp >= PROD0 / 32 && warp < MMAW) {
        // ===== v tile producers (one thread per row); x tile of the same step -> L2; these four
        // warps (one per TMEM lane quarter) also move the finished d/dv tiles out, two tiles behind
        // the one they stage, so the write-out is off the epilogue's critical path =====
        const int pt = t - PROD0;        // operand row = TMEM lane = feature row of the tile
        const int pw = warp - PROD0 / 32;
        const bool add_v = (PHASE == 1) || P.acc;
        auto readout_d1 = [&](int itp) {
            const int s = itp & 1;
            mbar_wait(&S.bars.d1_full[s], (itp >> 1) & 1);
            tc_fence_after();
            uint32_t r0[32], r1[32];
            const uint32_t taddr = tmem_base + ((uint32_t)(32 * pw) << 16) + COL_D1 + s * 64;
            tmem_ld32(taddr, r0);
            tmem_ld32(taddr + 32, r1);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&S.bars.d1_empty[s]);
            // transpose through shared memory: thread <-> feature row in TMEM, but a row of d/dv
            // (K contiguous floats) should leave with consecutive lanes on consecutive addresses
#pragma unroll
            for (int k = 0; k < 32; ++k) {
                S.stage[pt][k] = __uint_as_float(r0[k]);
                S.stage[pt][32 + k] = __uint_as_float(r1[k]);
            }
            named_bar_sync(3, NPROD);
            {
                // warp pw writes rows pw, pw + 4, ...: lane <-> float2 of the row; row pointers were
                // resolved (vidx) when the tile was staged
                const long long* rows = S.gvrow[itp % 3];
                const int k2 = lane;
                const bool kok = 2 * k2 < a.K;
                float* const gbase = a.grad_v + 2 * k2;
#pragma unroll 1
                for (int batch = 0; batch < 2; ++batch) {
                    float2 old[16];
                    long long off[16];
#pragma unroll
                    for (int rr = 0; rr < 16; ++rr) {  // every load before the first store
                        off[rr] = kok ? rows[pw + 4 * (rr + 16 * batch)] : -1;
                        old[rr] = (add_v && off[rr] >= 0) ? *reinterpret_cast<const float2*>(gbase + off[rr])
                                                          : make_float2(0.f, 0.f);
                    }
#pragma unroll
                    for (int rr = 0; rr < 16; ++rr) {
                        const int row = pw + 4 * (rr + 16 * batch);
                        const float v0 = S.stage[row][2 * k2], v1 = S.stage[row][2 * k2 + 1];
                        if (off[rr] >= 0)
                            *reinterpret_cast<float2*>(gbase + off[rr]) = make_float2(v0 + old[rr].x, v1 + old[rr].y);
                    }
                }
            }
            named_bar_sync(3, NPROD);  // the staging tile is free again
            if (pt == 0) TC_TRACE(1 + PHASE, itp, 11);
        };
        for (int it = 0; it < ntl; ++it) {
            const int s = it & 1, f0 = (cta + it * ncta) * 128;
            if (PHASE == 0) {
                const int nvalid = min(128, a.F - f0);
                for (int cell = pt; cell < ncell; cell += NPROD) {
                    const float* px = a.x + S.xoff[cell] + f0;
                    for (int o = 0; o < nvalid; o += 32) asm volatile("prefetch.global.L2 [%0];" ::"l"(px + o));
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(px + nvalid - 1));
                }
            }
            mbar_wait(&S.bars.v_empty[s], ((it >> 1) & 1) ^ 1);
            if (pt == 0) TC_TRACE(1 + PHASE, it, 0);
            produce_v_rows<CINV, NPROD>(a, S.v_hi[s], S.v_lo[s], f0, pt, GRAD ? S.gvrow[it % 3] : nullptr);
            fence_proxy_async();
            mbar_arrive(&S.bars.v_full[s]);
            if (pt == 0) TC_TRACE(1 + PHASE, it, 1);
            if (GRAD && it >= 2) readout_d1(it - 2);  // (its MMAs completed before v_empty[s] was released)
        }
        if (GRAD) {
            if (ntl >= 2) readout_d1(ntl - 2);
            readout_d1(ntl - 1);
            if (pt == 0) TC_TRACE(1 + PHASE, 6, 13);
        }
    } else if (warp == MMAW) {
        // ===== MMA issuer =====
        if (lane == 0) {
            auto issue_fwd = [&](int it) {
                const int s = it & 1;
                mbar_wait(&S.bars.v_full[s], (it >> 1) & 1);
                mbar_wait(&S.bars.acc_empty[s], ((it >> 1) & 1) ^ 1);
                tc_fence_after();
                issue_logits(tmem_base + COL_ACC + s * 128, S.v_hi[s], S.v_lo[s], S.u_hi, S.u_lo);
                umma_commit(&S.bars.acc_full[s]);
                if (!GRAD) umma_commit(&S.bars.v_empty[s]);
                TC_TRACE(1 + PHASE, it, 2);
            };
            issue_fwd(0);
            const uint32_t nks1 = (uint32_t)(ncell + 15) >> 4;  // K steps over cells
            for (int it = 0; it < ntl; ++it) {
                if (it + 1 < ntl) issue_fwd(it + 1);
                if (GRAD) {
                    const int s = it & 1;
                    mbar_wait(&S.bars.x_full, it & 1);
                    TC_TRACE(1 + PHASE, it, 3);
                    mbar_wait(&S.bars.d1_empty[s], ((it >> 1) & 1) ^ 1);
                    tc_fence_after();
                    const uint32_t xh = smem_u32(S.x_hi), xl = smem_u32(S.x_lo);
                    const uint32_t uh = smem_u32(S.u_hi), ul = smem_u32(S.u_lo);
                    const uint32_t vh = smem_u32(S.v_hi[s]), vl = smem_u32(S.v_lo[s]);
                    const uint32_t xs[3] = {xl, xh, xh};
                    // D1[feature, k] = sum_cell X[feature, cell] U[cell, k]
                    {
                        const uint32_t bs[3] = {uh, ul, uh};
                        const uint32_t d1 = tmem_base + COL_D1 + s * 64;
#pragma unroll
                        for (int c = 0; c < 3; ++c)
                            for (uint32_t ks = 0; ks < nks1; ++ks)
                                umma_bf16(d1, smem_desc(DESC_K, xs[c] + (ks >> 2) * OP_BYTES + (ks & 3) * 32),
                                          smem_desc(DESC_MN1, bs[c] + ks * 2048), IDESC_D1, (c | ks) ? 1u : 0u);
                        umma_commit(&S.bars.d1_full[s]);
                    }
                    // D2[cell, k] += sum_feature X[feature, cell] V[feature, k]
                    {
                        const uint32_t bs[3] = {vh, vl, vh};
                        const uint32_t d2 = tmem_base + COL_D2;
#pragma unroll
                        for (int c = 0; c < 3; ++c)
#pragma unroll
                            for (uint32_t ks = 0; ks < 8; ++ks)
                                umma_bf16(d2, smem_desc(DESC_MN_X, xs[c] + ks * 2048),
                                          smem_desc(DESC_MN1, bs[c] + ks * 2048), IDESC_D2, (it | c | ks) ? 1u : 0u);
                        umma_commit(&S.bars.x_empty);
                        umma_commit(&S.bars.v_empty[s]);
                        TC_TRACE(1 + PHASE, it, 4);
                    }
                }
            }
            if (GRAD) umma_commit(&S.bars.d2_full);
        }
    } else if (warp < EW) {
        // ===== epilogue =====
        const int q = warp & 3, h = warp >> 2;
        const int frow = 32 * q + lane;
        const uint32_t lane_addr = (uint32_t)(32 * q) << 16;
        // groups of 8 cells (= one 16-byte chunk of a coefficient-tile row); the NH threads that
        // share a feature split the groups in contiguous runs
        const int ngrp = (ncell + 7) >> 3;
        const int g_begin = (ngrp * h + NH - 1) / NH;
        const int g_end = (ngrp * (h + 1) + NH - 1) / NH;
        const bool skip_nan = a.reduce == 0 && !a.row_weight;  // nanmean semantics
        float loss_acc = 0.f;
        int bad_acc = 0;

        for (int it = 0; it < ntl; ++it) {
            const int s = it & 1;
            const int f0 = (cta + it * ncta) * 128;
            const int j = f0 + frow;
            const bool jok = j < a.F;
            int cur_b = -1;
            FeatParams fp = {};
            float gb = 0.f, gs = 0.f, gd = 0.f, gz = 0.f;
            // first parameter run of the upper-half thread is flushed after the tile barrier so that
            // the two threads sharing a feature update a table entry in a fixed order
            bool pend = false;
            int64_t pend_off = 0;
            float pend_b = 0.f, pend_s = 0.f, pend_d = 0.f, pend_z = 0.f;
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
            auto flush = [&]() {
                if (GRAD && cur_b >= 0 && jok) {
                    const int64_t off = (int64_t)cur_b * a.F + j;
                    const float vs = gs * fp.sg;
                    if (h == 1 && !pend) {
                        pend = true, pend_off = off, pend_b = gb, pend_s = vs, pend_d = gd, pend_z = gz;
                    } else {
                        write_param(off, gb, vs, gd, gz, (PHASE == 1) || P.acc || !NB1 || h == 1);
                    }
                }
                gb = gs = gd = gz = 0.f;
            };
            float s2 = 0.f, b2 = 0.f, lg1 = 0.f, dg1 = 0.f, lg2 = 0.f, dg2 = 0.f;
            f2 ngb2 = pk(0.f, 0.f), ngs2 = pk(0.f, 0.f), gd2 = pk(0.f, 0.f), loss2 = pk(0.f, 0.f);
            if (NB1) {
                fp = load_params<FAM>(a, 0, j, jok);
                cur_b = 0;
                s2 = fp.s * LOG2EF, b2 = fp.bias * LOG2EF;  // logits in base 2
                if (PHASE == 1 && !jok) b2 = -INFINITY;     // features beyond F: p = 0
                if (NBF && PHASE == 0) {  // log-gamma / digamma differences for counts 1 and 2
                    const float ti = 1.f / fp.th;
                    lg1 = fp.disp, dg1 = ti;
                    lg2 = logf(fp.th * (fp.th + 1.f)) - 0.69314718055994530942f;
                    dg2 = ti + 1.f / (fp.th + 1.f);
                }
            }

            // observations of the first group are requested before the logits are waited for
            float xn[8];
            // FAST: every load is in bounds without predicates -- feature column clamped to F - 1,
            // cells beyond ncell (and the group after the last one) read row offset 0; what they
            // compute carries weight 0
            const float* xj = a.x + (jok ? j : a.F - 1);
            if (FAST) asm volatile("" : "+l"(xj));  // keep the column base in a register pair (IMAD.WIDE per load)
            auto load_x = [&](int g) {
                if constexpr (FAST) {
                    const uint4 o0 = *reinterpret_cast<const uint4*>(&S.xoff32[8 * g]);
                    const uint4 o1 = *reinterpret_cast<const uint4*>(&S.xoff32[8 * g + 4]);
                    xn[0] = __ldg(xj + o0.x), xn[1] = __ldg(xj + o0.y), xn[2] = __ldg(xj + o0.z), xn[3] = __ldg(xj + o0.w);
                    xn[4] = __ldg(xj + o1.x), xn[5] = __ldg(xj + o1.y), xn[6] = __ldg(xj + o1.z), xn[7] = __ldg(xj + o1.w);
                } else {
#pragma unroll
                    for (int e8 = 0; e8 < 8; ++e8) {
                        const int cell = 8 * g + e8;
                        xn[e8] = (PHASE == 0 && jok && g < g_end && cell < ncell)
                                     ? __ldg(a.x + S.xoff[cell] + j) : 0.f;
                    }
                }
            };
            const f2 S2 = pk(s2, s2), B2 = pk(b2, b2), NS2 = pk(-fp.s, -fp.s), TH2 = pk(fp.th, fp.th);
            const f2 EPS2 = pk(GLUE_EPS, GLUE_EPS), M1 = pk(-1.f, -1.f), LN2_2 = pk(LN2F, LN2F);
            const f2 NLN2 = pk(-LN2F, -LN2F), DISP2 = pk(fp.disp, fp.disp);
            // FAST: log-gamma / digamma differences of the counts 0, 1, 2 as x (c1 + c2 x)
            const float lq_c2 = 0.5f * lg2 - lg1, lq_c1 = lg1 - lq_c2;
            const float dq_c2 = 0.5f * dg2 - dg1, dq_c1 = dg1 - dq_c2;
            const f2 LQ1 = pk(lq_c1, lq_c1), LQ2 = pk(lq_c2, lq_c2), DQ1 = pk(dq_c1, dq_c1), DQ2 = pk(dq_c2, dq_c2);
            const f2 M2 = pk(-2.f, -2.f);
            const float* wsrc = (FAST && !jok) ? S.zeros : S.wrow;
            load_x(g_begin);
            if (t == 0) TC_TRACE(1 + PHASE, it, 5);
            mbar_wait(&S.bars.acc_full[s], (it >> 1) & 1);
            tc_fence_after();
            if (t == 0) TC_TRACE(1 + PHASE, it, 6);
            if (FAST && GRAD && g_begin < g_end) {  // the coefficient tile of the previous step was consumed
                mbar_wait(&S.bars.x_empty, (it & 1) ^ 1);
                if (t == 0) TC_TRACE(1 + PHASE, it, 7);
            }
#pragma unroll 1
            for (int g = g_begin; g < g_end; ++g) {
                const int c0 = g * 8;
                uint32_t d[8];
                tmem_ld8(tmem_base + lane_addr + COL_ACC + s * 128 + c0, d);
                // per-cell scalars of the group (vector shared-memory loads)
                float lsev[8], auxv[8], wv[8];  // aux = library size (pass B) or c_i (pass C)
                int bv[8];
                if (PHASE == 0) {
                    const float4 w0 = *reinterpret_cast<const float4*>(&wsrc[c0]);
                    const float4 w1 = *reinterpret_cast<const float4*>(&wsrc[c0 + 4]);
                    wv[0] = w0.x, wv[1] = w0.y, wv[2] = w0.z, wv[3] = w0.w;
                    wv[4] = w1.x, wv[5] = w1.y, wv[6] = w1.z, wv[7] = w1.w;
                }
                {
                    const float4 l0 = *reinterpret_cast<const float4*>(&S.lse[c0]);
                    const float4 l1 = *reinterpret_cast<const float4*>(&S.lse[c0 + 4]);
                    const float* auxp = PHASE == 0 ? S.l : S.c;
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
                if (!FAST && GRAD && g == g_begin) {
                    mbar_wait(&S.bars.x_empty, (it & 1) ^ 1);
                    if (t == 0) TC_TRACE(1 + PHASE, it, 7);
                }
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
                        up(mul2(mul2(x2, add2(x2, M1)), add2(x2, M2)), p0, p1);  // x (x - 1) (x - 2)
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
                            } else if (!(x == 0.f || x == 1.f || x == 2.f)) {
                                float lg_t, dg_t;  // (locals: the out-of-line slow path takes addresses)
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
                        const f2 lgq2 = mul2(xq2, fma2(LQ2, xq2, LQ1));
                        const f2 lp2 = fma2(TH2, dlt2, fma2(x2, xl2, lgq2));
                        loss2 = fma2(W2, lp2, loss2);
                        if (GRAD) {
                            const f2 rat2 = mul2(mu2, mul2(r2, mt2));
                            const f2 dgq2 = mul2(xq2, fma2(DQ2, xq2, DQ1));
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
                } else if constexpr (NB1 && FAM == GLUE_NB) {
                    // ---- pass B, NB: packed pairs of cells in three stages ---------------------------
                    // (-G accumulated: nG = W (-(x - (x + theta) sig)) ratio; signs fixed at the flush)
                    f2 dot2[4], xs2[4], dlt2[4], xl2[4], ndd2[4], rat2[4], W2[4];
                    float lgv[8], dgv[8], xsc[8];
#pragma unroll
                    for (int pr_ = 0; pr_ < 4; ++pr_) {
                        const int e0 = 2 * pr_, e1 = e0 + 1;
                        float w01[2];
#pragma unroll
                        for (int q = 0; q < 2; ++q) {
                            const int e8 = e0 + q;
                            const float x = xv[e8];
                            const bool isnan_x = x != x;
                            bool valid = jok;  // cells beyond ncell carry weight 0 already
                            if (skip_nan && isnan_x) {
                                bad_acc += (valid && c0 + e8 < ncell) ? 1 : 0;
                                valid = false;
                            }
                            xsc[e8] = (skip_nan && isnan_x) ? 0.f : x;
                            w01[q] = valid ? wv[e8] : 0.f;
                            lgv[e8] = xsc[e8] == 1.f ? lg1 : (xsc[e8] == 2.f ? lg2 : 0.f);
                            dgv[e8] = xsc[e8] == 1.f ? dg1 : (xsc[e8] == 2.f ? dg2 : 0.f);
                        }
                        W2[pr_] = pk(w01[0], w01[1]);
                        xs2[pr_] = pk(xsc[e0], xsc[e1]);
                        dot2[pr_] = pk(__uint_as_float(d[e0]), __uint_as_float(d[e1]));
                        const f2 t2 = add2(fma2(S2, dot2[pr_], B2), pk(lsev[e0], lsev[e1]));
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
                        rat2[pr_] = mul2(mu2, mul2(r2, mt2));
                        dlt2[pr_] = fma2(l2t, NLN2, DISP2);               // log theta - log(m + theta)
                        xl2[pr_] = mul2(fma2(l2t, M1, l2m), LN2_2);       // log m - log(m + theta)
                        ndd2[pr_] = fma2(add2(xs2[pr_], TH2), sig2, mul2(xs2[pr_], M1));  // -(x - (x+theta) sig)
                    }
#pragma unroll
                    for (int e8 = 0; e8 < 8; ++e8)  // counts other than 0, 1, 2: rare
                        if (!(xsc[e8] == 0.f || xsc[e8] == 1.f || xsc[e8] == 2.f)) {
                            float lg_t, dg_t;  // (locals: the out-of-line slow path takes addresses)
                            nb_gamma_terms_t<true>(xsc[e8], fp.th, lg_t, dg_t);
                            lgv[e8] = lg_t, dgv[e8] = dg_t;
                        }
#pragma unroll
                    for (int pr_ = 0; pr_ < 4; ++pr_) {
                        const int e0 = 2 * pr_, e1 = e0 + 1;
                        const f2 lp2 = fma2(TH2, dlt2[pr_], fma2(xs2[pr_], xl2[pr_], pk(lgv[e0], lgv[e1])));
                        const f2 tl2 = fma2(TH2, add2(dlt2[pr_], pk(dgv[e0], dgv[e1])), ndd2[pr_]);
                        loss2 = fma2(W2[pr_], lp2, loss2);
                        if (GRAD) {
                            const f2 nG2 = mul2(W2[pr_], mul2(ndd2[pr_], rat2[pr_]));
                            ngb2 = add2(ngb2, nG2);
                            ngs2 = fma2(nG2, dot2[pr_], ngs2);
                            gd2 = fma2(W2[pr_], tl2, gd2);
                            up(mul2(nG2, NS2), coef[e0], coef[e1]);
                            float g0, g1;
                            up(nG2, g0, g1);
                            gch[e0] = -g0, gch[e1] = -g1;
                        } else {
                            coef[e0] = coef[e1] = 0.f;
                            gch[e0] = gch[e1] = 0.f;
                        }
                    }
                } else if constexpr (NB1 && NBF) {
                    // ---- pass B, NB / ZINB, straight-line in three stages -------------------------
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
                            float lg_t, dg_t;  // (locals: the out-of-line slow path takes addresses)
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
                    // coefficient tile -> shared memory (bf16 hi / lo): rows = features, 128 B = 64 cells
                    uint32_t hh[4], ll[4];
#pragma unroll
                    for (int e2 = 0; e2 < 4; ++e2) split2(coef[2 * e2], coef[2 * e2 + 1], hh[e2], ll[e2]);
                    const uint32_t off = (uint32_t)(c0 >> 6) * OP_BYTES + swz((uint32_t)frow, (uint32_t)((c0 & 63) >> 3));
                    st_shared_v4(smem_u32(S.x_hi) + off, hh[0], hh[1], hh[2], hh[3]);
                    st_shared_v4(smem_u32(S.x_lo) + off, ll[0], ll[1], ll[2], ll[3]);
                    if (NBF && PHASE == 0 && !FAST) {
                        // c_i partial: transpose-reduce the (32 feature lanes x 8 cells) block over lanes;
                        // afterwards lane L holds the warp total of cell c0 + 4 bit4(L) + 2 bit3(L) + bit2(L)
#pragma unroll
                        for (int off2 = 16, n = 4; n >= 1; off2 >>= 1, n >>= 1) {
                            const bool up = (lane & off2) != 0;
#pragma unroll
                            for (int k = 0; k < n; ++k) {
                                const float send = up ? gch[k] : gch[k + n];
                                const float keep = up ? gch[k + n] : gch[k];
                                gch[k] = keep + __shfl_xor_sync(FULL, send, off2);
                            }
                        }
                        float tot = gch[0];
                        tot += __shfl_xor_sync(FULL, tot, 2);
                        tot += __shfl_xor_sync(FULL, tot, 1);
                        if ((lane & 3) == 0) S.cw[it & 1][q][c0 + (lane >> 2)] = tot;
                    }
                }
            }
            if (NB1) {  // packed accumulators of the fast paths (they hold -G sums)
                float a0, a1;
                up(ngb2, a0, a1);
                gb -= a0 + a1;
                up(ngs2, a0, a1);
                gs -= a0 + a1;
                up(gd2, a0, a1);
                gd += a0 + a1;
                up(loss2, a0, a1);
                loss_acc += a0 + a1;
            }
            flush();
            if (t == 0) TC_TRACE(1 + PHASE, it, 8);
            if (t == 128) TC_TRACE(1 + PHASE, it, 9);
            if (GRAD) fence_proxy_async();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&S.bars.acc_empty[s]);
                if (GRAD) mbar_arrive(&S.bars.x_full);
            }
            if (GRAD) {
                named_bar_sync(1, EPI_T);
                if (t == 0) TC_TRACE(1 + PHASE, it, 10);
                if (pend) write_param(pend_off, pend_b, pend_s, pend_d, pend_z, true);
                if (NBF && PHASE == 0 && !FAST && t < ncell)
                    S.c_acc[t] += (S.cw[it & 1][0][t] + S.cw[it & 1][1][t]) + (S.cw[it & 1][2][t] + S.cw[it & 1][3][t]);
            }
        }
        if (GRAD) {
            // d/du partial of this CTA: rows = cells
            mbar_wait(&S.bars.d2_full, 0);
            tc_fence_after();
            if (t == 0) TC_TRACE(1 + PHASE, 6, 14);
            uint32_t r[DC];
            tmem_ld32(tmem_base + lane_addr + COL_D2 + DC * h, r);
            tmem_ld_wait();
            float4* dst = reinterpret_cast<float4*>(P.p.ws.apart + ((int64_t)cta * 128 + frow) * 64 + DC * h);
            float4 old[DC / 4];
#pragma unroll
            for (int k4 = 0; k4 < DC / 4; ++k4) old[k4] = PHASE == 1 ? dst[k4] : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int k4 = 0; k4 < DC / 4; ++k4)
                dst[k4] = make_float4(__uint_as_float(r[4 * k4]) + old[k4].x, __uint_as_float(r[4 * k4 + 1]) + old[k4].y,
                                      __uint_as_float(r[4 * k4 + 2]) + old[k4].z,
                                      __uint_as_float(r[4 * k4 + 3]) + old[k4].w);
            if (t == 0) TC_TRACE(1 + PHASE, 6, 15);
            if (NBF && PHASE == 0) {
                if constexpr (FAST) {
                    // c_i partial of this CTA = column K of the d/du accumulator (produce_v_rows<CINV>);
                    // rows of D2 are cells, so the thread that holds column K of cell `frow` writes it
                    float cval = 0.f;
#pragma unroll
                    for (int k = 0; k < DC; ++k)
                        if (DC * h + k == a.K) cval = __uint_as_float(r[k]);
                    if (a.K / DC == h && frow < ncell) P.p.ws.cpart[(int64_t)cta * a.B + P.cell0 + frow] = cval;
                } else {
                    named_bar_sync(1, EPI_T);
                    if (t < ncell) P.p.ws.cpart[(int64_t)cta * a.B + P.cell0 + t] = S.c_acc[t];
                }
            }
        }
        if (PHASE == 0) {
            const float wl = warp_sum(loss_acc);
            int wb = bad_acc;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) wb += __shfl_xor_sync(FULL, wb, o);
            if (lane == 0) S.red_l[warp] = wl, S.red_b[warp] = wb;
            named_bar_sync(2, EPI_T);
            if (t == 0) {
                float sl = 0.f;
                int sb = 0;
                for (int w = 0; w < EW; ++w) sl += S.red_l[w], sb += S.red_b[w];
                P.p.ws.losspart[P.chunk * (int)gridDim.x + cta] = sl;
                P.p.ws.badpart[P.chunk * (int)gridDim.x + cta] = sb;
            }
        }
    }
    tc_fence_before();
    if (PHASE == 0) __threadfence();
    __syncthreads();
    if (warp == MMAW) tmem_dealloc(tmem_base, TMEM_COLS);
    if (PHASE == 0) {
        // ---- the last CTA to arrive reduces the per-CTA partials: c_i, loss, nanmean factor -----
        if (t == 0) S.ticket = atomicAdd(P.p.ws.counters + 1, 1u);
        __syncthreads();
        if (S.ticket == (unsigned)(P.chunk + 1) * gridDim.x - 1) {  // tickets keep counting across cell blocks
            __threadfence();
            const int nc = (int)gridDim.x;
            if (GRAD && NBF && t < 384) {  // three strided sub-sums per cell, combined in a fixed order
                const int cell = t & 127, sub = t >> 7;
                float c = 0.f;
                if (cell < ncell) {
#pragma unroll 10
                    for (int q = sub; q < nc; q += 3) c += __ldcg(P.p.ws.cpart + (int64_t)q * a.B + P.cell0 + cell);
                }
                S.cpart3[sub][cell] = c;
            }
            if (warp == MMAW && P.chunk == P.nchunks - 1) {  // loss of the whole call: all cell blocks
                double loss = 0.0;
                long long bad = 0;
                for (int q = lane; q < nc * P.nchunks; q += 32) {
                    loss += (double)__ldcg(P.p.ws.losspart + q);
                    bad += __ldcg(P.p.ws.badpart + q);
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    loss += __shfl_xor_sync(FULL, loss, o);
                    bad += __shfl_xor_sync(FULL, bad, o);
                }
                if (lane == 0) {
                    const double total = (double)a.B * (double)a.F;
                    const double ratio = bad > 0 ? total / (total - (double)bad) : 1.0;
                    P.p.ws.rescale[0] = (float)ratio;
                    if (a.loss) a.loss[0] = (float)(loss * ratio / (double)P.p.gscale);
                }
            }
            __syncthreads();
            if (GRAD && NBF && t < ncell)
                P.p.ws.c[P.cell0 + t] = (S.cpart3[0][t] + S.cpart3[1][t]) + S.cpart3[2][t];
        }
    }
    if (t == 0) TC_TRACE(1 + PHASE, 7, 1);
}

// lse_i from the per-(CTA, column half) partials.  Block = 32 cells x 32 strided sub-ranges of
// the partials (consecutive lanes read consecutive cells, <= 10 loads in flight per thread);
// sub-results meet in a fixed order.
__global__ void __launch_bounds__(1024) tc_lse_kernel(DecWs ws, int nparts, int B) {
    __shared__ float red[2][32][33];
    griddep_launch();
    griddep_wait();
    const int lane = threadIdx.x & 31, sub = threadIdx.x >> 5;
    const int row = blockIdx.x * 32 + lane;
    const bool act = row < B;
    float pm[10], ps[10];
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        const int c = sub + 32 * i;
        const bool ok = act && c < nparts;
        pm[i] = ok ? ws.pmax[(int64_t)c * B + row] : -INFINITY;
        ps[i] = ok ? ws.psum[(int64_t)c * B + row] : 0.f;
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
    for (int i = 0; i < 10; ++i) sum += pm[i] > -INFINITY ? ps[i] * expf(pm[i] - mx) : 0.f;
    red[1][sub][lane] = sum;
    __syncthreads();
    if (act && sub == 0) {
        float tot = 0.f;
#pragma unroll 8
        for (int q = 0; q < 32; ++q) tot += red[1][q][lane];
        ws.lse[row] = mx + logf(tot);
    }
}

// d/du = sum over CTAs of the D2 partials; grid = B blocks (sorted position), 256 threads =
// 64 latent columns x 4 strided sub-sums combined in a fixed order
__global__ void __launch_bounds__(256) tc_grad_u_kernel(DecParams p, int ncta, int cell0) {
    const glue_decoder_args& a = p.a;
    const int pos = blockIdx.x, k = threadIdx.x & 63, sub = threadIdx.x >> 6;  // pos: cell within the block
    __shared__ float part[4][64];
    griddep_launch();
    griddep_wait();
    float acc = 0.f;
    for (int q = sub; q < ncta; q += 4) acc += p.ws.apart[((int64_t)q * 128 + pos) * 64 + k];
    part[sub][k] = acc;
    __syncthreads();
    if (sub == 0 && k < a.K && a.grad_u)
        a.grad_u[(int64_t)orig_row(a, cell0 + pos) * a.K + k] =
            (part[0][k] + part[1][k]) + (part[2][k] + part[3][k]);
}

size_t tc_ws_layout(const glue_decoder_args& a, int ncta, char* base, DecWs* ws) {
    size_t off = 0;
    auto take = [&](size_t bytes) {
        char* ptr = base ? base + off : nullptr;
        off += align_up(bytes, 256);
        return ptr;
    };
    const size_t nb = (size_t)ncta * a.B;
    float* pmax = (float*)take(2 * nb * 4);
    float* psum = (float*)take(2 * nb * 4);
    float* lse = (float*)take((size_t)a.B * 4);
    float* cpart = (float*)take(nb * 4);
    float* c = (float*)take((size_t)a.B * 4);
    float* apart = (float*)take((size_t)ncta * 128 * 64 * 4);
    const size_t nchunks = ((size_t)a.B + 127) / 128;
    float* losspart = (float*)take(nchunks * ncta * 4);
    int* badpart = (int*)take(nchunks * ncta * 4);
    float* rescale = (float*)take(4);
    unsigned int* counters = (unsigned int*)take(16);
    if (ws) *ws = DecWs{pmax, psum, lse, cpart, c, apart, losspart, badpart, rescale, counters};
    return off;
}

int tc_grid(const glue_decoder_args& a, int& ntiles) {
    // the fewest CTAs with the makespan of a full grid (see grid_for in decoder_tc3.cu)
    ntiles = (a.F + 127) / 128;
    const int sms = sm_count();
    if (ntiles <= sms) return ntiles;
    const int per_cta = (ntiles + sms - 1) / sms;
    return (ntiles + per_cta - 1) / per_cta;
}

template <int FAM, int PHASE, bool GRAD, bool NB1, bool FAST>
int launch_kernel(const TcParams& P, int ncta, cudaStream_t st) {
    const size_t smem = sizeof(SmemMain);
    auto k = tc_main_kernel<FAM, PHASE, GRAD, NB1, FAST>;
    GLUE_SMEM_ONCE(k, smem);
    GLUE_CUDA(launch_pdl(P.pdl != 0, k, dim3(ncta), dim3(NTHREADS), smem, st, P));
    GLUE_CHECK_LAUNCH();
    return 0;
}

template <int FAM, bool NB1, bool FAST>
int launch_phase(int phase, bool grad, const TcParams& P, int ncta, cudaStream_t st) {
    if (phase == 1) return launch_kernel<FAM, 1, true, NB1, false>(P, ncta, st);
    if (grad) return launch_kernel<FAM, 0, true, NB1, FAST>(P, ncta, st);
    return launch_kernel<FAM, 0, false, NB1, FAST>(P, ncta, st);
}

// variant bits (GLUE_TC_VARIANT): bit 0 = streamlined NB pass B (single parameter row)
constexpr int VARIANT_FAST = 1;
constexpr int TC_DEFAULT_PDL = 0;
constexpr int TC_DEFAULT_VARIANT = VARIANT_FAST;  // measured: 117.7 -> 99.1 us per ATAC fwd+bwd call

template <bool NB1>
int launch_family(int family, int phase, bool grad, int variant, const TcParams& P, int ncta, cudaStream_t st) {
    if (phase == 1) return launch_phase<GLUE_NB, NB1, false>(1, true, P, ncta, st);  // pass C is family independent
    if constexpr (NB1) {
        if (family == GLUE_NB && (variant & VARIANT_FAST)) return launch_phase<GLUE_NB, true, true>(0, grad, P, ncta, st);
    }
    switch (family) {
        case GLUE_NB: return launch_phase<GLUE_NB, NB1, false>(0, grad, P, ncta, st);
        case GLUE_ZINB: return launch_phase<GLUE_ZINB, NB1, false>(0, grad, P, ncta, st);
        case GLUE_NORMAL: return launch_phase<GLUE_NORMAL, NB1, false>(0, grad, P, ncta, st);
        default: return launch_phase<GLUE_ZILN, NB1, false>(0, grad, P, ncta, st);
    }
}

}  // namespace

static long long* g_trace = nullptr;

size_t tc_workspace_bytes(const glue_decoder_args& a) {
    int ntiles;
    const int ncta = tc_grid(a, ntiles);
    return tc_ws_layout(a, ncta, nullptr, nullptr);
}

bool tc_eligible(const glue_decoder_args& a, int mode) {
    if (mode != MODE_LOSS && mode != MODE_GRAD) return false;
    if (a.wmat) return false;
    if (a.B > 128 * TC_MAX_CHUNKS || a.K > 64 || (a.K & 1) || a.n_batches > NB_MAX) return false;
    if (a.n_batches > 1 && a.b && !a.row_order) return false;
    auto al8 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 7u) == 0; };
    if (!al8(a.u) || !al8(a.v) || !al8(a.grad_v)) return false;
    return true;
}

int tc_run(DecParams p, int mode, cudaStream_t st) {
    const glue_decoder_args& a = p.a;
    TcParams P;
    int ntiles;
    const int ncta = tc_grid(a, ntiles);
    const size_t need = tc_ws_layout(a, ncta, nullptr, nullptr);
    if (a.workspace_bytes < need) return GLUE_E_WORKSPACE;
    tc_ws_layout(a, ncta, (char*)a.workspace, &p.ws);
    p.ncta = ncta;
    p.gscale = a.grad_scale != 0.f ? a.grad_scale : 1.f;
    p.w0 = (float)(-1.0 / ((double)a.B * (double)a.F)) * p.gscale;
    const int nchunks = (a.B + 127) / 128;
    P.p = p, P.ntiles = ntiles, P.cell0 = 0, P.ncell = a.B, P.acc = 0, P.trace = g_trace;
    P.chunk = 0, P.nchunks = nchunks;
    P.pdl = TC_DEFAULT_PDL;
    const bool nbf = is_nb_family(a.family);
    const bool grad = mode == MODE_GRAD;
    const bool nb1 = a.n_batches == 1;
    // kernel variant: impl bit 8 switches the streamlined NB pass B off (A/B runs, tests)
    int variant = TC_DEFAULT_VARIANT;
    if (a.impl & 0x100) variant &= ~VARIANT_FAST;
    if (!(a.family == GLUE_NB && nb1 && a.K <= 62 && (long long)a.B * (long long)a.F < (1ll << 31)))
        variant &= ~VARIANT_FAST;
    // Cells are handled in blocks of <= 128 (soft-max rows are independent); feature-side outputs
    // (d/dv, parameter tables) accumulate from the second block on, in launch order.
    auto block = [&](int c) {
        P.chunk = c, P.cell0 = 128 * c, P.ncell = a.B - 128 * c < 128 ? a.B - 128 * c : 128;
        P.acc = c > 0 ? 1 : 0;
    };

    GLUE_CUDA(cudaMemsetAsync(p.ws.counters, 0, 16, st));
    if (grad && a.n_batches > 1) {
        const size_t tab = (size_t)a.n_batches * a.F * sizeof(float);
        if (a.grad_scale_lin) GLUE_CUDA(cudaMemsetAsync(a.grad_scale_lin, 0, tab, st));
        if (a.grad_bias) GLUE_CUDA(cudaMemsetAsync(a.grad_bias, 0, tab, st));
        if (a.grad_disp) GLUE_CUDA(cudaMemsetAsync(a.grad_disp, 0, tab, st));
        if (a.grad_zi_logits) GLUE_CUDA(cudaMemsetAsync(a.grad_zi_logits, 0, tab, st));
    }
    if (nbf) {
        const size_t smem = sizeof(SmemStats);
        GLUE_SMEM_ONCE(tc_rowstats_kernel, smem);
        for (int c = 0; c < nchunks; ++c) {
            block(c);
            // (the first launch follows a memset node: plain stream order)
            GLUE_CUDA(launch_pdl(P.pdl != 0 && c > 0, tc_rowstats_kernel, dim3(ncta), dim3(NTHREADS), smem, st, P));
            GLUE_CHECK_LAUNCH();
        }
        GLUE_CUDA(launch_pdl(P.pdl != 0, tc_lse_kernel, dim3((a.B + 31) / 32), dim3(1024), 0, st, p.ws, 2 * ncta,
                             a.B));  // 2 * ncta <= 320 partials
        GLUE_CHECK_LAUNCH();
    }
    for (int c = 0; c < nchunks; ++c) {
        block(c);
        int rc = nb1 ? launch_family<true>(a.family, 0, grad, variant, P, ncta, st)
                     : launch_family<false>(a.family, 0, grad, variant, P, ncta, st);
        if (rc) return rc;
        if (grad) {
            if (nbf) {
                rc = nb1 ? launch_family<true>(a.family, 1, true, variant, P, ncta, st)
                         : launch_family<false>(a.family, 1, true, variant, P, ncta, st);
                if (rc) return rc;
            }
            GLUE_CUDA(launch_pdl(P.pdl != 0, tc_grad_u_kernel, dim3(P.ncell), dim3(256), 0, st, p, ncta, P.cell0));
            GLUE_CHECK_LAUNCH();
        }
    }
    if (grad && a.reduce == 0 && !a.row_weight) {
        dec_rescale_kernel<<<sm_count(), 256, 0, st>>>(p);
        GLUE_CHECK_LAUNCH();
    }
    return 0;
}

}  // namespace glue

// Diagnostic: device buffer of 3 * 8 * 16 int64 that receives clock64 stamps of CTA 0 of the
// tensor-core decoder kernels (NULL switches tracing off).  Not part of the training path.
extern "C" void glue_tc_set_trace(void* device_buffer) { glue::g_trace = (long long*)device_buffer; }
