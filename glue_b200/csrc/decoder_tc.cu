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
// Warp roles (384 threads, 1 CTA / SM): warps 0-7 epilogue (warp w owns TMEM lanes
// 32 (w & 3) .. +31 and one half of the cell columns), warp 8 issues tcgen05.mma (one lane),
// warps 9-11 stage the v tiles (double buffered).  Pipelines are mbarrier based.
#include "decoder_common.cuh"
#include "tc_common.cuh"

namespace glue {
using namespace tc;

namespace {

constexpr int NTHREADS = 384;
constexpr int EPI_THREADS = 256;
constexpr int EPI_WARPS = 8;
constexpr int MMA_WARP = 8;
constexpr int PROD_T0 = 288;
constexpr int PROD_THREADS = 96;
constexpr int NB_MAX = 24;  // parameter rows the row-statistics table holds per v stage

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
    float lse[128], l[128], c[128];
    int bidx[128];
    float cw[2][4][128];
    float c_acc[128];
    float red_l[EPI_WARPS];
    int red_b[EPI_WARPS];
    Bars bars;
    uint32_t tmem_base;
};

struct __align__(1024) SmemStats {
    uint8_t u_hi[OP_BYTES], u_lo[OP_BYTES];
    uint8_t v_hi[2][OP_BYTES], v_lo[2][OP_BYTES];
    float2 ptab[2][NB_MAX][130];  // (softplus(scale_lin), bias) per (stage, batch, feature-in-tile)
    int bidx[128];
    Bars bars;
    uint32_t tmem_base;
};

struct TcParams {
    DecParams p;
    int ntiles;
    int cell0, ncell;  // cell block [cell0, cell0 + ncell) of the (sorted) minibatch, ncell <= 128
    int acc;           // outputs accumulate (a previous cell block already wrote them)
};

template <class T>
__device__ __forceinline__ T& carve(uint8_t* raw) {
    return *reinterpret_cast<T*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
}

__device__ __forceinline__ int orig_row(const glue_decoder_args& a, int pos) {
    return a.row_order ? a.row_order[pos] : pos;
}

// v tile producer: one thread per feature row
__device__ __forceinline__ void produce_v_rows(const glue_decoder_args& a, uint8_t* hi, uint8_t* lo, int f0, int pt) {
    for (int r = pt; r < 128; r += PROD_THREADS) {
        const int j = f0 + r;
        const float* src = nullptr;
        if (j < a.F) {
            const int64_t vr = a.vidx ? a.vidx[j] : (int64_t)j;
            src = a.v + vr * a.K;
        }
        stage_row(smem_u32(hi), smem_u32(lo), (uint32_t)r, src, a.K);
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
    extern __shared__ uint8_t smem_raw[];
    SmemStats& S = carve<SmemStats>(smem_raw);
    const glue_decoder_args& a = P.p.a;
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const int cta = blockIdx.x, ncta = gridDim.x;
    const int ntl = (P.ntiles - cta + ncta - 1) / ncta;

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

    if (warp > MMA_WARP) {
        const int pt = t - PROD_T0;
        for (int it = 0; it < ntl; ++it) {
            const int s = it & 1, f0 = (cta + it * ncta) * 128;
            mbar_wait(&S.bars.v_empty[s], ((it >> 1) & 1) ^ 1);
            produce_v_rows(a, S.v_hi[s], S.v_lo[s], f0, pt);
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
    } else {
        const int q = warp & 3, h = warp >> 2;
        const int cell = 32 * q + lane;
        const int my_b = S.bidx[cell];
        float m = -INFINITY, ssum = 0.f;
        for (int it = 0; it < ntl; ++it) {
            const int s = it & 1, f0 = (cta + it * ncta) * 128;
            const int nvalid = min(128, a.F - f0);
            mbar_wait(&S.bars.acc_full[s], (it >> 1) & 1);
            tc_fence_after();
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
}

// ---------------------------------------------------------------------------------------------
// pass B (PHASE 0: likelihood + all G terms) and pass C (PHASE 1: the -p_ij c_i soft-max terms).
// D[feature, cell] = V U^T; thread <-> feature.
// ---------------------------------------------------------------------------------------------
template <int FAM, int PHASE, bool GRAD>
__global__ void __launch_bounds__(NTHREADS, 1) tc_main_kernel(TcParams P) {
    extern __shared__ uint8_t smem_raw[];
    SmemMain& S = carve<SmemMain>(smem_raw);
    const glue_decoder_args& a = P.p.a;
    constexpr bool NBF = (FAM == GLUE_NB || FAM == GLUE_ZINB);
    constexpr bool ZI = (FAM == GLUE_ZINB || FAM == GLUE_ZILN);
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const int cta = blockIdx.x, ncta = gridDim.x;
    const int ntl = (P.ntiles - cta + ncta - 1) / ncta;
    const int ncell = P.ncell;

    if (t == 0) {
        for (int s = 0; s < 2; ++s) {
            mbar_init(&S.bars.v_full[s], PROD_THREADS);
            mbar_init(&S.bars.v_empty[s], 1);
            mbar_init(&S.bars.acc_full[s], 1);
            mbar_init(&S.bars.acc_empty[s], EPI_WARPS);
            mbar_init(&S.bars.d1_full[s], 1);
            mbar_init(&S.bars.d1_empty[s], EPI_WARPS);
        }
        mbar_init(&S.bars.x_full, EPI_WARPS);
        mbar_init(&S.bars.x_empty, 1);
        mbar_init(&S.bars.d2_full, 1);
        mbar_fence_init();
    }
    if (warp == MMA_WARP) tmem_alloc(&S.tmem_base, TMEM_COLS);
    if (t < 128) {
        const bool ok = t < ncell;
        const int pos = P.cell0 + t;
        const int orig = ok ? orig_row(a, pos) : 0;
        stage_row(smem_u32(S.u_hi), smem_u32(S.u_lo), (uint32_t)t, ok ? a.u + (int64_t)orig * a.K : nullptr, a.K);
        S.bidx[t] = (ok && a.b) ? (int)a.b[orig] : 0;
        S.xoff[t] = (long long)orig * a.F;
        S.l[t] = (ok && a.l) ? a.l[orig] : 1.f;
        S.lse[t] = (ok && NBF) ? P.p.ws.lse[pos] : 0.f;
        S.c[t] = (ok && PHASE == 1) ? P.p.ws.c[pos] : 0.f;
        S.c_acc[t] = 0.f;
    }
    if (GRAD) {  // cells beyond ncell must read as exact zeros in the coefficient tile
        uint4* xz = reinterpret_cast<uint4*>(S.x_hi);
        for (int i = t; i < (int)(4 * OP_BYTES / 16); i += NTHREADS) xz[i] = make_uint4(0, 0, 0, 0);
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = S.tmem_base;

    if (warp > MMA_WARP) {
        // ===== v tile producers =====
        const int pt = t - PROD_T0;
        for (int it = 0; it < ntl; ++it) {
            const int s = it & 1, f0 = (cta + it * ncta) * 128;
            mbar_wait(&S.bars.v_empty[s], ((it >> 1) & 1) ^ 1);
            produce_v_rows(a, S.v_hi[s], S.v_lo[s], f0, pt);
            fence_proxy_async();
            mbar_arrive(&S.bars.v_full[s]);
        }
    } else if (warp == MMA_WARP) {
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
            };
            issue_fwd(0);
            const uint32_t nks1 = (uint32_t)(ncell + 15) >> 4;  // K steps over cells
            for (int it = 0; it < ntl; ++it) {
                if (it + 1 < ntl) issue_fwd(it + 1);
                if (GRAD) {
                    const int s = it & 1;
                    mbar_wait(&S.bars.x_full, it & 1);
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
                    }
                }
            }
            if (GRAD) umma_commit(&S.bars.d2_full);
        }
    } else {
        // ===== epilogue =====
        const int q = warp & 3, h = warp >> 2;
        const int frow = 32 * q + lane;
        const uint32_t lane_addr = (uint32_t)(32 * q) << 16;
        const int nch = (ncell + 31) >> 5;
        const int ch_begin = h == 0 ? 0 : (nch + 1) / 2;
        const int ch_end = h == 0 ? (nch + 1) / 2 : nch;
        const bool add_v = (PHASE == 1) || P.acc;
        float loss_acc = 0.f;
        int bad_acc = 0;

        auto readout_d1 = [&](int itp) {
            const int s = itp & 1;
            const int j = (cta + itp * ncta) * 128 + frow;
            mbar_wait(&S.bars.d1_full[s], (itp >> 1) & 1);
            tc_fence_after();
            uint32_t r[32];
            tmem_ld32(tmem_base + lane_addr + COL_D1 + s * 64 + 32 * h, r);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&S.bars.d1_empty[s]);
            if (j < a.F && a.grad_v) {
                const int64_t vr = a.vidx ? a.vidx[j] : (int64_t)j;
                float* dst = a.grad_v + vr * a.K + 32 * h;
#pragma unroll
                for (int k2 = 0; k2 < 16; ++k2) {
                    if (32 * h + 2 * k2 < a.K) {
                        float2 val = make_float2(__uint_as_float(r[2 * k2]), __uint_as_float(r[2 * k2 + 1]));
                        float2* d2p = reinterpret_cast<float2*>(dst + 2 * k2);
                        if (add_v) {
                            const float2 old = *d2p;
                            val.x += old.x, val.y += old.y;
                        }
                        *d2p = val;
                    }
                }
            }
        };

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
                if (a.grad_bias) a.grad_bias[off] = add ? a.grad_bias[off] + vb : vb;
                if (a.grad_scale_lin) a.grad_scale_lin[off] = add ? a.grad_scale_lin[off] + vs : vs;
                if (PHASE == 0) {
                    if (a.grad_disp) a.grad_disp[off] = add ? a.grad_disp[off] + vd : vd;
                    if (ZI && a.grad_zi_logits) a.grad_zi_logits[off] = add ? a.grad_zi_logits[off] + vz : vz;
                }
            };
            auto flush = [&]() {
                if (GRAD && cur_b >= 0 && jok) {
                    const int64_t off = (int64_t)cur_b * a.F + j;
                    const float vs = gs * fp.sg;
                    if (h == 1 && !pend) {
                        pend = true, pend_off = off, pend_b = gb, pend_s = vs, pend_d = gd, pend_z = gz;
                    } else {
                        write_param(off, gb, vs, gd, gz, (PHASE == 1) || P.acc || a.n_batches > 1 || h == 1);
                    }
                }
                gb = gs = gd = gz = 0.f;
            };

            mbar_wait(&S.bars.acc_full[s], (it >> 1) & 1);
            tc_fence_after();
#pragma unroll 1
            for (int cc = ch_begin; cc < ch_end; ++cc) {
                const int c0 = cc * 32;
                uint32_t d[32];
                tmem_ld32(tmem_base + lane_addr + COL_ACC + s * 128 + c0, d);
                float xv[32];
                if (PHASE == 0) {
#pragma unroll
                    for (int r = 0; r < 32; ++r)
                        xv[r] = (jok && c0 + r < ncell) ? __ldg(a.x + S.xoff[c0 + r] + j) : 0.f;
                }
                tmem_ld_wait();
                if (GRAD && cc == ch_begin) mbar_wait(&S.bars.x_empty, (it & 1) ^ 1);
                float gch[32];
                const uint32_t xh = smem_u32(S.x_hi) + (uint32_t)(c0 >> 6) * OP_BYTES;
                const uint32_t xl = smem_u32(S.x_lo) + (uint32_t)(c0 >> 6) * OP_BYTES;
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    float coef[8];
#pragma unroll
                    for (int e8 = 0; e8 < 8; ++e8) {
                        const int r = 8 * g + e8;
                        const int cell = c0 + r;
                        const bool cok = cell < ncell;  // warp-uniform
                        float G = 0.f, cf = 0.f;
                        if (cok) {
                            const int bi = S.bidx[cell];
                            if (bi != cur_b) {
                                flush();
                                fp = load_params<FAM>(a, bi, j, jok);
                                cur_b = bi;
                            }
                            const float dot = __uint_as_float(d[r]);
                            const float av = fmaf(fp.s, dot, fp.bias);
                            if (PHASE == 0) {
                                const float x = xv[r];
                                const ElemOut e = element<FAM>(fp, av, x, S.lse[cell], S.l[cell]);
                                bool valid = jok;
                                if (a.reduce == 0 && x != x) {
                                    valid = false;
                                    bad_acc += jok ? 1 : 0;
                                }
                                const float W = valid ? P.p.w0 : 0.f;
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
                                const float pr = jok ? expf(av - S.lse[cell]) : 0.f;
                                const float w = pr * S.c[cell];
                                G = -w;
                                cf = G * fp.s;
                                gb += G;
                                gs += G * dot;
                            }
                        }
                        coef[e8] = cf;
                        gch[r] = G;
                    }
                    if (GRAD) {
                        // coefficient tile -> shared memory (bf16 hi / lo): rows = features, 128 B = 64 cells
                        uint32_t hh[4], ll[4];
#pragma unroll
                        for (int e2 = 0; e2 < 4; ++e2) split2(coef[2 * e2], coef[2 * e2 + 1], hh[e2], ll[e2]);
                        const uint32_t off = swz((uint32_t)frow, (uint32_t)(((c0 & 63) >> 3) + g));
                        st_shared_v4(xh + off, hh[0], hh[1], hh[2], hh[3]);
                        st_shared_v4(xl + off, ll[0], ll[1], ll[2], ll[3]);
                    }
                }
                if (GRAD) {
                    if (NBF && PHASE == 0) {
                        // c_i partial: transpose-reduce the 32 x 32 (feature lane x cell) block over lanes
#pragma unroll
                        for (int off = 16; off >= 1; off >>= 1) {
                            const bool up = (lane & off) != 0;
#pragma unroll
                            for (int k = 0; k < off; ++k) {
                                const float send = up ? gch[k] : gch[k + off];
                                const float keep = up ? gch[k + off] : gch[k];
                                gch[k] = keep + __shfl_xor_sync(FULL, send, off);
                            }
                        }
                        S.cw[it & 1][q][c0 + lane] = gch[0];
                    }
                }
            }
            flush();
            if (GRAD) fence_proxy_async();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&S.bars.acc_empty[s]);
                if (GRAD) mbar_arrive(&S.bars.x_full);
            }
            if (GRAD) {
                named_bar_sync(1, EPI_THREADS);
                if (pend) write_param(pend_off, pend_b, pend_s, pend_d, pend_z, true);
                if (NBF && PHASE == 0 && t < ncell)
                    S.c_acc[t] += (S.cw[it & 1][0][t] + S.cw[it & 1][1][t]) + (S.cw[it & 1][2][t] + S.cw[it & 1][3][t]);
                if (it > 0) readout_d1(it - 1);
            }
        }
        if (GRAD) {
            readout_d1(ntl - 1);
            // d/du partial of this CTA: rows = cells
            mbar_wait(&S.bars.d2_full, 0);
            tc_fence_after();
            uint32_t r[32];
            tmem_ld32(tmem_base + lane_addr + COL_D2 + 32 * h, r);
            tmem_ld_wait();
            float4* dst = reinterpret_cast<float4*>(P.p.ws.apart + ((int64_t)cta * 128 + frow) * 64 + 32 * h);
#pragma unroll
            for (int k4 = 0; k4 < 8; ++k4) {
                float4 val = make_float4(__uint_as_float(r[4 * k4]), __uint_as_float(r[4 * k4 + 1]),
                                         __uint_as_float(r[4 * k4 + 2]), __uint_as_float(r[4 * k4 + 3]));
                if (PHASE == 1) {
                    const float4 old = dst[k4];
                    val.x += old.x, val.y += old.y, val.z += old.z, val.w += old.w;
                }
                dst[k4] = val;
            }
            if (NBF && PHASE == 0) {
                named_bar_sync(1, EPI_THREADS);
                if (t < ncell) P.p.ws.cpart[(int64_t)cta * a.B + P.cell0 + t] = S.c_acc[t];
            }
        }
        if (PHASE == 0) {
            const float wl = warp_sum(loss_acc);
            int wb = bad_acc;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) wb += __shfl_xor_sync(FULL, wb, o);
            if (lane == 0) S.red_l[warp] = wl, S.red_b[warp] = wb;
            named_bar_sync(2, EPI_THREADS);
            if (t == 0) {
                float sl = 0.f;
                int sb = 0;
                for (int w = 0; w < EPI_WARPS; ++w) sl += S.red_l[w], sb += S.red_b[w];
                P.p.ws.losspart[cta] = sl;
                P.p.ws.badpart[cta] = sb;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == MMA_WARP) tmem_dealloc(tmem_base, TMEM_COLS);
}

// c_i = sum over CTAs of the partials; loss and the nanmean rescale factor (1 block of 128 threads)
__global__ void tc_finish_main_kernel(DecParams p, int ncta, int nbf, int grad) {
    const glue_decoder_args& a = p.a;
    const int t = threadIdx.x;
    if (grad && nbf)
        for (int i = t; i < a.B; i += blockDim.x) {
            float c = 0.f;
            for (int q = 0; q < ncta; ++q) c += p.ws.cpart[(int64_t)q * a.B + i];
            p.ws.c[i] = c;
        }
    if (t == 0) {
        double loss = 0.0;
        long long bad = 0;
        for (int q = 0; q < ncta; ++q) {
            loss += (double)p.ws.losspart[q];
            bad += p.ws.badpart[q];
        }
        const double total = (double)a.B * (double)a.F;
        const double ratio = bad > 0 ? total / (total - (double)bad) : 1.0;
        p.ws.rescale[0] = (float)ratio;
        if (a.loss) a.loss[0] = (float)(loss * ratio);
    }
}

// d/du = sum over CTAs of the D2 partials; grid = B blocks (sorted position), 64 threads
__global__ void tc_grad_u_kernel(DecParams p, int ncta) {
    const glue_decoder_args& a = p.a;
    const int pos = blockIdx.x, k = threadIdx.x;
    if (k >= a.K || !a.grad_u) return;
    float acc = 0.f;
    for (int q = 0; q < ncta; ++q) acc += p.ws.apart[((int64_t)q * 128 + pos) * 64 + k];
    a.grad_u[(int64_t)orig_row(a, pos) * a.K + k] = acc;
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
    float* losspart = (float*)take((size_t)ncta * 4);
    int* badpart = (int*)take((size_t)ncta * 4);
    float* rescale = (float*)take(4);
    if (ws) *ws = DecWs{pmax, psum, lse, cpart, c, apart, losspart, badpart, rescale};
    return off;
}

int tc_grid(const glue_decoder_args& a, int& ntiles) {
    ntiles = (a.F + 127) / 128;
    return ntiles < sm_count() ? ntiles : sm_count();
}

template <int FAM>
int launch_phase(int phase, bool grad, const TcParams& P, int ncta, cudaStream_t st) {
    const size_t smem = sizeof(SmemMain) + 1024;
#define GLUE_TC_LAUNCH(PH, GR)                                                                      \
    {                                                                                               \
        auto k = tc_main_kernel<FAM, PH, GR>;                                                       \
        GLUE_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        k<<<ncta, NTHREADS, smem, st>>>(P);                                                         \
    }
    if (phase == 1) GLUE_TC_LAUNCH(1, true)
    else if (grad) GLUE_TC_LAUNCH(0, true)
    else GLUE_TC_LAUNCH(0, false)
#undef GLUE_TC_LAUNCH
    GLUE_CHECK_LAUNCH();
    return 0;
}

}  // namespace

size_t tc_workspace_bytes(const glue_decoder_args& a) {
    int ntiles;
    const int ncta = tc_grid(a, ntiles);
    return tc_ws_layout(a, ncta, nullptr, nullptr);
}

bool tc_eligible(const glue_decoder_args& a, int mode) {
    if (mode != MODE_LOSS && mode != MODE_GRAD) return false;
    if (a.wmat) return false;
    if (a.B > 128 || a.K > 64 || (a.K & 1) || a.n_batches > NB_MAX) return false;
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
    p.w0 = (float)(-1.0 / ((double)a.B * (double)a.F));
    P.p = p, P.ntiles = ntiles, P.cell0 = 0, P.ncell = a.B, P.acc = 0;
    const bool nbf = is_nb_family(a.family);
    const bool grad = mode == MODE_GRAD;

    if (grad && a.n_batches > 1) {
        const size_t tab = (size_t)a.n_batches * a.F * sizeof(float);
        if (a.grad_scale_lin) GLUE_CUDA(cudaMemsetAsync(a.grad_scale_lin, 0, tab, st));
        if (a.grad_bias) GLUE_CUDA(cudaMemsetAsync(a.grad_bias, 0, tab, st));
        if (a.grad_disp) GLUE_CUDA(cudaMemsetAsync(a.grad_disp, 0, tab, st));
        if (a.grad_zi_logits) GLUE_CUDA(cudaMemsetAsync(a.grad_zi_logits, 0, tab, st));
    }
    if (nbf) {
        const size_t smem = sizeof(SmemStats) + 1024;
        GLUE_CUDA(cudaFuncSetAttribute(tc_rowstats_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        tc_rowstats_kernel<<<ncta, NTHREADS, smem, st>>>(P);
        GLUE_CHECK_LAUNCH();
        dec_lse_kernel<<<(a.B * 32 + 255) / 256, 256, 0, st>>>(p.ws, 2 * ncta, a.B);
        GLUE_CHECK_LAUNCH();
    }
    int rc = 0;
    switch (a.family) {
        case GLUE_NB: rc = launch_phase<GLUE_NB>(0, grad, P, ncta, st); break;
        case GLUE_ZINB: rc = launch_phase<GLUE_ZINB>(0, grad, P, ncta, st); break;
        case GLUE_NORMAL: rc = launch_phase<GLUE_NORMAL>(0, grad, P, ncta, st); break;
        default: rc = launch_phase<GLUE_ZILN>(0, grad, P, ncta, st); break;
    }
    if (rc) return rc;
    tc_finish_main_kernel<<<1, 128, 0, st>>>(p, ncta, nbf ? 1 : 0, grad ? 1 : 0);
    GLUE_CHECK_LAUNCH();
    if (grad) {
        if (nbf) {
            rc = a.family == GLUE_NB ? launch_phase<GLUE_NB>(1, true, P, ncta, st)
                                     : launch_phase<GLUE_ZINB>(1, true, P, ncta, st);
            if (rc) return rc;
        }
        tc_grad_u_kernel<<<a.B, 64, 0, st>>>(p, ncta);
        GLUE_CHECK_LAUNCH();
        if (a.reduce == 0) {
            dec_rescale_kernel<<<sm_count(), 256, 0, st>>>(p);
            GLUE_CHECK_LAUNCH();
        }
    }
    return 0;
}

}  // namespace glue
