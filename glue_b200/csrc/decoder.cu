// Fused data decoders (NB / ZINB / Normal / ZILN), forward + backward, sm_100a.
//
// Reference behaviour: scglue/models/sc.py:280-478 (decoder modules),
// scglue/models/prob.py:110-153 (ZILN / ZINB log_prob), torch NegativeBinomial /
// Normal / LogNormal log_prob, and the (nan)mean at scglue/models/scglue.py:342-347.
//
// Decomposition ("v1", CUDA-core FFMA): every CTA owns a contiguous range of
// features and walks ALL rows of the minibatch for it, so everything that is
// reduced over rows (d/dv, d/dscale_lin, d/dbias, d/ddisp, d/dzi) is CTA-local
// and written without atomics.  What is reduced over features (soft-max
// statistics, c_i = sum_j dL/dlogp_ij, d/du) goes through per-CTA partials in
// the workspace that are summed in a fixed order => bit-reproducible.
//
//   L = sum_ij W_ij * log_prob_ij      W_ij = -1/(B*F) (fused NLL) or wmat_ij
//
//   a_ij  = s_j <u_i, v_j> + bias_j                     s = softplus(scale_lin)
//   NB family:  logp_ij = a_ij - lse_i,  mu = exp(logp) l_i
//     G_ij  = dL/dlogp_ij,  c_i = sum_j G_ij,  dL/da_ij = G_ij - p_ij c_i
//   pass A  (rowstats) : lse_i                       reads v, u
//   pass B  (main)     : L, G-terms, c_i, A1, A2     reads x ONCE, v, u
//   pass C  (softmax)  : the  -p_ij c_i  terms       reads v, u (no x)
//   Normal family: only pass B (dL/da_ij = G_ij).
#include <stdlib.h>

#include "decoder_common.cuh"

namespace glue {

constexpr int TF = 128;  // features per tile == threads per CTA (thread <-> feature)
constexpr int RC = 32;   // rows per chunk (x tile, G/P tiles)
constexpr int RB = 128;  // rows whose u is resident in shared memory

template <int KP>
struct __align__(16) DecSmem {
    float u[RB][KP];
    float v[TF][KP];
    float G[RC][TF + 1];
    float P[RC][TF + 1];
    float x[2][RC][TF];
    float l[RB];
    float lse[RB];
    float c[RB];
    int b[RB];
    int row[RB];
    float cw[RC][4];
    float cw2[RC][4];
};


// ---- staging helpers -------------------------------------------------------
template <int KP>
__device__ __forceinline__ void stage_v_tile(DecSmem<KP>& S, const glue_decoder_args& a, int f0, int f_end) {
    for (int idx = threadIdx.x; idx < TF * KP; idx += TF) {
        const int r = idx / KP, k = idx - r * KP;
        const int jj = f0 + r;
        float val = 0.f;
        if (jj < f_end && k < a.K) {
            const int64_t vr = a.vidx ? a.vidx[jj] : (int64_t)jj;
            val = __ldg(a.v + vr * a.K + k);
        }
        S.v[r][k] = val;
    }
}

template <int KP>
__device__ __forceinline__ void stage_rows(DecSmem<KP>& S, const DecParams& p, int rb0, int nrows, bool want_lse,
                                           bool want_c) {
    const glue_decoder_args& a = p.a;
    for (int idx = threadIdx.x; idx < nrows * KP; idx += TF) {
        const int r = idx / KP, k = idx - r * KP;
        const int orig = a.row_order ? a.row_order[rb0 + r] : rb0 + r;
        S.u[r][k] = k < a.K ? __ldg(a.u + (int64_t)orig * a.K + k) : 0.f;
    }
    for (int r = threadIdx.x; r < nrows; r += TF) {
        const int orig = a.row_order ? a.row_order[rb0 + r] : rb0 + r;
        S.row[r] = orig;
        S.b[r] = a.b ? (int)a.b[orig] : 0;
        S.l[r] = a.l ? a.l[orig] : 1.f;
        S.lse[r] = want_lse ? p.ws.lse[rb0 + r] : 0.f;
        S.c[r] = want_c ? p.ws.c[rb0 + r] : 0.f;
    }
}

template <int KP>
__device__ __forceinline__ float dot_row(const float (&v)[KP], const float* __restrict__ urow) {
    float d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;
#pragma unroll
    for (int k4 = 0; k4 < KP / 4; ++k4) {
        const float4 uu = *reinterpret_cast<const float4*>(urow + 4 * k4);
        d0 = fmaf(uu.x, v[4 * k4 + 0], d0);
        d1 = fmaf(uu.y, v[4 * k4 + 1], d1);
        d2 = fmaf(uu.z, v[4 * k4 + 2], d2);
        d3 = fmaf(uu.w, v[4 * k4 + 3], d3);
    }
    return (d0 + d1) + (d2 + d3);
}

template <int KP>
__device__ __forceinline__ void axpy_row(float (&acc)[KP], float w, const float* __restrict__ urow) {
#pragma unroll
    for (int k4 = 0; k4 < KP / 4; ++k4) {
        const float4 uu = *reinterpret_cast<const float4*>(urow + 4 * k4);
        acc[4 * k4 + 0] = fmaf(w, uu.x, acc[4 * k4 + 0]);
        acc[4 * k4 + 1] = fmaf(w, uu.y, acc[4 * k4 + 1]);
        acc[4 * k4 + 2] = fmaf(w, uu.z, acc[4 * k4 + 2]);
        acc[4 * k4 + 3] = fmaf(w, uu.w, acc[4 * k4 + 3]);
    }
}

// ---- pass A: soft-max row statistics ---------------------------------------
template <int KP>
__global__ void __launch_bounds__(TF, 1) dec_rowstats_kernel(DecParams p) {
    extern __shared__ float4 smem_raw[];
    DecSmem<KP>& S = *reinterpret_cast<DecSmem<KP>*>(smem_raw);
    const glue_decoder_args& a = p.a;
    const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
    const int cta = blockIdx.x;
    const int f_begin = cta * p.feat_per_cta;
    const int f_end = min(a.F, f_begin + p.feat_per_cta);
    bool first = true;
    for (int f0 = f_begin; f0 < f_end; f0 += TF, first = false) {
        const int j = f0 + t;
        const bool jok = j < f_end;
        __syncthreads();
        stage_v_tile<KP>(S, a, f0, f_end);
        __syncthreads();
        float v[KP];
#pragma unroll
        for (int k = 0; k < KP; ++k) v[k] = S.v[t][k];
        int cur_b = -1;
        float s = 0.f, bias = 0.f;
        for (int rb0 = 0; rb0 < a.B; rb0 += RB) {
            const int nrows = min(RB, a.B - rb0);
            __syncthreads();
            stage_rows<KP>(S, p, rb0, nrows, false, false);
            __syncthreads();
            for (int c0 = 0; c0 < nrows; c0 += RC) {
                const int rows_here = min(RC, nrows - c0);
                for (int r = 0; r < rows_here; ++r) {
                    const int bi = S.b[c0 + r];
                    if (bi != cur_b) {
                        const int64_t off = (int64_t)bi * a.F + j;
                        s = jok ? softplusf(a.scale_lin[off]) : 0.f;
                        bias = jok ? a.bias[off] : 0.f;
                        cur_b = bi;
                    }
                    const float dot = dot_row<KP>(v, &S.u[c0 + r][0]);
                    S.G[r][t] = jok ? fmaf(s, dot, bias) : -INFINITY;
                }
                __syncthreads();
                if (lane < rows_here) {  // thread (row = lane, column block = wid)
                    float m = -INFINITY;
                    for (int q = 0; q < 32; ++q) m = fmaxf(m, S.G[lane][wid * 32 + q]);
                    float sum = 0.f;
                    if (m > -INFINITY)
                        for (int q = 0; q < 32; ++q) sum += expf(S.G[lane][wid * 32 + q] - m);
                    S.cw[lane][wid] = m;
                    S.cw2[lane][wid] = sum;
                }
                __syncthreads();
                if (t < rows_here) {
                    float m = fmaxf(fmaxf(S.cw[t][0], S.cw[t][1]), fmaxf(S.cw[t][2], S.cw[t][3]));
                    float sum = 0.f;
#pragma unroll
                    for (int w = 0; w < 4; ++w)
                        if (S.cw[t][w] > -INFINITY) sum += S.cw2[t][w] * expf(S.cw[t][w] - m);
                    const int64_t slot = (int64_t)cta * a.B + rb0 + c0 + t;
                    if (!first) {
                        const float pm = p.ws.pmax[slot], ps = p.ws.psum[slot];
                        const float nm = fmaxf(pm, m);
                        sum = (pm > -INFINITY ? ps * expf(pm - nm) : 0.f) +
                              (m > -INFINITY ? sum * expf(m - nm) : 0.f);
                        m = nm;
                    }
                    p.ws.pmax[slot] = m;
                    p.ws.psum[slot] = sum;
                }
                __syncthreads();
            }
        }
    }
}

// ---- pass B: main ----------------------------------------------------------
template <int KP, int FAM, int MODE>
__global__ void __launch_bounds__(TF, 1) dec_main_kernel(DecParams p) {
    extern __shared__ float4 smem_raw[];
    DecSmem<KP>& S = *reinterpret_cast<DecSmem<KP>*>(smem_raw);
    const glue_decoder_args& a = p.a;
    constexpr bool NBF = (FAM == GLUE_NB || FAM == GLUE_ZINB);
    constexpr bool GRAD = (MODE == MODE_GRAD);
    constexpr bool NEEDX = (MODE != MODE_MEAN);
    const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
    const int cta = blockIdx.x;
    const int f_begin = cta * p.feat_per_cta;
    const int f_end = min(a.F, f_begin + p.feat_per_cta);
    float loss_acc = 0.f;
    int bad_acc = 0;
    bool first = true;
    for (int f0 = f_begin; f0 < f_end; f0 += TF, first = false) {
        const int j = f0 + t;
        const bool jok = j < f_end;
        __syncthreads();
        stage_v_tile<KP>(S, a, f0, f_end);
        __syncthreads();
        float v[KP], gv[KP];
#pragma unroll
        for (int k = 0; k < KP; ++k) v[k] = S.v[t][k];
#pragma unroll
        for (int k = 0; k < KP; ++k) gv[k] = 0.f;
        int cur_b = -1;
        FeatParams fp = {};
        float gb = 0.f, gs = 0.f, gd = 0.f, gz = 0.f;
        auto flush = [&]() {
            if (GRAD && cur_b >= 0 && jok) {
                const int64_t off = (int64_t)cur_b * a.F + j;
                if (a.grad_bias) a.grad_bias[off] = gb;
                if (a.grad_scale_lin) a.grad_scale_lin[off] = gs * fp.sg;
                if (a.grad_disp) a.grad_disp[off] = gd;
                if (a.grad_zi_logits && (FAM == GLUE_ZINB || FAM == GLUE_ZILN)) a.grad_zi_logits[off] = gz;
            }
            gb = gs = gd = gz = 0.f;
        };
        for (int rb0 = 0; rb0 < a.B; rb0 += RB) {
            const int nrows = min(RB, a.B - rb0);
            __syncthreads();
            stage_rows<KP>(S, p, rb0, nrows, NBF, false);
            __syncthreads();
            const int nchunks = (nrows + RC - 1) / RC;
            auto issue_x = [&](int c) {
                if (NEEDX) {
                    const int rows_c = min(RC, nrows - c * RC);
                    for (int r = 0; r < rows_c; ++r)
                        cp_async4(&S.x[c & 1][r][t], a.x + (int64_t)S.row[c * RC + r] * a.F + (jok ? j : 0), jok);
                }
                cp_async_commit();
            };
            issue_x(0);
            for (int c = 0; c < nchunks; ++c) {
                if (c + 1 < nchunks) issue_x(c + 1);
                else cp_async_commit();
                cp_async_wait<1>();
                __syncthreads();
                const int rows_here = min(RC, nrows - c * RC);
                for (int r = 0; r < rows_here; ++r) {
                    const int rr = c * RC + r;
                    const int bi = S.b[rr];
                    if (bi != cur_b) {
                        flush();
                        fp = load_params<FAM>(a, bi, j, jok);
                        cur_b = bi;
                    }
                    const float dot = dot_row<KP>(v, &S.u[rr][0]);
                    const float av = fmaf(fp.s, dot, fp.bias);
                    const float xv = NEEDX ? S.x[c & 1][r][t] : 0.f;
                    const ElemOut e = element<FAM>(fp, av, xv, S.lse[rr], S.l[rr]);
                    if (MODE == MODE_LOGP || MODE == MODE_MEAN) {
                        if (jok) p.out[(int64_t)S.row[rr] * a.F + j] = MODE == MODE_LOGP ? e.lp : e.mean;
                        continue;
                    }
                    float W = a.wmat ? (jok ? a.wmat[(int64_t)S.row[rr] * a.F + j] : 0.f)
                                     : (a.row_weight ? -a.row_weight[S.row[rr]] * p.gscale : p.w0);
                    bool valid = jok;
                    if (a.reduce == 0 && !a.row_weight && xv != xv) {  // nanmean: NaN observations carry no weight
                        valid = false;
                        bad_acc += jok ? 1 : 0;
                    }
                    if (!valid) W = 0.f;
                    loss_acc += valid ? W * e.lp : 0.f;
                    if constexpr (GRAD) {
                        const float G = valid ? W * e.g : 0.f;
                        const float coef = G * fp.s;
                        axpy_row<KP>(gv, coef, &S.u[rr][0]);
                        gb += G;
                        gs += G * dot;
                        gd += valid ? W * e.gd : 0.f;
                        gz += valid ? W * e.gz : 0.f;
                        S.G[r][t] = coef;
                        if (NBF) {
                            S.P[r][t] = jok ? e.prob * fp.s : 0.f;
                            const float csum = warp_sum(G);
                            if (lane == 0) S.cw[r][wid] = csum;
                        }
                    }
                }
                if constexpr (GRAD) {
                    __syncthreads();
                    // phase 2: thread (row = lane, float4 chunks wid, wid+4, ...) contracts the
                    // G / P tiles with the v tile over the 128 features of this tile.
                    if (lane < rows_here) {
                        constexpr int NQ = (KP / 4 + 3) / 4;
                        float4 a1[NQ], a2[NQ];
#pragma unroll
                        for (int q = 0; q < NQ; ++q) a1[q] = a2[q] = make_float4(0.f, 0.f, 0.f, 0.f);
                        for (int jj = 0; jj < TF; ++jj) {
                            const float g = S.G[lane][jj];
                            const float pp = NBF ? S.P[lane][jj] : 0.f;
#pragma unroll
                            for (int q = 0; q < NQ; ++q) {
                                const int ch = wid + 4 * q;
                                if (ch < KP / 4) {
                                    const float4 vv = *reinterpret_cast<const float4*>(&S.v[jj][4 * ch]);
                                    a1[q].x = fmaf(g, vv.x, a1[q].x), a1[q].y = fmaf(g, vv.y, a1[q].y);
                                    a1[q].z = fmaf(g, vv.z, a1[q].z), a1[q].w = fmaf(g, vv.w, a1[q].w);
                                    if (NBF) {
                                        a2[q].x = fmaf(pp, vv.x, a2[q].x), a2[q].y = fmaf(pp, vv.y, a2[q].y);
                                        a2[q].z = fmaf(pp, vv.z, a2[q].z), a2[q].w = fmaf(pp, vv.w, a2[q].w);
                                    }
                                }
                            }
                        }
                        const int pos = rb0 + c * RC + lane;
                        float* ap = p.ws.apart + ((int64_t)cta * a.B + pos) * (2 * KP);
#pragma unroll
                        for (int q = 0; q < NQ; ++q) {
                            const int ch = wid + 4 * q;
                            if (ch < KP / 4) {
                                float4* d1 = reinterpret_cast<float4*>(ap + 4 * ch);
                                float4* d2 = reinterpret_cast<float4*>(ap + KP + 4 * ch);
                                if (!first) {
                                    const float4 o1 = *d1;
                                    a1[q].x += o1.x, a1[q].y += o1.y, a1[q].z += o1.z, a1[q].w += o1.w;
                                    if (NBF) {
                                        const float4 o2 = *d2;
                                        a2[q].x += o2.x, a2[q].y += o2.y, a2[q].z += o2.z, a2[q].w += o2.w;
                                    }
                                }
                                *d1 = a1[q];
                                if (NBF) *d2 = a2[q];
                            }
                        }
                        if (NBF && wid == 0) {
                            float cs = (S.cw[lane][0] + S.cw[lane][1]) + (S.cw[lane][2] + S.cw[lane][3]);
                            const int64_t slot = (int64_t)cta * a.B + pos;
                            if (!first) cs += p.ws.cpart[slot];
                            p.ws.cpart[slot] = cs;
                        }
                    }
                }
                __syncthreads();
            }
        }
        flush();
        if (GRAD) if (a.grad_v) {
            // transpose through shared memory so that d/dv rows leave with coalesced stores
            __syncthreads();
#pragma unroll
            for (int k = 0; k < KP; ++k) S.v[t][k] = gv[k];
            __syncthreads();
            for (int idx = t; idx < TF * a.K; idx += TF) {
                const int r = idx / a.K, k = idx - r * a.K;
                const int jj = f0 + r;
                if (jj < f_end) {
                    const int64_t vr = a.vidx ? a.vidx[jj] : (int64_t)jj;
                    a.grad_v[vr * a.K + k] = S.v[r][k];
                }
            }
        }
    }
    if (MODE == MODE_LOSS || MODE == MODE_GRAD) {
        // deterministic block reduction of the loss / skipped-NaN counters
        __shared__ float red_l[TF / 32];
        __shared__ int red_b[TF / 32];
        const float wl = warp_sum(loss_acc);
        int wb = bad_acc;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) wb += __shfl_xor_sync(FULL, wb, o);
        if (lane == 0) red_l[wid] = wl, red_b[wid] = wb;
        __syncthreads();
        if (t == 0) {
            p.ws.losspart[cta] = (red_l[0] + red_l[1]) + (red_l[2] + red_l[3]);
            p.ws.badpart[cta] = red_b[0] + red_b[1] + red_b[2] + red_b[3];
        }
    }
}

// ---- reduce: c_i, d/du, loss ------------------------------------------------
// grid = B blocks (sorted row position), KP..128 threads.
template <int KP>
__global__ void dec_reduce_kernel(DecParams p, int nbf, int grad) {
    const glue_decoder_args& a = p.a;
    const int pos = blockIdx.x;
    const int t = threadIdx.x;
    __shared__ float c_sh;
    if (grad) {
        if (t == 0) {
            float c = 0.f;
            if (nbf)
                for (int q = 0; q < p.ncta; ++q) c += p.ws.cpart[(int64_t)q * a.B + pos];
            c_sh = c;
            p.ws.c[pos] = c;
        }
        __syncthreads();
        if (t < a.K && a.grad_u) {
            float a1 = 0.f, a2 = 0.f;
            for (int q = 0; q < p.ncta; ++q) {
                const float* ap = p.ws.apart + ((int64_t)q * a.B + pos) * (2 * KP);
                a1 += ap[t];
                if (nbf) a2 += ap[KP + t];
            }
            const int orig = a.row_order ? a.row_order[pos] : pos;
            a.grad_u[(int64_t)orig * a.K + t] = a1 - c_sh * a2;
        }
    }
    if (pos == 0 && t == 0) {
        double loss = 0.0;
        long long bad = 0;
        for (int q = 0; q < p.ncta; ++q) {
            loss += (double)p.ws.losspart[q];
            bad += p.ws.badpart[q];
        }
        const double total = (double)a.B * (double)a.F;
        const double ratio = (bad > 0 && !a.wmat) ? total / (total - (double)bad) : 1.0;
        p.ws.rescale[0] = (float)ratio;
        if (a.loss) a.loss[0] = (float)(loss * ratio / (double)p.gscale);
    }
}

// ---- pass C: the -p_ij c_i terms of the soft-max backward --------------------
template <int KP>
__global__ void __launch_bounds__(TF, 1) dec_softmax_bwd_kernel(DecParams p) {
    extern __shared__ float4 smem_raw[];
    DecSmem<KP>& S = *reinterpret_cast<DecSmem<KP>*>(smem_raw);
    const glue_decoder_args& a = p.a;
    const int t = threadIdx.x;
    const int cta = blockIdx.x;
    const int f_begin = cta * p.feat_per_cta;
    const int f_end = min(a.F, f_begin + p.feat_per_cta);
    for (int f0 = f_begin; f0 < f_end; f0 += TF) {
        const int j = f0 + t;
        const bool jok = j < f_end;
        __syncthreads();
        stage_v_tile<KP>(S, a, f0, f_end);
        __syncthreads();
        float v[KP], tv[KP];
#pragma unroll
        for (int k = 0; k < KP; ++k) v[k] = S.v[t][k], tv[k] = 0.f;
        int cur_b = -1;
        float s = 0.f, sg = 0.f, bias = 0.f, tb = 0.f, ts = 0.f;
        auto flush = [&]() {
            if (cur_b >= 0 && jok) {
                const int64_t off = (int64_t)cur_b * a.F + j;
                if (a.grad_bias) a.grad_bias[off] -= tb;
                if (a.grad_scale_lin) a.grad_scale_lin[off] -= ts * sg;
            }
            tb = ts = 0.f;
        };
        for (int rb0 = 0; rb0 < a.B; rb0 += RB) {
            const int nrows = min(RB, a.B - rb0);
            __syncthreads();
            stage_rows<KP>(S, p, rb0, nrows, true, true);
            __syncthreads();
            for (int r = 0; r < nrows; ++r) {
                const int bi = S.b[r];
                if (bi != cur_b) {
                    flush();
                    const int64_t off = (int64_t)bi * a.F + j;
                    const float sl = jok ? a.scale_lin[off] : 0.f;
                    s = softplusf(sl), sg = sigmoidf(sl);
                    bias = jok ? a.bias[off] : 0.f;
                    cur_b = bi;
                }
                const float dot = dot_row<KP>(v, &S.u[r][0]);
                const float pr = jok ? expf(fmaf(s, dot, bias) - S.lse[r]) : 0.f;
                const float w = pr * S.c[r];
                axpy_row<KP>(tv, w * s, &S.u[r][0]);
                tb += w;
                ts += w * dot;
            }
        }
        flush();
        if (a.grad_v) {
            __syncthreads();
#pragma unroll
            for (int k = 0; k < KP; ++k) S.v[t][k] = tv[k];
            __syncthreads();
            for (int idx = t; idx < TF * a.K; idx += TF) {
                const int r = idx / a.K, k = idx - r * a.K;
                const int jj = f0 + r;
                if (jj < f_end) {
                    const int64_t vr = a.vidx ? a.vidx[jj] : (int64_t)jj;
                    a.grad_v[vr * a.K + k] -= S.v[r][k];
                }
            }
        }
    }
}

// ======================================================================= host
static int pick_kp(int K) {
    if (K <= 16) return 16;
    if (K <= 52) return 52;
    if (K <= 64) return 64;
    return 0;
}

static void plan(const glue_decoder_args& a, int& ncta, int& feat_per_cta) {
    const int sms = sm_count();
    int per = (a.F + sms - 1) / sms;
    per = (per + 31) / 32 * 32;
    feat_per_cta = per;
    ncta = (a.F + per - 1) / per;
}

static size_t ws_layout(const glue_decoder_args& a, int ncta, int KP, char* base, DecWs* ws) {
    size_t off = 0;
    auto take = [&](size_t bytes) {
        char* ptr = base ? base + off : nullptr;
        off += align_up(bytes, 256);
        return ptr;
    };
    const size_t nb = (size_t)ncta * a.B;
    float* pmax = (float*)take(nb * 4);
    float* psum = (float*)take(nb * 4);
    float* lse = (float*)take((size_t)a.B * 4);
    float* cpart = (float*)take(nb * 4);
    float* c = (float*)take((size_t)a.B * 4);
    float* apart = (float*)take(nb * 2 * KP * 4);
    float* losspart = (float*)take((size_t)ncta * 4);
    int* badpart = (int*)take((size_t)ncta * 4);
    float* rescale = (float*)take(4);
    if (ws) *ws = DecWs{pmax, psum, lse, cpart, c, apart, losspart, badpart, rescale, nullptr};
    return off;
}

static int validate(const glue_decoder_args* a, bool need_x) {
    if (!a) return GLUE_E_BADARG;
    if (a->B <= 0 || a->F <= 0 || a->K <= 0 || a->n_batches <= 0) return GLUE_E_BADARG;
    if (a->family < GLUE_NB || a->family > GLUE_ZILN) return GLUE_E_UNSUPPORTED;
    if (pick_kp(a->K) == 0) return GLUE_E_UNSUPPORTED;
    if (!a->u || !a->v || !a->scale_lin || !a->bias || !a->disp || !a->workspace) return GLUE_E_BADARG;
    if (need_x && !a->x) return GLUE_E_BADARG;
    if (is_nb_family(a->family) && !a->l) return GLUE_E_BADARG;
    if ((a->family == GLUE_ZINB || a->family == GLUE_ZILN) && !a->zi_logits) return GLUE_E_BADARG;
    return 0;
}

template <int KP, int FAM>
static int launch_main(int mode, const DecParams& p, int ncta, size_t smem, cudaStream_t st) {
#define GLUE_LAUNCH_MODE(M)                                                                          \
    case M: {                                                                                        \
        auto k = dec_main_kernel<KP, FAM, M>;                                                        \
        GLUE_SMEM_ONCE(k, smem); \
        k<<<ncta, TF, smem, st>>>(p);                                                                \
        break;                                                                                       \
    }
    switch (mode) {
        GLUE_LAUNCH_MODE(MODE_LOSS)
        GLUE_LAUNCH_MODE(MODE_GRAD)
        GLUE_LAUNCH_MODE(MODE_LOGP)
        GLUE_LAUNCH_MODE(MODE_MEAN)
        default:
            return GLUE_E_BADARG;
    }
#undef GLUE_LAUNCH_MODE
    GLUE_CHECK_LAUNCH();
    return 0;
}

template <int KP>
static int run(const glue_decoder_args* args, int mode, float* out, cudaStream_t st) {
    DecParams p;
    p.a = *args;
    const glue_decoder_args& a = p.a;
    plan(a, p.ncta, p.feat_per_cta);
    const size_t need = ws_layout(a, p.ncta, KP, nullptr, nullptr);
    if (a.workspace_bytes < need) return GLUE_E_WORKSPACE;
    ws_layout(a, p.ncta, KP, (char*)a.workspace, &p.ws);
    const float gscale = (a.grad_scale != 0.f && !a.wmat) ? a.grad_scale : 1.f;
    p.gscale = gscale;
    p.w0 = (float)(-1.0 / ((double)a.B * (double)a.F)) * gscale;
    p.out = out;
    const bool nbf = is_nb_family(a.family);
    const size_t smem = sizeof(DecSmem<KP>);

    if (mode == MODE_GRAD) {
        // rows of the parameter tables that no cell of this minibatch uses get zero gradient
        const size_t tab = (size_t)a.n_batches * a.F * sizeof(float);
        if (a.n_batches > 1) {
            if (a.grad_scale_lin) GLUE_CUDA(cudaMemsetAsync(a.grad_scale_lin, 0, tab, st));
            if (a.grad_bias) GLUE_CUDA(cudaMemsetAsync(a.grad_bias, 0, tab, st));
            if (a.grad_disp) GLUE_CUDA(cudaMemsetAsync(a.grad_disp, 0, tab, st));
            if (a.grad_zi_logits) GLUE_CUDA(cudaMemsetAsync(a.grad_zi_logits, 0, tab, st));
        }
    }
    if (nbf) {
        auto k = dec_rowstats_kernel<KP>;
        GLUE_SMEM_ONCE(k, smem);
        k<<<p.ncta, TF, smem, st>>>(p);
        GLUE_CHECK_LAUNCH();
        dec_lse_kernel<<<(a.B * 32 + 255) / 256, 256, 0, st>>>(p.ws, p.ncta, a.B);
        GLUE_CHECK_LAUNCH();
    }
    int rc = 0;
    switch (a.family) {
        case GLUE_NB: rc = launch_main<KP, GLUE_NB>(mode, p, p.ncta, smem, st); break;
        case GLUE_ZINB: rc = launch_main<KP, GLUE_ZINB>(mode, p, p.ncta, smem, st); break;
        case GLUE_NORMAL: rc = launch_main<KP, GLUE_NORMAL>(mode, p, p.ncta, smem, st); break;
        default: rc = launch_main<KP, GLUE_ZILN>(mode, p, p.ncta, smem, st); break;
    }
    if (rc) return rc;
    if (mode == MODE_LOSS || mode == MODE_GRAD) {
        dec_reduce_kernel<KP><<<a.B, 64, 0, st>>>(p, nbf ? 1 : 0, mode == MODE_GRAD ? 1 : 0);
        GLUE_CHECK_LAUNCH();
    }
    if (mode == MODE_GRAD) {
        if (nbf) {
            auto k = dec_softmax_bwd_kernel<KP>;
            GLUE_SMEM_ONCE(k, smem);
            k<<<p.ncta, TF, smem, st>>>(p);
            GLUE_CHECK_LAUNCH();
        }
        if (!a.wmat && !a.row_weight && a.reduce == 0) {
            dec_rescale_kernel<<<sm_count(), 256, 0, st>>>(p);
            GLUE_CHECK_LAUNCH();
        }
    }
    return 0;
}

// args->impl pins a decoder implementation (A/B comparisons, tests of every path).  Default: the
// TMA-staged tensor-core kernels where they apply, then the per-row-staged ones, then CUDA cores.
static int forced_version(const glue_decoder_args* a) { return a->impl & 0xff; }

static glue_decoder_multi_args as_multi(const glue_decoder_args* args) {
    glue_decoder_multi_args m = {};
    m.shared = *args;
    m.n_terms = 1;
    m.u[0] = args->u, m.row_weight[0] = args->row_weight, m.grad_scale[0] = args->grad_scale;
    m.loss[0] = args->loss, m.grad_u[0] = args->grad_u;
    return m;
}

static int dispatch(const glue_decoder_args* args, int mode, float* out, void* stream) {
    const int rc = validate(args, mode != MODE_MEAN);
    if (rc) return rc;
    if ((mode == MODE_LOGP || mode == MODE_MEAN) && !out) return GLUE_E_BADARG;
    cudaStream_t st = (cudaStream_t)stream;
    const int forced = forced_version(args);
    if (forced == 0 || forced == 3) {
        const glue_decoder_multi_args m = as_multi(args);
        if (tc3_eligible(m, mode) && args->workspace_bytes >= tc3_workspace_bytes(*args, 1)) return tc3_run(m, mode, st);
    }
    if (forced != 1 && tc_eligible(*args, mode)) {
        DecParams p;
        p.a = *args;
        p.out = out;
        return tc_run(p, mode, st);
    }
    switch (pick_kp(args->K)) {
        case 16: return run<16>(args, mode, out, st);
        case 52: return run<52>(args, mode, out, st);
        default: return run<64>(args, mode, out, st);
    }
}

static int dispatch_multi(const glue_decoder_multi_args* m, int mode, void* stream) {
    if (!m || m->n_terms < 1 || m->n_terms > GLUE_DECODER_MAX_TERMS) return GLUE_E_BADARG;
    glue_decoder_args probe = m->shared;  // the shared fields are validated with the first term's
    probe.u = m->u[0], probe.row_weight = m->row_weight[0], probe.loss = m->loss[0], probe.grad_u = m->grad_u[0];
    probe.wmat = nullptr;
    const int rc = validate(&probe, true);
    if (rc) return rc;
    for (int s = 0; s < m->n_terms; ++s)
        if (!m->u[s]) return GLUE_E_BADARG;
    const int forced = forced_version(&m->shared);
    if ((forced != 0 && forced != 3) || !tc3_eligible(*m, mode)) return GLUE_E_UNSUPPORTED;
    if (m->shared.workspace_bytes < tc3_workspace_bytes(m->shared, m->n_terms)) return GLUE_E_WORKSPACE;
    return tc3_run(*m, mode, (cudaStream_t)stream);
}

}  // namespace glue

extern "C" size_t glue_decoder_workspace_bytes(int32_t family, int32_t B, int32_t F, int32_t K,
                                               int32_t n_batches) {
    glue_decoder_args a = {};
    a.family = family, a.B = B, a.F = F, a.K = K, a.n_batches = n_batches;
    const int KP = glue::pick_kp(K);
    if (B <= 0 || F <= 0 || KP == 0) return 0;
    int ncta, per;
    glue::plan(a, ncta, per);
    size_t need = glue::ws_layout(a, ncta, KP, nullptr, nullptr);
    const size_t v2 = glue::tc_workspace_bytes(a);
    need = v2 > need ? v2 : need;
    if (B <= 128) {
        const size_t v3 = glue::tc3_workspace_bytes(a, 1);
        need = v3 > need ? v3 : need;
    }
    return need;
}

extern "C" size_t glue_decoder_multi_workspace_bytes(int32_t family, int32_t B, int32_t F, int32_t K,
                                                     int32_t n_batches, int32_t n_terms) {
    glue_decoder_args a = {};
    a.family = family, a.B = B, a.F = F, a.K = K, a.n_batches = n_batches;
    if (B <= 0 || B > 128 || F <= 0 || K <= 0 || K > 64 || n_terms < 1 || n_terms > GLUE_DECODER_MAX_TERMS) return 0;
    return glue::tc3_workspace_bytes(a, n_terms);
}

extern "C" int glue_decoder_nll_fwd_bwd(const glue_decoder_args* args, void* stream) {
    return glue::dispatch(args, glue::MODE_GRAD, nullptr, stream);
}
extern "C" int glue_decoder_nll_fwd(const glue_decoder_args* args, void* stream) {
    return glue::dispatch(args, glue::MODE_LOSS, nullptr, stream);
}
extern "C" int glue_decoder_nll_multi_fwd_bwd(const glue_decoder_multi_args* args, void* stream) {
    return glue::dispatch_multi(args, glue::MODE_GRAD, stream);
}
extern "C" int glue_decoder_nll_multi_fwd(const glue_decoder_multi_args* args, void* stream) {
    return glue::dispatch_multi(args, glue::MODE_LOSS, stream);
}
extern "C" int glue_decoder_log_prob(const glue_decoder_args* args, float* out, void* stream) {
    return glue::dispatch(args, glue::MODE_LOGP, out, stream);
}
extern "C" int glue_decoder_mean(const glue_decoder_args* args, float* out, void* stream) {
    return glue::dispatch(args, glue::MODE_MEAN, out, stream);
}
