// Pieces shared by the two fused-decoder implementations (decoder.cu: CUDA-core "v1",
// decoder_tc.cu: tcgen05 "v2"): workspace carve-up, per-feature parameters and the
// per-element likelihood / gradient math.
#pragma once
#include "common.cuh"

namespace glue {

enum Mode { MODE_LOSS = 0, MODE_GRAD = 1, MODE_LOGP = 2, MODE_MEAN = 3 };

struct DecWs {  // workspace carve-up (device pointers)
    float* pmax;     // [ncta][B]
    float* psum;     // [ncta][B]
    float* lse;      // [B]
    float* cpart;    // [ncta][B]
    float* c;        // [B]
    float* apart;    // [ncta][B][2][KP]
    float* losspart; // [ncta]
    int* badpart;    // [ncta]  number of NaN observations skipped
    float* rescale;  // [1]  (B*F)/count when NaNs were skipped, else 1
    unsigned int* counters;  // [4] arrival tickets of the "last CTA reduces" tails (tensor-core path)
};

struct DecParams {
    glue_decoder_args a;
    DecWs ws;
    int ncta;
    int feat_per_cta;
    float w0;      // uniform weight -grad_scale/(B*F) when wmat == NULL
    float gscale;  // grad_scale (1 when unset): folded into the weights, divided out of the loss
    float* out;  // MODE_LOGP / MODE_MEAN
};

__host__ __device__ inline bool is_nb_family(int fam) { return fam == GLUE_NB || fam == GLUE_ZINB; }

// Per-(batch row, feature) parameters, reloaded when the (sorted) batch index changes.
struct FeatParams {
    float s, sg;       // softplus(scale_lin), sigmoid(scale_lin)
    float bias;
    float disp;        // log_theta | std_lin
    float th;          // exp(log_theta)          (NB family)
    float std_inv, log_std, sg_std;  // Normal family
    float zi, spz, sgz, ezi;          // zero inflation
};

template <int FAM>
__device__ __forceinline__ FeatParams load_params(const glue_decoder_args& a, int bi, int j, bool jok) {
    FeatParams p;
    const int64_t off = (int64_t)bi * a.F + j;
    const float sl = jok ? a.scale_lin[off] : 0.f;
    p.s = softplusf(sl);
    p.sg = sigmoidf(sl);
    p.bias = jok ? a.bias[off] : 0.f;
    p.disp = jok ? a.disp[off] : 0.f;
    p.th = 1.f, p.std_inv = 1.f, p.log_std = 0.f, p.sg_std = 0.f;
    if (FAM == GLUE_NB || FAM == GLUE_ZINB) {
        p.th = expf(p.disp);
    } else {
        const float sd = softplusf(p.disp) + GLUE_EPS;
        p.std_inv = 1.f / sd;
        p.log_std = logf(sd);
        p.sg_std = sigmoidf(p.disp);
    }
    p.zi = 0.f, p.spz = 0.f, p.sgz = 0.f, p.ezi = 0.f;
    if (FAM == GLUE_ZINB || FAM == GLUE_ZILN) {
        p.zi = jok ? a.zi_logits[off] : 0.f;
        p.spz = softplusf(p.zi);
        p.sgz = sigmoidf(p.zi);
        p.ezi = expf(p.zi);
    }
    return p;
}

struct ElemOut {
    float lp;    // log_prob
    float g;     // d lp / d logp (NB family) or d lp / d loc (Normal family)
    float gd;    // d lp / d disp
    float gz;    // d lp / d zi_logits
    float prob;  // softmax probability (NB family)
    float mean;  // distribution mean
};

// FAST selects the hardware approximations (ex2 / lg2 / rcp based, ~2^-21 relative error) used
// by the tensor-core path; the CUDA-core path keeps the libdevice versions.
template <bool FAST>
__device__ __forceinline__ float exf(float x) { return FAST ? __expf(x) : expf(x); }
template <bool FAST>
__device__ __forceinline__ float lgf(float x) { return FAST ? __logf(x) : logf(x); }
template <bool FAST>
__device__ __forceinline__ float dvf(float a, float b) { return FAST ? __fdividef(a, b) : a / b; }

// rare path of nb_gamma_terms (non-integer or large counts), kept out of line
static __device__ __noinline__ void nb_gamma_terms_slow(float x, float theta, float* lg, float* dg) {
    *lg = lgammaf(theta + x) - lgammaf(theta) - lgammaf(1.f + x);
    *dg = digammaf(theta + x) - digammaf(theta);
}

template <bool FAST>
__device__ __forceinline__ void nb_gamma_terms_t(float x, float theta, float& lg, float& dg) {
    if (!FAST) {
        nb_gamma_terms(x, theta, lg, dg);
        return;
    }
    lg = 0.f, dg = 0.f;
    if (x == 0.f) return;
    if (x <= 12.f && x == floorf(x)) {
        float prod = 1.f, h = 0.f, quot = 1.f;
        const int n = (int)x;
#pragma unroll 1
        for (int m = 0; m < n; ++m) {
            const float t = theta + (float)m;
            prod *= t;
            quot *= (float)(m + 1);
            h += __fdividef(1.f, t);
        }
        lg = __logf(prod) - __logf(quot);
        dg = h;
        return;
    }
    nb_gamma_terms_slow(x, theta, &lg, &dg);
}

template <int FAM, bool FAST = false>
__device__ __forceinline__ ElemOut element(const FeatParams& p, float a, float x, float lse, float l) {
    ElemOut o;
    o.gz = 0.f;
    o.prob = 0.f;
    if (FAM == GLUE_NB || FAM == GLUE_ZINB) {
        const float logp = a - lse;
        const float pr = exf<FAST>(logp);
        const float mu = pr * l;
        const float m = mu + GLUE_EPS;
        const float Lm = lgf<FAST>(m), Lt = lgf<FAST>(m + p.th);
        float lg, dg;
        nb_gamma_terms_t<FAST>(x, p.th, lg, dg);
        const float lp = p.th * (p.disp - Lt) + x * (Lm - Lt) + lg;
        const float sig = dvf<FAST>(m, m + p.th);
        const float d = x - (x + p.th) * sig;
        const float tl = p.th * ((p.disp - Lt) + dg) - d;
        float wnb = 1.f;
        o.lp = lp;
        if (FAM == GLUE_ZINB) {
            if (fabsf(x) < GLUE_EPS) {
                const float e1 = exf<FAST>(lp);
                const float E = e1 + p.ezi + GLUE_EPS;
                o.lp = lgf<FAST>(E) - p.spz;
                wnb = dvf<FAST>(e1, E);
                o.gz = dvf<FAST>(p.ezi, E) - p.sgz;
            } else {
                o.lp = lp - p.spz;
                o.gz = -p.sgz;
            }
        }
        o.g = wnb * d * dvf<FAST>(mu, m);
        o.gd = wnb * tl;
        o.prob = pr;
        o.mean = m;
    } else {
        if (FAM == GLUE_ZILN && fabsf(x) < GLUE_EPS) {
            o.lp = p.zi - p.spz;
            o.g = 0.f;
            o.gd = 0.f;
            o.gz = 1.f - p.sgz;
        } else {
            const float y = FAM == GLUE_ZILN ? lgf<FAST>(x) : x;
            const float z = (y - a) * p.std_inv;
            o.lp = -0.5f * z * z - p.log_std - 0.91893853320467274178f;
            o.g = z * p.std_inv;
            o.gd = (z * z - 1.f) * p.std_inv * p.sg_std;
            if (FAM == GLUE_ZILN) {
                o.lp -= y + p.spz;
                o.gz = -p.sgz;
            }
        }
        if (FAM == GLUE_ZILN) {
            const float sd = 1.f / p.std_inv;
            o.mean = exf<FAST>(a + 0.5f * sd * sd);
        } else {
            o.mean = a;
        }
    }
    return o;
}


// v2 (tensor-core) path, decoder_tc.cu
bool tc_eligible(const glue_decoder_args& a, int mode);
size_t tc_workspace_bytes(const glue_decoder_args& a);
int tc_run(DecParams p, int mode, cudaStream_t st);

// v3 (TMA-staged, several terms per launch sequence), decoder_tc3.cu
bool tc3_eligible(const glue_decoder_multi_args& m, int mode);
size_t tc3_workspace_bytes(const glue_decoder_args& a, int n_terms);
int tc3_run(const glue_decoder_multi_args& m, int mode, cudaStream_t st);

// ---- small kernels shared by both paths (one copy per translation unit) ------------------------
// one warp per row: lse_i = log sum_part psum * exp(pmax - M) + M
static __global__ void dec_lse_kernel(DecWs ws, int nparts, int B) {
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= B) return;
    float m = -INFINITY;
    for (int c = lane; c < nparts; c += 32) m = fmaxf(m, ws.pmax[(int64_t)c * B + row]);
    m = warp_max(m);
    float s = 0.f;
    for (int c = lane; c < nparts; c += 32) {
        const float pm = ws.pmax[(int64_t)c * B + row];
        if (pm > -INFINITY) s += ws.psum[(int64_t)c * B + row] * expf(pm - m);
    }
    s = warp_sum(s);
    if (lane == 0) ws.lse[row] = m + logf(s);
}

// only does work when NaN observations were skipped (nanmean denominator)
static __global__ void dec_rescale_kernel(DecParams p) {
    const float ratio = p.ws.rescale[0];
    if (ratio == 1.f) return;
    const glue_decoder_args& a = p.a;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nth = (int64_t)gridDim.x * blockDim.x;
    if (a.grad_u)
        for (int64_t i = tid; i < (int64_t)a.B * a.K; i += nth) a.grad_u[i] *= ratio;
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

}  // namespace glue
