// KL( N(mean, std) || N(0, 1) ) summed over a [n] block of a data encoder's posterior, and its
// gradient (sm_100a).
//
// The reference evaluates  D.kl_divergence(u, prior).sum(dim=1).mean() / n_features  per modality
// (scglue/models/scglue.py:349-353) -- with the standard-normal prior (scglue/models/sc.py:640-660)
//   kl = 0.5 (std^2 + mean^2 - 1) - log std
// per element.  Through torch.distributions that is ~10 elementwise launches forward and ~15
// backward per modality, all on the serial tail of the training step; here it is one launch
// each way.  [B, K] is a few thousand elements: one block, fixed reduction order.
#include "common.cuh"

namespace glue {
namespace {

constexpr int KL_THREADS = 512;

__global__ void __launch_bounds__(KL_THREADS) kl_unit_normal_fwd_kernel(const float* __restrict__ mean,
                                                                        const float* __restrict__ std_, int64_t n,
                                                                        float scale, float* __restrict__ out) {
    __shared__ float red[KL_THREADS / 32];
    float acc = 0.f;
    for (int64_t i = threadIdx.x; i < n; i += KL_THREADS) {
        const float m = mean[i], s = std_[i];
        acc += 0.5f * (fmaf(s, s, m * m) - 1.f) - logf(s);
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        float v = threadIdx.x < KL_THREADS / 32 ? red[threadIdx.x] : 0.f;
        v = warp_sum(v);
        if (threadIdx.x == 0) out[0] = v * scale;
    }
}

__global__ void __launch_bounds__(256) kl_unit_normal_bwd_kernel(const float* __restrict__ mean,
                                                                 const float* __restrict__ std_, int64_t n, float scale,
                                                                 const float* __restrict__ upstream,
                                                                 float* __restrict__ gmean, float* __restrict__ gstd) {
    const float c = upstream[0] * scale;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float s = std_[i];
        gmean[i] = c * mean[i];
        gstd[i] = c * (s - 1.f / s);
    }
}

// ---------------------------------------------------------------------------------------------
// Discriminator loss (scglue/models/scglue.py:326-329): F.cross_entropy(logits, target, reduction="none"),
// times the per-cell weight, summed, over the number of cells -- and its gradient in the same launch, with the
// coefficient the objective gives the term folded in (the generator's -lam_align, the discriminator update's
// +lam_align).  Through torch ops: five launches forward, five backward, on the serial stretch between the
// discriminator's forward and backward kernels at the end of the step.  One block, rows strided over the
// threads, fixed summation order.
// ---------------------------------------------------------------------------------------------
constexpr int CE_THREADS = 256;
constexpr int CE_MAX_CLASSES = 32;

__global__ void __launch_bounds__(CE_THREADS) weighted_ce_kernel(const float* __restrict__ logits,
                                                                 const int64_t* __restrict__ target,
                                                                 const float* __restrict__ weight, int64_t n, int C,
                                                                 float scale, float coef, float* __restrict__ out,
                                                                 float* __restrict__ grad) {
    __shared__ float red[CE_THREADS / 32];
    float acc = 0.f;
    for (int64_t i = threadIdx.x; i < n; i += CE_THREADS) {
        const float* row = logits + i * C;
        float x[CE_MAX_CLASSES];
        float mx = -INFINITY;
#pragma unroll
        for (int c = 0; c < CE_MAX_CLASSES; ++c)
            if (c < C) x[c] = row[c], mx = fmaxf(mx, x[c]);
        float sum = 0.f;
#pragma unroll
        for (int c = 0; c < CE_MAX_CLASSES; ++c)
            if (c < C) x[c] = expf(x[c] - mx), sum += x[c];
        const int t = (int)target[i];
        const float w = weight[i];
        acc += w * (mx + logf(sum) - row[t]);
        if (grad) {
            const float g = coef * scale * w, inv = 1.f / sum;
#pragma unroll
            for (int c = 0; c < CE_MAX_CLASSES; ++c)
                if (c < C) grad[i * C + c] = g * (x[c] * inv - (c == t ? 1.f : 0.f));
        }
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        float v = threadIdx.x < CE_THREADS / 32 ? red[threadIdx.x] : 0.f;
        v = warp_sum(v);
        if (threadIdx.x == 0) out[0] = v * scale;
    }
}

// ---------------------------------------------------------------------------------------------
// Paired cells (scglue/models/scglue.py:618-622, 654-660): mean embedding of every cell over the
// modalities it is observed in, and the cosine term
//   cos_loss = sum_k [n_k > 0] (1 - sum_i p_ki cos(u_ki, m_i) / max(n_k, 1)),   n_k = sum_i p_ki,
//   m_i = sum_k p_ki u_ki / sum_k p_ki,   cos(x, y) = x.y / (max(|x|, eps) max(|y|, eps)), eps = 1e-8
// (F.cosine_similarity).  Through torch ops that is ~30 launches forward and ~80 backward on the
// serial tail of the step.  One warp per cell; the sums over cells are taken by one block in a
// fixed order.
// ---------------------------------------------------------------------------------------------
constexpr int PM_MAX_MOD = 8;
constexpr int PM_ROWS = 8;  // cells (warps) per block
constexpr float PM_EPS = 1e-8f;

struct PmRow {
    float m[2];   // this lane's two latent columns of m_i
    float nm;     // |m_i|
    float cnt;    // sum_k p_ki
};

// lane handles latent columns lane and lane + 32 (K <= 64)
__device__ __forceinline__ PmRow pm_mean(const float* __restrict__ u, const float* __restrict__ pf, int M, int B, int K,
                                         int i, int lane) {
    PmRow r;
    r.m[0] = r.m[1] = 0.f, r.cnt = 0.f;
    for (int k = 0; k < M; ++k) {
        const float p = pf[(int64_t)k * B + i];
        const float* row = u + ((int64_t)k * B + i) * K;
        r.cnt += p;
        if (lane < K) r.m[0] = fmaf(p, row[lane], r.m[0]);
        if (lane + 32 < K) r.m[1] = fmaf(p, row[lane + 32], r.m[1]);
    }
    r.m[0] /= r.cnt, r.m[1] /= r.cnt;
    r.nm = sqrtf(warp_sum(fmaf(r.m[0], r.m[0], r.m[1] * r.m[1])));
    return r;
}

// rows: one warp per cell, PM_ROWS cells per block; writes umean and p_ki cos(u_ki, m_i)
__global__ void __launch_bounds__(PM_ROWS * 32) paired_mean_cos_rows_kernel(const float* __restrict__ u,
                                                                            const float* __restrict__ pf, int M, int B,
                                                                            int K, float* __restrict__ umean,
                                                                            float* __restrict__ psim) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = blockIdx.x * PM_ROWS + warp; i < B; i += gridDim.x * PM_ROWS) {
        const PmRow r = pm_mean(u, pf, M, B, K, i, lane);
        if (lane < K) umean[(int64_t)i * K + lane] = r.m[0];
        if (lane + 32 < K) umean[(int64_t)i * K + lane + 32] = r.m[1];
        for (int k = 0; k < M; ++k) {
            const float p = pf[(int64_t)k * B + i];
            const float* row = u + ((int64_t)k * B + i) * K;
            const float x0 = lane < K ? row[lane] : 0.f, x1 = lane + 32 < K ? row[lane + 32] : 0.f;
            const float dot = warp_sum(fmaf(x0, r.m[0], x1 * r.m[1]));
            const float nu = sqrtf(warp_sum(fmaf(x0, x0, x1 * x1)));
            const float sim = dot / (fmaxf(nu, PM_EPS) * fmaxf(r.nm, PM_EPS));
            if (lane == 0) psim[(int64_t)k * B + i] = p * sim;
        }
    }
}

// n_k, the masked sums of the cosines and the loss: one block, fixed order
__global__ void __launch_bounds__(256) paired_cos_reduce_kernel(const float* __restrict__ pf,
                                                                const float* __restrict__ psim, int M, int B,
                                                                float* __restrict__ cos_out, float* __restrict__ nk_out) {
    __shared__ float red_s[8], red_n[8];
    __shared__ float cosv;
    if (threadIdx.x == 0) cosv = 0.f;
    for (int k = 0; k < M; ++k) {
        float s = 0.f, n = 0.f;
        for (int i = threadIdx.x; i < B; i += 256) s += psim[(int64_t)k * B + i], n += pf[(int64_t)k * B + i];
        s = warp_sum(s), n = warp_sum(n);
        __syncthreads();
        if ((threadIdx.x & 31) == 0) red_s[threadIdx.x >> 5] = s, red_n[threadIdx.x >> 5] = n;
        __syncthreads();
        if (threadIdx.x == 0) {
            float ts = 0.f, tn = 0.f;
            for (int w = 0; w < 8; ++w) ts += red_s[w], tn += red_n[w];
            nk_out[k] = tn;
            if (tn > 0.f) cosv += 1.f - ts / fmaxf(tn, 1.f);
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) cos_out[0] = cosv;
}

__global__ void __launch_bounds__(PM_ROWS * 32) paired_mean_cos_bwd_kernel(const float* __restrict__ u,
                                                                         const float* __restrict__ pf,
                                                                         const float* __restrict__ nk, int M, int B,
                                                                         int K, const float* __restrict__ g_umean,
                                                                         const float* __restrict__ g_cos,
                                                                         float* __restrict__ g_u) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float gc = g_cos ? g_cos[0] : 0.f;
    for (int i = blockIdx.x * PM_ROWS + warp; i < B; i += gridDim.x * PM_ROWS) {
        const PmRow r = pm_mean(u, pf, M, B, K, i, lane);
        const float dm = fmaxf(r.nm, PM_EPS);
        // gradient reaching m_i: from the consumers of umean and from every cosine it takes part in
        float gm0 = (g_umean && lane < K) ? g_umean[(int64_t)i * K + lane] : 0.f;
        float gm1 = (g_umean && lane + 32 < K) ? g_umean[(int64_t)i * K + lane + 32] : 0.f;
        float du0[PM_MAX_MOD], du1[PM_MAX_MOD];  // direct d cos_loss / d u_ki
#pragma unroll
        for (int k = 0; k < PM_MAX_MOD; ++k) {
            du0[k] = du1[k] = 0.f;
            if (k < M) {
                const float p = pf[(int64_t)k * B + i];
                const float n = nk[k];
                const float a = (n > 0.f) ? -gc * p / fmaxf(n, 1.f) : 0.f;  // d loss / d sim_ki
                const float* row = u + ((int64_t)k * B + i) * K;
                const float x0 = lane < K ? row[lane] : 0.f, x1 = lane + 32 < K ? row[lane + 32] : 0.f;
                const float dot = warp_sum(fmaf(x0, r.m[0], x1 * r.m[1]));
                const float nu = sqrtf(warp_sum(fmaf(x0, x0, x1 * x1)));
                const float du = fmaxf(nu, PM_EPS);
                const float inv = 1.f / (du * dm);
                const float sim = dot * inv;
                // d sim / d x = m inv - sim x / |x|^2  (the second term only where the norm is not clamped)
                const float cu = nu > PM_EPS ? sim / (nu * nu) : 0.f;
                const float cm = r.nm > PM_EPS ? sim / (r.nm * r.nm) : 0.f;
                du0[k] = a * (r.m[0] * inv - cu * x0), du1[k] = a * (r.m[1] * inv - cu * x1);
                gm0 += a * (x0 * inv - cm * r.m[0]), gm1 += a * (x1 * inv - cm * r.m[1]);
            }
        }
#pragma unroll
        for (int k = 0; k < PM_MAX_MOD; ++k) {
            if (k < M) {
                const float w = pf[(int64_t)k * B + i] / r.cnt;  // d m_i / d u_ki
                float* out = g_u + ((int64_t)k * B + i) * K;
                if (lane < K) out[lane] = fmaf(w, gm0, du0[k]);
                if (lane + 32 < K) out[lane + 32] = fmaf(w, gm1, du1[k]);
            }
        }
    }
}

}  // namespace
}  // namespace glue

extern "C" int glue_kl_unit_normal_fwd(const float* mean, const float* std_, int64_t n, float scale, float* out,
                                       void* stream) {
    if (!mean || !std_ || !out || n <= 0) return GLUE_E_BADARG;
    glue::kl_unit_normal_fwd_kernel<<<1, glue::KL_THREADS, 0, (cudaStream_t)stream>>>(mean, std_, n, scale, out);
    GLUE_CHECK_LAUNCH();
    return 0;
}

extern "C" int glue_kl_unit_normal_bwd(const float* mean, const float* std_, int64_t n, float scale,
                                       const float* upstream, float* grad_mean, float* grad_std, void* stream) {
    if (!mean || !std_ || !upstream || !grad_mean || !grad_std || n <= 0) return GLUE_E_BADARG;
    const int64_t want = (n + 255) / 256;
    const int cap = 4 * glue::sm_count();
    const int grid = (int)(want > cap ? cap : want);
    glue::kl_unit_normal_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(mean, std_, n, scale, upstream, grad_mean,
                                                                            grad_std);
    GLUE_CHECK_LAUNCH();
    return 0;
}

extern "C" int glue_weighted_ce_fwd(const float* logits, const int64_t* target, const float* weight, int64_t n,
                                    int32_t n_classes, float scale, float coef, float* out_loss, float* grad_logits,
                                    void* stream) {
    if (!logits || !target || !weight || !out_loss || n <= 0) return GLUE_E_BADARG;
    if (n_classes < 1 || n_classes > glue::CE_MAX_CLASSES) return GLUE_E_UNSUPPORTED;
    glue::weighted_ce_kernel<<<1, glue::CE_THREADS, 0, (cudaStream_t)stream>>>(logits, target, weight, n, n_classes,
                                                                             scale, coef, out_loss, grad_logits);
    GLUE_CHECK_LAUNCH();
    return 0;
}

extern "C" int glue_paired_mean_cos_fwd(const float* ustack, const float* pmask, int32_t n_modalities, int64_t B,
                                        int32_t K, float* umean, float* cos_loss, float* n_k, float* psim,
                                        void* stream) {
    if (!ustack || !pmask || !umean || !cos_loss || !n_k || !psim || B <= 0) return GLUE_E_BADARG;
    if (n_modalities < 1 || n_modalities > glue::PM_MAX_MOD || K < 1 || K > 64) return GLUE_E_UNSUPPORTED;
    const int grid = (int)((B + glue::PM_ROWS - 1) / glue::PM_ROWS);
    glue::paired_mean_cos_rows_kernel<<<grid, glue::PM_ROWS * 32, 0, (cudaStream_t)stream>>>(
        ustack, pmask, n_modalities, (int)B, K, umean, psim);
    GLUE_CHECK_LAUNCH();
    glue::paired_cos_reduce_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(pmask, psim, n_modalities, (int)B, cos_loss,
                                                                      n_k);
    GLUE_CHECK_LAUNCH();
    return 0;
}

extern "C" int glue_paired_mean_cos_bwd(const float* ustack, const float* pmask, const float* n_k,
                                        int32_t n_modalities, int64_t B, int32_t K, const float* grad_umean,
                                        const float* grad_cos, float* grad_ustack, void* stream) {
    if (!ustack || !pmask || !n_k || !grad_ustack || B <= 0) return GLUE_E_BADARG;
    if (n_modalities < 1 || n_modalities > glue::PM_MAX_MOD || K < 1 || K > 64) return GLUE_E_UNSUPPORTED;
    const int grid = (int)((B + glue::PM_ROWS - 1) / glue::PM_ROWS);
    glue::paired_mean_cos_bwd_kernel<<<grid, glue::PM_ROWS * 32, 0, (cudaStream_t)stream>>>(
        ustack, pmask, n_k, n_modalities, (int)B, K, grad_umean, grad_cos, grad_ustack);
    GLUE_CHECK_LAUNCH();
    return 0;
}
