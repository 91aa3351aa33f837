/*
 * glue_b200.h -- C ABI of the B200-native SCGLUE hot path (libglue_b200.so, sm_100a).
 *
 * This is the drop-in boundary.  The reference (gao-lab/GLUE, scglue 0.4.1) is
 * pure Python: its "operator interface" for this path is the set of
 * torch.nn.Module.forward methods listed below.  Each entry point replaces the
 * ATen op sequence behind one of them and is bound from Python with ctypes
 * (glue_b200/_lib.py); INTEGRATION.md shows the stub a reference maintainer
 * would add.
 *
 * Conventions (all entry points):
 *   - every pointer is a DEVICE pointer to memory owned by the caller; the
 *     library never allocates, frees, synchronises or keeps state between calls;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *   - tensors are dense row-major fp32 unless noted; indices are int64 where the
 *     reference uses int64 (eidx, b, vidx) and int32 inside the CSR;
 *   - return value: 0 on success, a positive cudaError_t, or a negative
 *     GLUE_E_* argument error.  glue_error_string() decodes both.
 *   - one process drives one GPU; calls on different streams are re-entrant.
 */
#ifndef GLUE_B200_H
#define GLUE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GLUE_B200_ABI_VERSION 6

#define GLUE_E_BADARG (-1)      /* NULL pointer / non-positive size               */
#define GLUE_E_UNSUPPORTED (-2) /* latent dim > 64, unknown family, ...           */
#define GLUE_E_WORKSPACE (-3)   /* workspace too small                            */
#define GLUE_E_ALIGN (-4)       /* pointer not aligned as documented              */

int glue_abi_version(void);
/* "sm_100a" -- the only architecture the library is compiled for. */
const char *glue_arch(void);
const char *glue_error_string(int code);
/* Diagnostic: kernels launched by this library in this process so far (the one piece of
 * process-wide state; it never influences results). */
unsigned long long glue_kernel_launches(void);

/* ------------------------------------------------------------------------- *
 * GraphConv  --  replaces scglue/models/nn.py:26-57 (GraphConv.forward) and,
 * called with the transposed CSR, its autograd backward.
 *
 *   out[r, :] = sum_{e in [indptr[r], indptr[r+1])} val[e] * in[col[e], :]
 *
 * CSR rows are edge targets (forward) or edge sources (backward), edges keep
 * their original relative order inside a row, val = esgn * enorm.  One warp per
 * row, vectorised row gathers, no atomics => bit-reproducible.
 * `in`/`out`: [n_rows, K] with 8-byte aligned rows when K is even.
 * ------------------------------------------------------------------------- */
int glue_graphconv_csr(const int32_t *indptr, const int32_t *col, const float *val,
                       const float *in, float *out, int64_t n_rows, int32_t K, void *stream);

/* Device CSR build (stable counting sort by `key`); bit-exact with
 * argsort(key, kind="stable") / bincount / cumsum.
 *   key, other : [E] int64 (row key, column payload), val: [E]
 *   indptr: [n_rows+1] int32, col: [E] int32, out_val: [E]
 * `workspace` must hold glue_csr_build_workspace_bytes(E, n_rows) bytes. */
size_t glue_csr_build_workspace_bytes(int64_t E, int64_t n_rows);
int glue_csr_build(const int64_t *key, const int64_t *other, const float *val, int64_t E,
                   int64_t n_rows, int32_t *indptr, int32_t *col, float *out_val,
                   void *workspace, size_t workspace_bytes, void *stream);

/* ------------------------------------------------------------------------- *
 * GraphEncoder head -- replaces, for the whole vertex table, the two Linear layers +
 * softplus of scglue/models/sc.py:44-46, the reparameterised sample
 * (scglue/models/scglue.py:329) and the KL term against N(0, 1)
 * (scglue/models/scglue.py:339-340), forward and backward:
 *
 *   mu = ptr W_loc^T + b_loc,  sigma = softplus(ptr W_std^T + b_std) + 1e-7
 *   vsamp = mu + noise * sigma               (noise == NULL: vsamp = mu)
 *   kl_sum[0] = sum_ik ( -log sigma + (sigma^2 + mu^2) / 2 - 1/2 )
 *
 * ptr / noise / mu / sigma / vsamp: [V, K]; W_*: [K, K] ([out, in] as torch.nn.Linear);
 * K even, <= 64.  Backward takes d L / d vsamp (NULL = zeros) and the DEVICE scalar
 * kl_coef[0] = d L / d kl_sum (NULL = 0) and writes every gradient; weight gradients
 * are reduced in a fixed order (bit-reproducible).
 * workspace: glue_graphenc_workspace_bytes(V, K), shared by forward and backward.
 * ------------------------------------------------------------------------- */
size_t glue_graphenc_workspace_bytes(int64_t V, int32_t K);
int glue_graphenc_head_fwd(const float *ptr, const float *w_loc, const float *b_loc,
                           const float *w_std, const float *b_std, const float *noise, int64_t V,
                           int32_t K, float *mu, float *sigma, float *vsamp, float *kl_sum,
                           void *workspace, size_t workspace_bytes, void *stream);
int glue_graphenc_head_bwd(const float *ptr, const float *w_loc, const float *w_std,
                           const float *noise, const float *mu, const float *sigma,
                           const float *grad_vsamp, const float *kl_coef, int64_t V, int32_t K,
                           float *grad_ptr, float *grad_w_loc, float *grad_b_loc,
                           float *grad_w_std, float *grad_b_std, void *workspace,
                           size_t workspace_bytes, void *stream);
/* Same, with the kernels pinned (A/B runs, tests of both paths): impl 0 = the library picks (the
 * K x K products per 64-vertex tile on warp-level tensor cores, three TF32 products per fp32
 * product), 1 = FFMA kernels, 2 = tensor-core kernels. */
int glue_graphenc_head_fwd_impl(const float *ptr, const float *w_loc, const float *b_loc,
                                const float *w_std, const float *b_std, const float *noise, int64_t V,
                                int32_t K, float *mu, float *sigma, float *vsamp, float *kl_sum,
                                void *workspace, size_t workspace_bytes, int32_t impl, void *stream);
int glue_graphenc_head_bwd_impl(const float *ptr, const float *w_loc, const float *w_std,
                                const float *noise, const float *mu, const float *sigma,
                                const float *grad_vsamp, const float *kl_coef, int64_t V, int32_t K,
                                float *grad_ptr, float *grad_w_loc, float *grad_b_loc,
                                float *grad_w_std, float *grad_b_std, void *workspace,
                                size_t workspace_bytes, int32_t impl, void *stream);

/* ------------------------------------------------------------------------- *
 * GraphDecoder + Bernoulli log-prob + pos/neg balanced mean -- replaces
 * scglue/models/sc.py:54-59 and scglue/models/scglue.py:331-338.
 *
 *   logit[e] = esgn[e] * <v[eidx[0,e]], v[eidx[1,e]]>
 *   nll[e]   = softplus(logit[e]) - ewt[e] * logit[e]
 *   out_loss[0] = ( sum_{ewt==0} nll / max(n_neg,1) + sum_{ewt!=0} nll / max(n_pos,1) )
 *                 / ((n_pos>0) + (n_neg>0))                (no host sync)
 *
 * eidx: [2, E] int64 (row 0 sources, row 1 targets).  Optional outputs (NULL to
 * skip): out_logits [E], out_nll [E], out_coef [E] = d out_loss / d logit[e].
 * workspace: glue_graphdec_workspace_bytes(E, V).
 * ------------------------------------------------------------------------- */
size_t glue_graphdec_workspace_bytes(int64_t E, int64_t V);
int glue_graphdec_bce_fwd(const float *v, const int64_t *eidx, const float *esgn,
                          const float *ewt, int64_t E, int64_t V, int32_t K, float *out_loss,
                          float *out_logits, float *out_nll, float *out_coef, void *workspace,
                          size_t workspace_bytes, void *stream);
/* Backward of the above w.r.t. v by segmented reduction over the edge
 * incidence list sorted by vertex (deterministic, no float atomics):
 *   grad_v[w, :] = sum_{e: s_e = w} coef[e] esgn[e] v[t_e] + sum_{e: t_e = w} coef[e] esgn[e] v[s_e]
 * `coef` is any per-edge upstream d L / d logit[e]; every row of grad_v [V, K] is
 * written (zeros for untouched vertices). */
int glue_graphdec_bwd(const float *v, const int64_t *eidx, const float *esgn, const float *coef,
                      int64_t E, int64_t V, int32_t K, float *grad_v, void *workspace,
                      size_t workspace_bytes, void *stream);
/* Same, with every coefficient multiplied by the device scalar upstream[0]
 * (autograd's d L / d out_loss) so that no extra pass over grad_v is needed. */
int glue_graphdec_bwd_scaled(const float *v, const int64_t *eidx, const float *esgn,
                             const float *coef, const float *upstream, int64_t E, int64_t V,
                             int32_t K, float *grad_v, void *workspace, size_t workspace_bytes,
                             void *stream);
/* The two halves of glue_graphdec_bwd_scaled.  The vertex-sorted incidence list depends on
 * `eidx` only: glue_graphdec_sort_incidence can run right after the forward kernel (same
 * workspace, kept alive), leaving memset + segmented reduction for the backward pass. */
int glue_graphdec_sort_incidence(const int64_t *eidx, int64_t E, int64_t V, void *workspace,
                                 size_t workspace_bytes, void *stream);
int glue_graphdec_bwd_sorted(const float *v, const int64_t *eidx, const float *esgn,
                             const float *coef, const float *upstream, int64_t E, int64_t V,
                             int32_t K, float *grad_v, void *workspace, size_t workspace_bytes,
                             void *stream);

/* The pair the training step calls (scglue/models/scglue.py:331-338 and its autograd backward).
 * glue_graph_nll_fwd: ONE kernel (+ a 4-byte memset): the last CTA to arrive reduces the per-block
 * sums in a fixed order, writes out_loss and leaves the two class weights in the workspace;
 * out_coef [E] (may be NULL) = sigma(logit) - ewt, NOT yet class-weighted.  prepare_bwd != 0 also
 * builds the vertex-sorted incidence list in the workspace (glue_graphdec_sort_incidence), which
 * must then stay alive until glue_graph_nll_bwd has run.
 * glue_graph_nll_bwd: grad_v [V, K] = d (upstream[0] * out_loss) / d v by segmented reduction;
 * class weights are applied per entry (by ewt[e] != 0), untouched vertices get zero rows from the
 * kernel itself (no memset), sums run in list order (bit-reproducible). */
int glue_graph_nll_fwd(const float *v, const int64_t *eidx, const float *esgn, const float *ewt,
                       int64_t E, int64_t V, int32_t K, float *out_loss, float *out_coef,
                       int32_t prepare_bwd, void *workspace, size_t workspace_bytes, void *stream);
int glue_graph_nll_bwd(const float *v, const int64_t *eidx, const float *esgn, const float *ewt,
                       const float *coef, const float *upstream, int64_t E, int64_t V, int32_t K,
                       float *grad_v, void *workspace, size_t workspace_bytes, void *stream);

/* ------------------------------------------------------------------------- *
 * Fused data decoders -- replace forward + log_prob + (nan)mean + backward of
 *   NB     scglue/models/sc.py:372-397 + torch NegativeBinomial.log_prob
 *   ZINB   scglue/models/sc.py:447-478 + scglue/models/prob.py:144-153
 *   Normal scglue/models/sc.py:280-308 + torch Normal.log_prob
 *   ZILN   scglue/models/sc.py:340-369 + scglue/models/prob.py:110-118
 * and the reduction at scglue/models/scglue.py:342-347.
 *
 *   a[i,j]  = softplus(scale_lin[b_i,j]) * <u_i, v_j> + bias[b_i,j]
 *   NB/ZINB: mu = softmax_j(a) * l_i ; NB(theta=exp(log_theta[b_i,j]), logits=log(mu+1e-7)-log_theta)
 *   Normal/ZILN: loc = a ; std = softplus(std_lin[b_i,j]) + 1e-7
 * The [B,F] logit matrix lives only in registers / shared memory.
 * ------------------------------------------------------------------------- */
enum glue_decoder_family { GLUE_NB = 0, GLUE_ZINB = 1, GLUE_NORMAL = 2, GLUE_ZILN = 3 };

typedef struct glue_decoder_args {
    int32_t family;    /* enum glue_decoder_family                                   */
    int32_t B;         /* rows (cells) in this call                                  */
    int32_t F;         /* features                                                   */
    int32_t K;         /* latent dimension, 1..64                                    */
    int32_t n_batches; /* rows of the per-feature parameter tables                   */
    int32_t reduce;    /* 0: mean over non-NaN x (nanmean)  1: plain mean            */
    int32_t impl;      /* 0: the library picks the kernels.  Diagnostics (A/B runs, tests of
                          every path): low byte 1 = CUDA-core kernels, 2 = tcgen05 kernels
                          with per-row operand staging, 3 = tcgen05 kernels with TMA-staged
                          operands; bit 8 = without the streamlined NB pass B; bit 9 = CTA 0
                          of the TMA-staged kernels writes a clock64 timeline into the
                          first 6 KB of the (1024-byte aligned) workspace.  A pinned path
                          that cannot take the call falls through to the next one.        */
    float grad_scale;  /* gradients are those of grad_scale * loss (the coefficient with
                          which the caller adds `loss` to its objective), `loss` itself is
                          reported unscaled; 0 is read as 1                             */
    /* inputs */
    const float *u;           /* [B, K]                                              */
    const float *v;           /* [n_vrows, K] feature-latent table                   */
    const int64_t *vidx;      /* [F] row of `v` holding feature j; NULL = identity   */
    const int64_t *b;         /* [B] parameter row per cell; NULL = all zero         */
    const int32_t *row_order; /* [B] cells sorted by b (stable); NULL = identity.
                                 Required when b has more than one distinct value.   */
    const float *l;           /* [B] library size (NB/ZINB); NULL for Normal/ZILN    */
    const float *x;           /* [B, F] observations                                 */
    const float *scale_lin;   /* [n_batches, F]                                      */
    const float *bias;        /* [n_batches, F]                                      */
    const float *disp;        /* [n_batches, F] log_theta (NB/ZINB) or std_lin       */
    const float *zi_logits;   /* [n_batches, F] (ZINB/ZILN) else NULL                */
    const float *wmat;        /* optional [B, F] upstream d L / d log_prob; when
                                 non-NULL the objective is L = sum wmat*log_prob and
                                 `reduce`/`loss` are ignored                         */
    const float *row_weight;  /* optional [B] per-cell weights w_i >= 0: the objective
                                 becomes loss = -sum_i w_i sum_j log_prob_ij (masked /
                                 weighted row subsets with static shapes, e.g.
                                 w_i = mask_i / (n_masked * F) for the paired-cell losses,
                                 scglue/models/scglue.py:603-640); NaN observations are
                                 not skipped in this mode                            */
    /* outputs (any may be NULL when its gradient is not wanted) */
    float *loss;           /* [1]  -mean(log_prob)                                   */
    float *grad_u;         /* [B, K]                                                 */
    float *grad_v;         /* [n_vrows, K]; only rows vidx[0..F) are written         */
    float *grad_scale_lin; /* [n_batches, F] (all rows written)                      */
    float *grad_bias;
    float *grad_disp;
    float *grad_zi_logits;
    void *workspace; /* glue_decoder_workspace_bytes()                               */
    size_t workspace_bytes;
} glue_decoder_args;

size_t glue_decoder_workspace_bytes(int32_t family, int32_t B, int32_t F, int32_t K,
                                    int32_t n_batches);
/* loss and all gradients of loss in one call (training step). */
int glue_decoder_nll_fwd_bwd(const glue_decoder_args *args, void *stream);
/* loss only (validation / get_losses). */
int glue_decoder_nll_fwd(const glue_decoder_args *args, void *stream);
/* Materialising variants for API users of Distribution.log_prob / .mean
 * (scglue/models/scglue.py:1355 decode_data): out is [B, F]. */
int glue_decoder_log_prob(const glue_decoder_args *args, float *out, void *stream);
int glue_decoder_mean(const glue_decoder_args *args, float *out, void *stream);

/* ------------------------------------------------------------------------- *
 * Several loss terms of one modality in one launch sequence.  PairedSCGLUETrainer
 * (scglue/models/scglue.py:603-652) evaluates up to three decoder terms per modality on the
 * same observations: reconstruction from the cell's own embedding, from the mean embedding
 * of its paired modalities (joint cross) and from the other modality's embedding (real
 * cross).  They share x, v, b, l and the parameter tables and differ in u, the per-cell
 * weights and the coefficient in the objective.  Fields of `shared` that are per term
 * (u, row_weight, grad_scale, loss, grad_u) and wmat are ignored; grad_v and the
 * parameter-table gradients receive the SUM over the terms of grad_scale[t] * d loss[t]
 * (what autograd would accumulate), loss[t] is reported unscaled.  With reduce == 0
 * (nanmean) and n_terms > 1 every term needs a row_weight.  B <= 128 cells.
 * ------------------------------------------------------------------------- */
#define GLUE_DECODER_MAX_TERMS 3
typedef struct glue_decoder_multi_args {
    glue_decoder_args shared;
    int32_t n_terms; /* 1 .. GLUE_DECODER_MAX_TERMS */
    const float *u[GLUE_DECODER_MAX_TERMS];          /* [B, K] each                    */
    const float *row_weight[GLUE_DECODER_MAX_TERMS]; /* [B] each or NULL (uniform mean) */
    float grad_scale[GLUE_DECODER_MAX_TERMS];        /* 0 is read as 1                  */
    float *loss[GLUE_DECODER_MAX_TERMS];             /* [1] each, may be NULL           */
    float *grad_u[GLUE_DECODER_MAX_TERMS];           /* [B, K] each, may be NULL        */
} glue_decoder_multi_args;
size_t glue_decoder_multi_workspace_bytes(int32_t family, int32_t B, int32_t F, int32_t K,
                                          int32_t n_batches, int32_t n_terms);
/* Returns GLUE_E_UNSUPPORTED when the shape is outside what the fused path takes (the
 * caller then issues one glue_decoder_nll_fwd_bwd per term). */
int glue_decoder_nll_multi_fwd_bwd(const glue_decoder_multi_args *args, void *stream);
int glue_decoder_nll_multi_fwd(const glue_decoder_multi_args *args, void *stream);

/* ------------------------------------------------------------------------- *
 * Fused RMSprop over a flat range -- replaces the multi-tensor launch sequence of
 * torch.optim.RMSprop(momentum=0, centered=False, weight_decay=0), the optimizer
 * scglue compiles its trainers with (scglue/models/scglue.py:889-942):
 *   square_avg = alpha * square_avg + (1 - alpha) * grad^2
 *   param     -= lr * grad / (sqrt(square_avg) + eps)
 * param / grad / square_avg: [n] fp32 (16-byte aligned ranges take the vector path).
 * ------------------------------------------------------------------------- */
int glue_rmsprop_flat(float *param, const float *grad, float *square_avg, int64_t n, float lr,
                      float alpha, float eps, void *stream);

/* n <= 16 device-to-device copies (dst[s] <- src[s], bytes[s] bytes) in one launch: the minibatch of a
 * captured training step going into the step's static input buffers (the reference's
 * SCGLUETrainer.format_data, scglue/models/scglue.py:255-289, moves each tensor separately).  The three
 * arrays are host arrays. */
int glue_multi_copy(int32_t n, const void *const *src, void *const *dst, const int64_t *bytes,
                    void *stream);

/* ------------------------------------------------------------------------- *
 * Host side of the data path: one graph epoch of the guidance-graph negative sampler.
 * Replaces  GraphDataset.propose_shuffle  (scglue/models/data.py:794-822): n_pos = round(sum(ewt)) positive
 * edges drawn with replacement (p = ewt / sum), neg_samples corrupted copies of each whose target is redrawn
 * from the vertex distribution until (src, dst, sign) is not a real edge, weights 1 / 0, then one permutation --
 * draw for draw on the legacy MT19937 stream of numpy.random.RandomState(seed), so the edge stream is the
 * reference's bit for bit.  All pointers are HOST pointers; no device work, no stream.
 *   esrc / edst [n_edges] int64, esgn [n_edges] fp32   the graph's edge triplets
 *   ecdf [n_edges], vcdf [n_vertices] fp64             cumulative distributions, normalised so the last entry is 1
 *   elut [2^elut_bits + 1], vlut [...] int32           lut[b] = number of cdf entries <= b / 2^bits
 *   edge_keys [n_keys] int64, sorted, unique           (src * n_vertices + dst) * 2 + (sign > 0) of every real edge
 *   key_filter [2^filter_bits / 8] u8                  bit (slot & 7) of byte (slot >> 3) set for slot =
 *                                                      (key * 0x9E3779B97F4A7C15 mod 2^64) >> (64 - filter_bits)
 *                                                      of every real edge (a filter without false negatives)
 *   out_src / out_dst [n] int64, out_wt / out_sgn [n] fp32,  n = n_pos * (1 + neg_samples) < 2^32
 * Thread-safe (no shared state): the look-ahead threads of the data side call it concurrently.
 * ------------------------------------------------------------------------- */
int glue_sample_graph_epoch(uint32_t seed, const int64_t *esrc, const int64_t *edst, const float *esgn,
                            int64_t n_edges, const double *ecdf, const int32_t *elut, int32_t elut_bits,
                            const double *vcdf, const int32_t *vlut, int32_t vlut_bits, int64_t n_vertices,
                            const int64_t *edge_keys, int64_t n_keys, const uint8_t *key_filter,
                            int32_t filter_bits, int64_t n_pos, int32_t neg_samples, int64_t *out_src,
                            int64_t *out_dst, float *out_wt, float *out_sgn);

/* ------------------------------------------------------------------------- *
 * KL( N(mean, std) || N(0, 1) ) of a data encoder's posterior.
 * Replaces  D.kl_divergence(u, prior).sum(dim=1).mean() / n_features
 * (scglue/models/scglue.py:349-353) for the standard-normal prior of
 * scglue/models/sc.py:640-660:
 *   out[0] = scale * sum_i ( 0.5 (std_i^2 + mean_i^2 - 1) - log std_i )
 *   grad_mean = upstream[0] * scale * mean,  grad_std = upstream[0] * scale * (std - 1 / std)
 * mean / std / grads: [n] fp32; out, upstream: device scalars.  One block forward
 * (fixed summation order).
 * ------------------------------------------------------------------------- */
int glue_kl_unit_normal_fwd(const float *mean, const float *std_, int64_t n, float scale, float *out,
                            void *stream);
int glue_kl_unit_normal_bwd(const float *mean, const float *std_, int64_t n, float scale,
                            const float *upstream, float *grad_mean, float *grad_std, void *stream);

/* ------------------------------------------------------------------------- *
 * Discriminator loss and its gradient in one launch.
 * Replaces  (F.cross_entropy(logits, target, reduction="none") * weight).sum() / weight.numel()
 * (scglue/models/scglue.py:326-329) and its autograd backward:
 *   out_loss[0]       = scale * sum_i weight_i * (logsumexp(logits_i) - logits[i, target_i])
 *   grad_logits[i, c] = coef * scale * weight_i * (softmax(logits_i)_c - [c == target_i])   (NULL: not wanted)
 * coef is the coefficient with which the caller's objective takes the term (+lam_align in the discriminator
 * update, -lam_align in the generator's loss): folded here, autograd hands grad_logits on as it is.
 * logits / grad_logits [n, n_classes] fp32 row-major (n_classes <= 32, else GLUE_E_UNSUPPORTED),
 * target [n] int64 in [0, n_classes), weight [n] fp32.  One block, fixed summation order.
 * ------------------------------------------------------------------------- */
int glue_weighted_ce_fwd(const float *logits, const int64_t *target, const float *weight, int64_t n,
                         int32_t n_classes, float scale, float coef, float *out_loss, float *grad_logits,
                         void *stream);

/* ------------------------------------------------------------------------- *
 * Paired cells: mean embedding over the modalities a cell is observed in and the
 * cosine term of PairedSCGLUETrainer (scglue/models/scglue.py:618-622, 654-660):
 *   umean_i  = sum_k p_ki u_ki / sum_k p_ki
 *   cos_loss = sum_k [n_k > 0] (1 - sum_i p_ki cos(u_ki, umean_i) / max(n_k, 1)),  n_k = sum_i p_ki
 * with cos as F.cosine_similarity (eps = 1e-8).  ustack: [M, B, K], pmask: [M, B]
 * (0 / 1 as fp32), umean: [B, K], cos_loss: [1], n_k: [M], psim: [M, B] scratch;
 * M <= 8, K <= 64.  bwd: grad_ustack [M, B, K] from grad_umean [B, K] (may be NULL)
 * and the device scalar grad_cos (may be NULL).  One warp per cell; sums over cells
 * in a fixed order.
 * ------------------------------------------------------------------------- */
int glue_paired_mean_cos_fwd(const float *ustack, const float *pmask, int32_t n_modalities, int64_t B,
                             int32_t K, float *umean, float *cos_loss, float *n_k, float *psim,
                             void *stream);
int glue_paired_mean_cos_bwd(const float *ustack, const float *pmask, const float *n_k,
                             int32_t n_modalities, int64_t B, int32_t K, const float *grad_umean,
                             const float *grad_cos, float *grad_ustack, void *stream);

/* ------------------------------------------------------------------------- *
 * Hidden block of the data encoders after its Linear layer
 * (scglue/models/sc.py:62-178: Linear -> LeakyReLU(0.2) -> BatchNorm1d -> Dropout):
 * activation, training-mode batch normalisation (batch statistics, running
 * statistics and batch counter updated as torch.nn.BatchNorm1d does) and dropout
 * in one launch; one launch for the backward pass.
 *   z [B, H]: Linear output; gamma / beta / running_mean / running_var [H];
 *   rng_state [3] (uint64: seed, call counter, ticket -- advanced on the device, so a
 *   captured step draws a fresh mask at every replay; may be NULL when p_drop == 0);
 *   out [B, H]; save_mean / save_rstd [H] and keep_mask [B, H] (bytes) feed the backward.
 * bwd: grad_z [B, H], grad_gamma / grad_beta [H] from grad_out [B, H].
 * ------------------------------------------------------------------------- */
int glue_act_bn_dropout_fwd(const float *z, const float *gamma, const float *beta, float *running_mean,
                            float *running_var, int64_t *num_batches_tracked, uint64_t *rng_state,
                            int64_t B, int32_t H, float negative_slope, float eps, float momentum,
                            float p_drop, float *out, float *save_mean, float *save_rstd,
                            uint8_t *keep_mask, void *stream);
int glue_act_bn_dropout_bwd(const float *z, const float *gamma, const float *save_mean,
                            const float *save_rstd, const uint8_t *keep_mask, const float *grad_out,
                            int64_t B, int32_t H, float negative_slope, float p_drop, float *grad_z,
                            float *grad_gamma, float *grad_beta, void *stream);

/* ------------------------------------------------------------------------- *
 * Whole data-encoder MLP (scglue/models/sc.py:62-178: h_depth x [Linear -> LeakyReLU(0.2) ->
 * BatchNorm1d -> Dropout], then the `loc` / `std_lin` heads; forward at sc.py:136-178) -- one launch
 * per hidden block and one for both heads, forward and backward.  B <= 128 rows, every operand width
 * (Din, H, k_above) a multiple of 4 and <= 256, 16-byte aligned tensors; other shapes return
 * GLUE_E_UNSUPPORTED and the caller uses the per-op path (Linear + glue_act_bn_dropout_*).
 *
 * glue_mlp_hidden_fwd: out = dropout(bn(leaky_relu(in W^T + bias))); gamma == NULL skips the
 *   BatchNorm (discriminator blocks, sc.py:576-624).  z [B, H] (the Linear output), save_mean /
 *   save_rstd [H] and keep_mask [B, H] feed the backward pass and may be NULL when no gradient is
 *   needed.  BatchNorm state and rng_state exactly as glue_act_bn_dropout_fwd (same masks).
 * glue_mlp_heads_fwd: loc = in Wloc^T + bloc, std = softplus(in Wstd^T + bstd) + 1e-7;
 *   spre [B, O] = the pre-activation of std (for the backward; may be NULL).
 * glue_mlp_heads_bwd: from d_loc, d_std [B, O]: G [B, 2 O] = [d_loc | d_std sigmoid(spre)], and
 *   the gradients of both heads' weights [O, H] and biases [O].
 * glue_mlp_hidden_bwd (top block first): g_above [B, k_above] is the gradient reaching the output
 *   of the layer above this block (G of the heads, or dz of the block above) and
 *   w_above0 / w_above1 that layer's weights ([n0, H] and [k_above - n0, H]; w_above1 NULL when
 *   n0 == k_above): the kernel forms g_above W_above restricted to its columns, runs dropout /
 *   BatchNorm / LeakyReLU backward and writes dz [B, H] (NULL: not needed), dW [H, Din] = dz^T in,
 *   db, dgamma, dbeta [H].  All reductions over the batch run in a fixed order.
 * ------------------------------------------------------------------------- */
int glue_mlp_hidden_fwd(const float *in, int64_t ld_in, const float *W, const float *bias,
                        const float *gamma, const float *beta, float *running_mean, float *running_var,
                        int64_t *num_batches_tracked, uint64_t *rng_state, int64_t B, int32_t Din,
                        int32_t H, float negative_slope, float eps, float momentum, float p_drop, float *z,
                        float *out, float *save_mean, float *save_rstd, uint8_t *keep_mask, void *stream);
int glue_mlp_heads_fwd(const float *in, int64_t ld_in, const float *Wloc, const float *bloc,
                       const float *Wstd, const float *bstd, int64_t B, int32_t H, int32_t O, float *loc,
                       float *std, float *spre, void *stream);
int glue_mlp_heads_bwd(const float *h, int64_t ld_h, const float *d_loc, const float *d_std,
                       const float *spre, int64_t B, int32_t H, int32_t O, float *G, float *dWloc,
                       float *dbloc, float *dWstd, float *dbstd, void *stream);
int glue_mlp_hidden_bwd(const float *g_above, int32_t k_above, const float *w_above0,
                        const float *w_above1, int32_t n0, const float *z, const float *gamma,
                        const float *save_mean, const float *save_rstd, const uint8_t *keep_mask,
                        const float *in, int64_t ld_in, int32_t Din, int64_t B, int32_t H,
                        float negative_slope, float p_drop, float *dz, float *dW, float *db, float *dgamma,
                        float *dbeta, void *stream);

/* ------------------------------------------------------------------------- *
 * Diagnostic: run `nk` tcgen05.mma (kind::f16, bf16 operands, fp32 accumulate,
 * M = 128) steps on caller-built shared-memory operand images and return the
 * accumulator tile out[128][N].  a_img / b_img are device buffers copied
 * verbatim into 1024-byte aligned shared memory; a_desc_base / b_desc_base are
 * UMMA shared-memory descriptors without the start address; a_off / b_off are
 * HOST arrays of nk start-address byte offsets.  tests/test_tc_gpu.py uses it to
 * pin the operand layouts (K-major / MN-major, 128-byte swizzle) the fused
 * decoder relies on against a plain matrix product.
 * ------------------------------------------------------------------------- */
int glue_tc_selftest(const void *a_img, uint32_t a_bytes, const void *b_img, uint32_t b_bytes,
                     uint64_t a_desc_base, uint64_t b_desc_base, uint32_t idesc, uint32_t nk,
                     const uint32_t *a_off, const uint32_t *b_off, uint32_t N, float *out,
                     void *stream);

/* Diagnostic: device buffer of 3*8*16 int64 receiving clock64 stamps of CTA 0 of the tensor-core
 * decoder kernels (pipeline timeline, tools/tc_trace.py); NULL switches it off (default). */
void glue_tc_set_trace(void *device_buffer);

#ifdef __cplusplus
}
#endif
#endif /* GLUE_B200_H */
