// Hidden block of the data encoders after its Linear layer: LeakyReLU -> BatchNorm1d (training
// mode) -> Dropout in one launch, and its backward in one launch (sm_100a).
//
// scglue/models/sc.py:62-178 builds every hidden block as
//   Linear -> LeakyReLU(0.2) -> BatchNorm1d -> Dropout(p).
// Through ATen the part after the GEMM is seven launches forward (activation, batch counter,
// batch statistics, running-statistics update, normalisation, dropout) and four backward on
// [B, H] = [128, 256] tensors -- launch latency, not work.  Two encoder passes per step and two
// blocks per encoder put ~80 us of that on the step's critical path.
//
// Semantics are torch's: a = leaky_relu(z); mean / biased variance over the batch;
// y = (a - mean) rstd gamma + beta; running_mean/var <- (1 - m) r + m (mean | var B / (B - 1));
// num_batches_tracked += 1; out = y keep / (1 - p), keep ~ Bernoulli(1 - p).  The keep mask comes
// from Philox4x32-10 keyed by a per-module seed and a per-call counter that lives on the device
// (so that a captured step draws fresh masks at every replay); it is returned as bytes for the
// backward pass.  Column statistics are two-pass (mean, then centred squares).
//
// Layout: block = 32 columns x all rows, 8 warps; lane <-> column (coalesced 128 B rows), warps
// stride over the rows; the 8 per-warp partials of a column are combined in a fixed order.
#include "common.cuh"
#include "philox.cuh"

namespace glue {
namespace {

constexpr int MLP_WARPS = 8;
constexpr int MLP_THREADS = MLP_WARPS * 32;

__device__ __forceinline__ float block_col_sum(float v, float (*red)[32], int warp, int lane) {
    __syncthreads();  // (red may still be read from the previous reduction)
    red[warp][lane] = v;
    __syncthreads();
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < MLP_WARPS; ++w) s += red[w][lane];
    return s;
}

struct BlockArgs {
    const float* z;             // [B, H] Linear output
    const float* gamma;         // [H]
    const float* beta;          // [H]
    float* running_mean;        // [H] (may be NULL)
    float* running_var;         // [H] (may be NULL)
    long long* num_batches;     // [1] (may be NULL)
    unsigned long long* rng;    // [3]: seed, call counter, ticket (NULL: no dropout)
    float* out;                 // [B, H]
    float* save_mean;           // [H]
    float* save_rstd;           // [H]
    unsigned char* keep;        // [B, H] (NULL when p == 0)
    int B, H;
    float slope, eps, momentum, p;
};

__global__ void __launch_bounds__(MLP_THREADS) act_bn_dropout_fwd_kernel(BlockArgs a) {
    __shared__ float red[MLP_WARPS][32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int col = blockIdx.x * 32 + lane;
    const bool cok = col < a.H;
    const float invB = 1.f / (float)a.B;
    // pass 1: column mean of the activation
    float s = 0.f;
    for (int r = warp; r < a.B; r += MLP_WARPS) {
        const float z = cok ? a.z[(int64_t)r * a.H + col] : 0.f;
        s += z > 0.f ? z : a.slope * z;
    }
    const float mean = block_col_sum(s, red, warp, lane) * invB;
    // pass 2: centred sum of squares
    float q = 0.f;
    for (int r = warp; r < a.B; r += MLP_WARPS) {
        const float z = cok ? a.z[(int64_t)r * a.H + col] : 0.f;
        const float d = (z > 0.f ? z : a.slope * z) - mean;
        q = fmaf(d, d, q);
    }
    const float var = block_col_sum(q, red, warp, lane) * invB;
    const float rstd = rsqrtf(var + a.eps);
    if (warp == 0 && cok) {
        a.save_mean[col] = mean, a.save_rstd[col] = rstd;
        if (a.running_mean) a.running_mean[col] = fmaf(a.momentum, mean - a.running_mean[col], a.running_mean[col]);
        if (a.running_var) {
            const float unbiased = a.B > 1 ? var * ((float)a.B / (float)(a.B - 1)) : var;
            a.running_var[col] = fmaf(a.momentum, unbiased - a.running_var[col], a.running_var[col]);
        }
    }
    // pass 3: normalise, scale / shift, drop
    const bool drop = a.p > 0.f && a.rng != nullptr;
    unsigned long long seed = 0, call = 0;
    if (drop) seed = a.rng[0], call = a.rng[1];
    const uint32_t threshold = drop ? (uint32_t)fminf(a.p * 4294967296.f, 4294967295.f) : 0u;
    const float keep_scale = drop ? 1.f / (1.f - a.p) : 1.f;
    const float g = cok ? a.gamma[col] * rstd : 0.f, b = cok ? a.beta[col] : 0.f;
    for (int r = warp; r < a.B; r += MLP_WARPS) {
        if (!cok) continue;
        const int64_t off = (int64_t)r * a.H + col;
        const float z = a.z[off];
        const float y = fmaf((z > 0.f ? z : a.slope * z) - mean, g, b);
        bool keep = true;
        if (drop) {
            keep = keep_flag(seed, call, r, col, threshold);
            a.keep[off] = keep ? 1 : 0;
        }
        a.out[off] = keep ? y * keep_scale : 0.f;
    }
    // the last block to finish advances the call counter (every block has read it by then)
    if (threadIdx.x == 0) {
        if (blockIdx.x == 0 && a.num_batches) a.num_batches[0] += 1;
        if (drop) {
            __threadfence();
            const unsigned long long t = atomicAdd(a.rng + 2, 1ull);
            if (t == gridDim.x - 1) a.rng[2] = 0ull, a.rng[1] = call + 1ull;
        }
    }
}

struct BlockBwdArgs {
    const float* z;
    const float* gamma;
    const float* save_mean;
    const float* save_rstd;
    const unsigned char* keep;  // NULL: no dropout
    const float* grad_out;      // [B, H]
    float* grad_z;              // [B, H]
    float* grad_gamma;          // [H]
    float* grad_beta;           // [H]
    int B, H;
    float slope, p;
};

__global__ void __launch_bounds__(MLP_THREADS) act_bn_dropout_bwd_kernel(BlockBwdArgs a) {
    __shared__ float red[MLP_WARPS][32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int col = blockIdx.x * 32 + lane;
    const bool cok = col < a.H;
    const float keep_scale = a.keep ? 1.f / (1.f - a.p) : 1.f;
    const float mean = cok ? a.save_mean[col] : 0.f, rstd = cok ? a.save_rstd[col] : 0.f;
    // pass 1: sum_i g, sum_i g xhat   (g = gradient reaching the BatchNorm output)
    float sg = 0.f, sgx = 0.f;
    for (int r = warp; r < a.B; r += MLP_WARPS) {
        if (!cok) continue;
        const int64_t off = (int64_t)r * a.H + col;
        float g = a.grad_out[off] * keep_scale;
        if (a.keep && !a.keep[off]) g = 0.f;
        const float z = a.z[off];
        const float xhat = ((z > 0.f ? z : a.slope * z) - mean) * rstd;
        sg += g, sgx = fmaf(g, xhat, sgx);
    }
    const float gb = block_col_sum(sg, red, warp, lane);
    const float gg = block_col_sum(sgx, red, warp, lane);
    if (warp == 0 && cok) {
        if (a.grad_gamma) a.grad_gamma[col] = gg;
        if (a.grad_beta) a.grad_beta[col] = gb;
    }
    // pass 2: d L / d z
    const float invB = 1.f / (float)a.B;
    const float c = cok ? a.gamma[col] * rstd : 0.f;
    for (int r = warp; r < a.B; r += MLP_WARPS) {
        if (!cok) continue;
        const int64_t off = (int64_t)r * a.H + col;
        float g = a.grad_out[off] * keep_scale;
        if (a.keep && !a.keep[off]) g = 0.f;
        const float z = a.z[off];
        const float xhat = ((z > 0.f ? z : a.slope * z) - mean) * rstd;
        const float ga = c * (g - invB * (gb + xhat * gg));
        a.grad_z[off] = z > 0.f ? ga : a.slope * ga;
    }
}

}  // namespace
}  // namespace glue

extern "C" int glue_act_bn_dropout_fwd(const float* z, const float* gamma, const float* beta, float* running_mean,
                                       float* running_var, int64_t* num_batches_tracked, uint64_t* rng_state,
                                       int64_t B, int32_t H, float negative_slope, float eps, float momentum,
                                       float p_drop, float* out, float* save_mean, float* save_rstd,
                                       uint8_t* keep_mask, void* stream) {
    if (!z || !gamma || !beta || !out || !save_mean || !save_rstd || B <= 0 || H <= 0) return GLUE_E_BADARG;
    if (p_drop < 0.f || p_drop >= 1.f) return GLUE_E_BADARG;
    if (p_drop > 0.f && (!rng_state || !keep_mask)) return GLUE_E_BADARG;
    if (B >= ((int64_t)1 << 31)) return GLUE_E_UNSUPPORTED;
    glue::BlockArgs a;
    a.z = z, a.gamma = gamma, a.beta = beta, a.running_mean = running_mean, a.running_var = running_var;
    a.num_batches = (long long*)num_batches_tracked, a.rng = (unsigned long long*)rng_state;
    a.out = out, a.save_mean = save_mean, a.save_rstd = save_rstd, a.keep = p_drop > 0.f ? keep_mask : nullptr;
    a.B = (int)B, a.H = H, a.slope = negative_slope, a.eps = eps, a.momentum = momentum, a.p = p_drop;
    glue::act_bn_dropout_fwd_kernel<<<(H + 31) / 32, glue::MLP_THREADS, 0, (cudaStream_t)stream>>>(a);
    GLUE_CHECK_LAUNCH();
    return 0;
}

extern "C" int glue_act_bn_dropout_bwd(const float* z, const float* gamma, const float* save_mean,
                                       const float* save_rstd, const uint8_t* keep_mask, const float* grad_out,
                                       int64_t B, int32_t H, float negative_slope, float p_drop, float* grad_z,
                                       float* grad_gamma, float* grad_beta, void* stream) {
    if (!z || !gamma || !save_mean || !save_rstd || !grad_out || !grad_z || B <= 0 || H <= 0) return GLUE_E_BADARG;
    if (B >= ((int64_t)1 << 31)) return GLUE_E_UNSUPPORTED;
    glue::BlockBwdArgs a;
    a.z = z, a.gamma = gamma, a.save_mean = save_mean, a.save_rstd = save_rstd;
    a.keep = p_drop > 0.f ? keep_mask : nullptr, a.grad_out = grad_out, a.grad_z = grad_z;
    a.grad_gamma = grad_gamma, a.grad_beta = grad_beta, a.B = (int)B, a.H = H, a.slope = negative_slope, a.p = p_drop;
    glue::act_bn_dropout_bwd_kernel<<<(H + 31) / 32, glue::MLP_THREADS, 0, (cudaStream_t)stream>>>(a);
    GLUE_CHECK_LAUNCH();
    return 0;
}
