// Small memory-bound stages of the encoder / decoder hidden layers:
// library size, dropout, BatchNorm1d over the minibatch (forward + backward), column sums,
// ReLU/dropout backward, Adam.  Reference: nn/networks.py:119-125 (Encoder.forward),
// module/spVIPESmodule.py:435 (library), scvi FCLayers BatchNorm1d(momentum=0.01, eps=0.001).
#include "common.cuh"
#include "../../include/spvipes_b200.h"

// ---------------------------------------------------------------------------------------
// library[b] = log(sum_g log1p(x[b, g]))      (reference :433-435, quirk Q2)
// ---------------------------------------------------------------------------------------
template <int SRC>
__global__ void library_kernel(const void* __restrict__ X, long ldx, const int* __restrict__ rows, int B, int G,
                               float* __restrict__ lib) {
    // one CTA (128 threads) per row
    __shared__ float red[4];
    const int b = blockIdx.x, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    long r = rows ? (long)rows[b] : (long)b;
    float s = 0.0f;
    for (int g = threadIdx.x; g < G; g += blockDim.x) s += load_src<SRC>(X, r * ldx + g);
    s = warp_sum(s);
    if (lane == 0) red[w] = s;
    __syncthreads();
    if (threadIdx.x == 0) lib[b] = logf(red[0] + red[1] + red[2] + red[3]);
}

extern "C" int spv_library_size(int src, const void* X, long long ldx, const int* rows, int B, int G, float* lib, void* stream) {
    if (!X || !lib || B <= 0 || G <= 0) return SPV_ERR_ARG;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (src == SPV_SRC_U16_LOG1P) library_kernel<SPV_SRC_U16_LOG1P><<<B, 128, 0, st>>>(X, ldx, rows, B, G, lib);
    else if (src == SPV_SRC_F32_LOG1P) library_kernel<SPV_SRC_F32_LOG1P><<<B, 128, 0, st>>>(X, ldx, rows, B, G, lib);
    else return SPV_ERR_ARG;
    SPV_CHECK_LAUNCH();
    return SPV_OK;
}

// ---------------------------------------------------------------------------------------
// dropout (in place): h *= mask.  mask != null: explicit multiplier (0 or 1/(1-p));
// else Philox keep-mask keyed by (seed, stream_id, *step, element index).
// ---------------------------------------------------------------------------------------
__global__ void dropout_kernel(float* __restrict__ h, long ld, int B, int C, const float* __restrict__ mask, long ldm,
                               float p, unsigned long long seed, unsigned int stream_id, const int* __restrict__ step) {
    long total = (long)B * C;
    unsigned int stp = step ? (unsigned int)*step : 0u;
    float inv_keep = 1.0f / (1.0f - p);
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        int c = (int)(i % C);
        long b = i / C;
        float m;
        if (mask) m = mask[b * ldm + c];
        else m = philox_uniform(seed, stream_id, stp, (unsigned long long)i) <= (1.0f - p) ? inv_keep : 0.0f;
        h[b * ld + c] *= m;
    }
}

extern "C" int spv_dropout(float* h, long long ld, int B, int C, const float* mask, long long ldm, float p,
                           unsigned long long seed, unsigned int stream_id, const int* step, void* stream) {
    if (!h || B <= 0 || C <= 0 || p < 0.0f || p >= 1.0f) return SPV_ERR_ARG;
    if (!mask && p == 0.0f) return SPV_OK;
    long total = (long)B * C;
    int blocks = (int)min((long)148 * 8, (total + 255) / 256);
    dropout_kernel<<<blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(h, ld, B, C, mask, ldm, p, seed, stream_id, step);
    SPV_CHECK_LAUNCH();
    return SPV_OK;
}

// dy <- dy * (y > 0 ? (mask ? mask : scale) : 0)      (ReLU [+ dropout] backward, y = saved output)
__global__ void relu_bwd_kernel(float* __restrict__ dy, long lddy, const float* __restrict__ y, long ldy, int B, int C,
                                const float* __restrict__ mask, long ldm, float scale) {
    long total = (long)B * C;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        int c = (int)(i % C);
        long b = i / C;
        float yv = y[b * ldy + c];
        float m = mask ? mask[b * ldm + c] : scale;
        float* d = dy + b * lddy + c;
        *d = yv > 0.0f ? (*d) * m : 0.0f;
    }
}

extern "C" int spv_relu_bwd(float* dy, long long lddy, const float* y, long long ldy, int B, int C, const float* mask,
                            long long ldm, float scale, void* stream) {
    if (!dy || !y || B <= 0 || C <= 0) return SPV_ERR_ARG;
    long total = (long)B * C;
    int blocks = (int)min((long)148 * 8, (total + 255) / 256);
    relu_bwd_kernel<<<blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(dy, lddy, y, ldy, B, C, mask, ldm, scale);
    SPV_CHECK_LAUNCH();
    return SPV_OK;
}

// ---------------------------------------------------------------------------------------
// BatchNorm1d over the minibatch.  8 columns (one 32-byte sector per row) x 64 row lanes per CTA: the column count is
// small (70 .. 256), so narrow CTAs are what spreads the rows of one column block over enough SMs.
// ---------------------------------------------------------------------------------------
#define BN_TX 8    // columns per CTA: one 32-byte sector per row
#define BN_TY 64   // row lanes

__device__ __forceinline__ float col_reduce(float v, float (*red)[BN_TX]) {
    const int tx = threadIdx.x % BN_TX, ty = threadIdx.x / BN_TX;
    __syncthreads();
    red[ty][tx] = v;
    __syncthreads();
    float s = 0.0f;
#pragma unroll
    for (int i = 0; i < BN_TY; ++i) s += red[i][tx];
    return s;
}

__global__ void __launch_bounds__(BN_TX* BN_TY) bn_fwd_kernel(const float* __restrict__ x, long ldx, float* __restrict__ y,
                                                               long ldy, int B, int C, const float* __restrict__ gamma,
                                                               const float* __restrict__ beta, float eps, float momentum,
                                                               float* __restrict__ running_mean, float* __restrict__ running_var,
                                                               float* __restrict__ save_mean, float* __restrict__ save_invstd,
                                                               int training, int relu) {
    __shared__ float red[BN_TY][BN_TX];
    const int tx = threadIdx.x % BN_TX, ty = threadIdx.x / BN_TX;
    const int c = blockIdx.x * BN_TX + tx;
    const bool ok = c < C;
    float mean, var;
    if (training) {
        float s = 0.0f;
        if (ok) for (int b = ty; b < B; b += BN_TY) s += x[(long)b * ldx + c];
        mean = col_reduce(s, red) / (float)B;
        float q = 0.0f;
        if (ok) for (int b = ty; b < B; b += BN_TY) { float d = x[(long)b * ldx + c] - mean; q += d * d; }
        var = col_reduce(q, red) / (float)B;  // biased, used for normalisation
        if (ok && ty == 0) {
            float unb = var * ((float)B / (float)max(B - 1, 1));
            running_mean[c] = (1.0f - momentum) * running_mean[c] + momentum * mean;
            running_var[c] = (1.0f - momentum) * running_var[c] + momentum * unb;
        }
    } else {
        mean = ok ? running_mean[c] : 0.0f;
        var = ok ? running_var[c] : 1.0f;
    }
    if (!ok) return;
    float invstd = 1.0f / sqrtf(var + eps);
    if (ty == 0) {
        if (save_mean) save_mean[c] = mean;
        if (save_invstd) save_invstd[c] = invstd;
    }
    float g = gamma[c], bt = beta[c];
    for (int b = ty; b < B; b += BN_TY) {
        float v = (x[(long)b * ldx + c] - mean) * invstd * g + bt;
        if (relu) v = fmaxf(v, 0.0f);
        y[(long)b * ldy + c] = v;
    }
}

extern "C" int spv_bn_fwd(const float* x, long long ldx, float* y, long long ldy, int B, int C, const float* gamma,
                          const float* beta, float eps, float momentum, float* running_mean, float* running_var,
                          float* save_mean, float* save_invstd, int training, int relu, void* stream) {
    if (!x || !y || !gamma || !beta || !running_mean || !running_var || B <= 0 || C <= 0) return SPV_ERR_ARG;
    bn_fwd_kernel<<<(C + BN_TX - 1) / BN_TX, BN_TX * BN_TY, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        x, ldx, y, ldy, B, C, gamma, beta, eps, momentum, running_mean, running_var, save_mean, save_invstd, training, relu);
    SPV_CHECK_LAUNCH();
    return SPV_OK;
}

// training-mode backward.  y_relu != null: the forward applied ReLU after the affine; dy is masked by y_relu > 0.
__global__ void __launch_bounds__(BN_TX* BN_TY) bn_bwd_kernel(const float* __restrict__ dy, long lddy, const float* __restrict__ x,
                                                               long ldx, const float* __restrict__ y_relu, long ldy,
                                                               float* __restrict__ dx, long lddx, int B, int C,
                                                               const float* __restrict__ gamma, const float* __restrict__ save_mean,
                                                               const float* __restrict__ save_invstd, float* __restrict__ dgamma,
                                                               float* __restrict__ dbeta) {
    __shared__ float red[BN_TY][BN_TX];
    const int tx = threadIdx.x % BN_TX, ty = threadIdx.x / BN_TX;
    const int c = blockIdx.x * BN_TX + tx;
    const bool ok = c < C;
    float mean = ok ? save_mean[c] : 0.0f, invstd = ok ? save_invstd[c] : 0.0f;
    float s1 = 0.0f, s2 = 0.0f;
    if (ok)
        for (int b = ty; b < B; b += BN_TY) {
            float d = dy[(long)b * lddy + c];
            if (y_relu && !(y_relu[(long)b * ldy + c] > 0.0f)) d = 0.0f;
            float xh = (x[(long)b * ldx + c] - mean) * invstd;
            s1 += d;
            s2 += d * xh;
        }
    s1 = col_reduce(s1, red);
    s2 = col_reduce(s2, red);
    if (!ok) return;
    if (ty == 0) {
        dgamma[c] = s2;
        dbeta[c] = s1;
    }
    float g = gamma[c] * invstd, m1 = s1 / (float)B, m2 = s2 / (float)B;
    for (int b = ty; b < B; b += BN_TY) {
        float d = dy[(long)b * lddy + c];
        if (y_relu && !(y_relu[(long)b * ldy + c] > 0.0f)) d = 0.0f;
        float xh = (x[(long)b * ldx + c] - mean) * invstd;
        dx[(long)b * lddx + c] = g * (d - m1 - xh * m2);
    }
}

extern "C" int spv_bn_bwd(const float* dy, long long lddy, const float* x, long long ldx, const float* y_relu, long long ldy,
                          float* dx, long long lddx, int B, int C, const float* gamma, const float* save_mean,
                          const float* save_invstd, float* dgamma, float* dbeta, void* stream) {
    if (!dy || !x || !dx || !gamma || !save_mean || !save_invstd || !dgamma || !dbeta || B <= 0 || C <= 0) return SPV_ERR_ARG;
    bn_bwd_kernel<<<(C + BN_TX - 1) / BN_TX, BN_TX * BN_TY, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        dy, lddy, x, ldx, y_relu, ldy, dx, lddx, B, C, gamma, save_mean, save_invstd, dgamma, dbeta);
    SPV_CHECK_LAUNCH();
    return SPV_OK;
}

__global__ void __launch_bounds__(BN_TX* BN_TY) colsum_kernel(const float* __restrict__ x, long ldx, int B, int C, float* __restrict__ out) {
    __shared__ float red[BN_TY][BN_TX];
    const int tx = threadIdx.x % BN_TX, ty = threadIdx.x / BN_TX;
    const int c = blockIdx.x * BN_TX + tx;
    float s = 0.0f;
    if (c < C) for (int b = ty; b < B; b += BN_TY) s += x[(long)b * ldx + c];
    s = col_reduce(s, red);
    if (c < C && ty == 0) out[c] = s;
}

extern "C" int spv_colsum(const float* x, long long ldx, int B, int C, float* out, void* stream) {
    if (!x || !out || B <= 0 || C <= 0) return SPV_ERR_ARG;
    colsum_kernel<<<(C + BN_TX - 1) / BN_TX, BN_TX * BN_TY, 0, reinterpret_cast<cudaStream_t>(stream)>>>(x, ldx, B, C, out);
    SPV_CHECK_LAUNCH();
    return SPV_OK;
}

// ---------------------------------------------------------------------------------------
// Adam, torch.optim.Adam semantics with L2 weight decay folded into the gradient
// (scvi TrainingPlan defaults: lr 1e-3, eps 0.01, weight_decay 1e-6).  *step is the 1-based
// step count held on the device so the launch can be replayed from a CUDA graph;
// spv_adam_tick increments it.
// ---------------------------------------------------------------------------------------
__global__ void adam_tick_kernel(int* step) { *step += 1; }

extern "C" int spv_adam_tick(int* step, void* stream) {
    if (!step) return SPV_ERR_ARG;
    adam_tick_kernel<<<1, 1, 0, reinterpret_cast<cudaStream_t>(stream)>>>(step);
    SPV_CHECK_LAUNCH();
    return SPV_OK;
}

__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                            long n, float lr, float b1, float b2, float eps, float wd, float grad_scale,
                            const int* __restrict__ step) {
    const int t = *step;
    const float bc1 = 1.0f - powf(b1, (float)t), bc2 = 1.0f - powf(b2, (float)t);
    const float step_size = lr / bc1, inv_sqrt_bc2 = 1.0f / sqrtf(bc2);
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        float pi = p[i];
        float gi = g[i] * grad_scale + wd * pi;
        float mi = b1 * m[i] + (1.0f - b1) * gi;
        float vi = b2 * v[i] + (1.0f - b2) * gi * gi;
        m[i] = mi;
        v[i] = vi;
        float denom = sqrtf(vi) * inv_sqrt_bc2 + eps;
        p[i] = pi - step_size * (mi / denom);
    }
}

extern "C" int spv_adam(float* p, const float* g, float* m, float* v, long long n, float lr, float b1, float b2, float eps,
                        float wd, float grad_scale, const int* step, void* stream) {
    if (!p || !g || !m || !v || !step || n <= 0) return SPV_ERR_ARG;
    int blocks = (int)min((long long)148 * 16, (n + 255) / 256);
    adam_kernel<<<blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(p, g, m, v, n, lr, b1, b2, eps, wd, grad_scale, step);
    SPV_CHECK_LAUNCH();
    return SPV_OK;
}
