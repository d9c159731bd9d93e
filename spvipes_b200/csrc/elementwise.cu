// Small memory-bound stages of the encoder / decoder hidden layers:
// library size, dropout, BatchNorm1d over the minibatch (forward + backward), column sums,
// ReLU/dropout backward, Adam.  Reference: nn/networks.py:119-125 (Encoder.forward),
// module/spVIPESmodule.py:435 (library), scvi FCLayers BatchNorm1d(momentum=0.01, eps=0.001).
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cooperative_groups.h>
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

// uint16 counts: 16-byte loads (8 genes) when the row is 16-byte aligned, log1p of counts < 256 from a shared-memory table filled
// with the same log1pf (bit-identical to computing it in place), 256 threads per cell
__global__ void __launch_bounds__(256) library_u16_kernel(const unsigned short* __restrict__ X, long ldx, const int* __restrict__ rows, int B,
                                                          int G, float* __restrict__ lib) {
    __shared__ float lut[256];
    __shared__ float red[8];
    const int b = blockIdx.x, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    lut[threadIdx.x] = threadIdx.x == 0 ? 0.0f : log1pf((float)threadIdx.x);
    const unsigned short* row = X + (rows ? (long)rows[b] : (long)b) * ldx;
    __syncthreads();
    const int nvec = ((reinterpret_cast<uintptr_t>(row) & 15) == 0) ? G / 8 : 0;
    float s = 0.0f;
    for (int v = threadIdx.x; v < nvec; v += 256) {
        const uint4 raw = __ldg(reinterpret_cast<const uint4*>(row) + v);
        const unsigned int wd[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const unsigned int c0 = wd[j] & 0xffffu, c1 = wd[j] >> 16;
            s += c0 < 256u ? lut[c0] : log1pf((float)c0);
            s += c1 < 256u ? lut[c1] : log1pf((float)c1);
        }
    }
    for (int g = nvec * 8 + threadIdx.x; g < G; g += 256) {
        const unsigned int c = row[g];
        s += c < 256u ? lut[c] : log1pf((float)c);
    }
    s = warp_sum(s);
    if (lane == 0) red[w] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.0f;
#pragma unroll
        for (int i = 0; i < 8; ++i) t += red[i];
        lib[b] = logf(t);
    }
}

extern "C" int spv_library_size(int src, const void* X, long long ldx, const int* rows, int B, int G, float* lib, void* stream) {
    if (!X || !lib || B <= 0 || G <= 0) return SPV_ERR_ARG;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (src == SPV_SRC_U16_LOG1P) library_u16_kernel<<<B, 256, 0, st>>>(reinterpret_cast<const unsigned short*>(X), ldx, rows, B, G, lib);
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

// CACHED (B <= BN_TY * BN_R): a thread's rows of the column stay in registers, so the column is read from global memory
// once, with all loads in flight together, instead of once per pass (sum, centred squares, normalise): these kernels are
// bound by the chain of dependent memory round trips, not by bandwidth.
#define BN_R 16
template <bool CACHED>
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
    float xv[CACHED ? BN_R : 1];
    if (CACHED) {
#pragma unroll
        for (int r = 0; r < BN_R; ++r) {
            const int b = ty + r * BN_TY;
            xv[r] = (ok && b < B) ? x[(long)b * ldx + c] : 0.0f;
        }
    }
    const float g = ok ? gamma[c] : 0.0f, bt = ok ? beta[c] : 0.0f;
    const float rmean = ok ? running_mean[c] : 0.0f, rvar = ok ? running_var[c] : 1.0f;
    float mean = rmean, var = rvar;
    if (training) {
        float s = 0.0f;
        if (CACHED) {
#pragma unroll
            for (int r = 0; r < BN_R; ++r) s += xv[r];
        } else if (ok) {
            for (int b = ty; b < B; b += BN_TY) s += x[(long)b * ldx + c];
        }
        mean = col_reduce(s, red) / (float)B;
        float q = 0.0f;
        if (CACHED) {
#pragma unroll
            for (int r = 0; r < BN_R; ++r) {
                const float d = xv[r] - mean;
                if (ty + r * BN_TY < B) q += d * d;
            }
        } else if (ok) {
            for (int b = ty; b < B; b += BN_TY) { float d = x[(long)b * ldx + c] - mean; q += d * d; }
        }
        var = col_reduce(q, red) / (float)B;  // biased, used for normalisation
        if (ok && ty == 0) {
            float unb = var * ((float)B / (float)max(B - 1, 1));
            running_mean[c] = (1.0f - momentum) * rmean + momentum * mean;
            running_var[c] = (1.0f - momentum) * rvar + momentum * unb;
        }
    }
    if (!ok) return;
    float invstd = 1.0f / sqrtf(var + eps);
    if (ty == 0) {
        if (save_mean) save_mean[c] = mean;
        if (save_invstd) save_invstd[c] = invstd;
    }
    if (CACHED) {
#pragma unroll
        for (int r = 0; r < BN_R; ++r) {
            const int b = ty + r * BN_TY;
            float v = (xv[r] - mean) * invstd * g + bt;
            if (relu) v = fmaxf(v, 0.0f);
            if (b < B) y[(long)b * ldy + c] = v;
        }
    } else {
        for (int b = ty; b < B; b += BN_TY) {
            float v = (x[(long)b * ldx + c] - mean) * invstd * g + bt;
            if (relu) v = fmaxf(v, 0.0f);
            y[(long)b * ldy + c] = v;
        }
    }
}

extern "C" int spv_bn_fwd(const float* x, long long ldx, float* y, long long ldy, int B, int C, const float* gamma,
                          const float* beta, float eps, float momentum, float* running_mean, float* running_var,
                          float* save_mean, float* save_invstd, int training, int relu, void* stream) {
    if (!x || !y || !gamma || !beta || !running_mean || !running_var || B <= 0 || C <= 0) return SPV_ERR_ARG;
    const dim3 grid((C + BN_TX - 1) / BN_TX), block(BN_TX * BN_TY);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (B <= BN_TY * BN_R)
        bn_fwd_kernel<true><<<grid, block, 0, st>>>(x, ldx, y, ldy, B, C, gamma, beta, eps, momentum, running_mean, running_var,
                                                    save_mean, save_invstd, training, relu);
    else
        bn_fwd_kernel<false><<<grid, block, 0, st>>>(x, ldx, y, ldy, B, C, gamma, beta, eps, momentum, running_mean, running_var,
                                                     save_mean, save_invstd, training, relu);
    SPV_CHECK_LAUNCH();
    return SPV_OK;
}

// training-mode backward.  y_relu != null: the forward applied ReLU after the affine; dy is masked by y_relu > 0.
template <bool CACHED>
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
    float dv[CACHED ? BN_R : 1], xh[CACHED ? BN_R : 1];
    if (CACHED) {
        float yr[BN_R];
#pragma unroll
        for (int r = 0; r < BN_R; ++r) {
            const int b = ty + r * BN_TY;
            const bool in = ok && b < B;
            dv[r] = in ? dy[(long)b * lddy + c] : 0.0f;
            xh[r] = in ? x[(long)b * ldx + c] : 0.0f;
            yr[r] = (in && y_relu) ? y_relu[(long)b * ldy + c] : 1.0f;
        }
#pragma unroll
        for (int r = 0; r < BN_R; ++r)
            if (!(yr[r] > 0.0f)) dv[r] = 0.0f;
    }
    const float mean = ok ? save_mean[c] : 0.0f, invstd = ok ? save_invstd[c] : 0.0f;
    const float gam = ok ? gamma[c] : 0.0f;
    float s1 = 0.0f, s2 = 0.0f;
    if (CACHED) {
#pragma unroll
        for (int r = 0; r < BN_R; ++r) {
            xh[r] = (ty + r * BN_TY < B) ? (xh[r] - mean) * invstd : 0.0f;
            s1 += dv[r];
            s2 += dv[r] * xh[r];
        }
    } else if (ok) {
        for (int b = ty; b < B; b += BN_TY) {
            float d = dy[(long)b * lddy + c];
            if (y_relu && !(y_relu[(long)b * ldy + c] > 0.0f)) d = 0.0f;
            float xhat = (x[(long)b * ldx + c] - mean) * invstd;
            s1 += d;
            s2 += d * xhat;
        }
    }
    s1 = col_reduce(s1, red);
    s2 = col_reduce(s2, red);
    if (!ok) return;
    if (ty == 0) {
        dgamma[c] = s2;
        dbeta[c] = s1;
    }
    float g = gam * invstd, m1 = s1 / (float)B, m2 = s2 / (float)B;
    if (CACHED) {
#pragma unroll
        for (int r = 0; r < BN_R; ++r) {
            const int b = ty + r * BN_TY;
            if (b < B) dx[(long)b * lddx + c] = g * (dv[r] - m1 - xh[r] * m2);
        }
    } else {
        for (int b = ty; b < B; b += BN_TY) {
            float d = dy[(long)b * lddy + c];
            if (y_relu && !(y_relu[(long)b * ldy + c] > 0.0f)) d = 0.0f;
            float xhat = (x[(long)b * ldx + c] - mean) * invstd;
            dx[(long)b * lddx + c] = g * (d - m1 - xhat * m2);
        }
    }
}

// Larger minibatches (BN_TY * BN_R < B <= 8 x that): a thread-block CLUSTER of up to 8 CTAs along the rows, each caching its
// 1024 rows of the 8 columns in registers (one read of the inputs, every load in flight at once); the two column sums are
// exchanged through distributed shared memory and added in rank order (deterministic).  The uncached single-CTA form above
// walks the column three times in dependent round trips (45-55 us at B = 2048 on the critical path of the backward).
__global__ void __launch_bounds__(BN_TX* BN_TY) bn_bwd_cluster_kernel(const float* __restrict__ dy, long lddy, const float* __restrict__ x,
                                                                       long ldx, const float* __restrict__ y_relu, long ldy,
                                                                       float* __restrict__ dx, long lddx, int B, int C,
                                                                       const float* __restrict__ gamma, const float* __restrict__ save_mean,
                                                                       const float* __restrict__ save_invstd, float* __restrict__ dgamma,
                                                                       float* __restrict__ dbeta) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    __shared__ float red[BN_TY][BN_TX];
    __shared__ float part[2][BN_TX];
    const int tx = threadIdx.x % BN_TX, ty = threadIdx.x / BN_TX;
    const int c = blockIdx.x * BN_TX + tx;
    const bool ok = c < C;
    const int nrank = (int)cluster.num_blocks(), rank = (int)cluster.block_rank();
    const int r_base = rank * (BN_TY * BN_R);
    float dv[BN_R], xh[BN_R], yr[BN_R];
#pragma unroll
    for (int r = 0; r < BN_R; ++r) {
        const int b = r_base + ty + r * BN_TY;
        const bool in = ok && b < B;
        dv[r] = in ? dy[(long)b * lddy + c] : 0.0f;
        xh[r] = in ? x[(long)b * ldx + c] : 0.0f;
        yr[r] = (in && y_relu) ? y_relu[(long)b * ldy + c] : 1.0f;
    }
    const float mean = ok ? save_mean[c] : 0.0f, invstd = ok ? save_invstd[c] : 0.0f;
    const float gam = ok ? gamma[c] : 0.0f;
    float s1 = 0.0f, s2 = 0.0f;
#pragma unroll
    for (int r = 0; r < BN_R; ++r) {
        if (!(yr[r] > 0.0f)) dv[r] = 0.0f;
        xh[r] = (r_base + ty + r * BN_TY < B) ? (xh[r] - mean) * invstd : 0.0f;
        s1 += dv[r];
        s2 += dv[r] * xh[r];
    }
    s1 = col_reduce(s1, red);
    s2 = col_reduce(s2, red);
    if (ty == 0) {
        part[0][tx] = s1;
        part[1][tx] = s2;
    }
    cluster.sync();
    s1 = s2 = 0.0f;
    for (int k = 0; k < nrank; ++k) {  // rank order: every CTA of the cluster gets bitwise the same totals
        const float* rp = cluster.map_shared_rank(&part[0][0], k);
        s1 += rp[tx];
        s2 += rp[BN_TX + tx];
    }
    cluster.sync();  // nobody leaves (and frees its shared memory) while a peer may still be reading it
    if (!ok) return;
    if (ty == 0 && rank == 0) {
        dgamma[c] = s2;
        dbeta[c] = s1;
    }
    const float g = gam * invstd, m1 = s1 / (float)B, m2 = s2 / (float)B;
#pragma unroll
    for (int r = 0; r < BN_R; ++r) {
        const int b = r_base + ty + r * BN_TY;
        if (b < B) dx[(long)b * lddx + c] = g * (dv[r] - m1 - xh[r] * m2);
    }
}

extern "C" int spv_bn_bwd(const float* dy, long long lddy, const float* x, long long ldx, const float* y_relu, long long ldy,
                          float* dx, long long lddx, int B, int C, const float* gamma, const float* save_mean,
                          const float* save_invstd, float* dgamma, float* dbeta, void* stream) {
    if (!dy || !x || !dx || !gamma || !save_mean || !save_invstd || !dgamma || !dbeta || B <= 0 || C <= 0) return SPV_ERR_ARG;
    const dim3 grid((C + BN_TX - 1) / BN_TX), block(BN_TX * BN_TY);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (B <= BN_TY * BN_R)
        bn_bwd_kernel<true><<<grid, block, 0, st>>>(dy, lddy, x, ldx, y_relu, ldy, dx, lddx, B, C, gamma, save_mean, save_invstd,
                                                    dgamma, dbeta);
    else if (B <= 8 * BN_TY * BN_R) {
        const int nrank = (B + BN_TY * BN_R - 1) / (BN_TY * BN_R);
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(grid.x, nrank);
        cfg.blockDim = block;
        cfg.dynamicSmemBytes = 0;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 1;
        attr[0].val.clusterDim.y = nrank;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        long lddy_ = lddy, ldx_ = ldx, ldy_ = ldy, lddx_ = lddx;
        if (cudaLaunchKernelEx(&cfg, bn_bwd_cluster_kernel, dy, lddy_, x, ldx_, y_relu, ldy_, dx, lddx_, B, C, gamma, save_mean, save_invstd,
                               dgamma, dbeta) != cudaSuccess)
            return SPV_ERR_LAUNCH;
    } else
        bn_bwd_kernel<false><<<grid, block, 0, st>>>(dy, lddy, x, ldx, y_relu, ldy, dx, lddx, B, C, gamma, save_mean, save_invstd,
                                                     dgamma, dbeta);
    SPV_CHECK_LAUNCH();
    return SPV_OK;
}

__global__ void __launch_bounds__(BN_TX* BN_TY) colsum_kernel(const float* __restrict__ x, long ldx, int B, int C, float* __restrict__ out) {
    __shared__ float red[BN_TY][BN_TX];
    const int tx = threadIdx.x % BN_TX, ty = threadIdx.x / BN_TX;
    const int c = blockIdx.x * BN_TX + tx;
    float s = 0.0f;
    if (c < C) {
#pragma unroll 8
        for (int b = ty; b < B; b += BN_TY) s += x[(long)b * ldx + c];
    }
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

// Optional bf16 staging: up to ADAM_MAX_SEGS row-major [rows, cols] blocks of the flat parameter vector are also written,
// freshly updated, as bf16 into the (padded, ld_dst >= cols) tensor-core operand buffers, so that the next step does not
// start with conversion kernels.  The last CTA to finish advances *step (ticket counter, reset to zero).
#define ADAM_MAX_SEGS 8
struct AdamSegs {
    long long begin[ADAM_MAX_SEGS], end[ADAM_MAX_SEGS], ld[ADAM_MAX_SEGS];
    __nv_bfloat16* dst[ADAM_MAX_SEGS];
    __nv_bfloat16* dst_lo[ADAM_MAX_SEGS];  // optional bf16 residual plane (split-operand GEMMs), same pitch
    int cols[ADAM_MAX_SEGS];
    int f16[ADAM_MAX_SEGS];                // destination holds fp16 instead of bf16 values (no residual plane then)
    float inv_cols[ADAM_MAX_SEGS];
    int n;
};

__device__ __forceinline__ void adam_stage(const AdamSegs& sg, long long idx, float val) {
#pragma unroll 1
    for (int s = 0; s < sg.n; ++s) {
        if (idx >= sg.begin[s] && idx < sg.end[s]) {
            const int j = (int)(idx - sg.begin[s]);  // blocks are < 2^24 elements wide enough for the float estimate + fix-up
            const int cols = sg.cols[s];
            int r = __float2int_rd((float)j * sg.inv_cols[s]);
            int c = j - r * cols;
            if (c < 0) { --r; c += cols; }
            else if (c >= cols) { ++r; c -= cols; }
            if (sg.f16[s]) {
                reinterpret_cast<__half*>(sg.dst[s])[(long long)r * sg.ld[s] + c] = to_half_sat(val);
                return;
            }
            const __nv_bfloat16 hi = __float2bfloat16(val);
            sg.dst[s][(long long)r * sg.ld[s] + c] = hi;
            if (sg.dst_lo[s]) sg.dst_lo[s][(long long)r * sg.ld[s] + c] = __float2bfloat16(val - __bfloat162float(hi));
            return;
        }
    }
}

__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                   float* __restrict__ v, long n, float lr, float b1, float b2, float eps, float wd,
                                                   float grad_scale, int* step, int* ticket, const __grid_constant__ AdamSegs sg) {
    const int t = *step + (ticket ? 1 : 0);  // with a ticket counter this launch owns the increment of the step count
    const float bc1 = 1.0f - powf(b1, (float)t), bc2 = 1.0f - powf(b2, (float)t);
    const float step_size = lr / bc1, inv_sqrt_bc2 = 1.0f / sqrtf(bc2);
    const long n4 = n >> 2;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n4; i += (long)gridDim.x * blockDim.x) {
        float4 p4 = reinterpret_cast<float4*>(p)[i];
        const float4 g4 = __ldg(reinterpret_cast<const float4*>(g) + i);
        float4 m4 = reinterpret_cast<float4*>(m)[i];
        float4 v4 = reinterpret_cast<float4*>(v)[i];
        float* pp = &p4.x;
        const float* gg = &g4.x;
        float* mm = &m4.x;
        float* vv = &v4.x;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float gi = gg[k] * grad_scale + wd * pp[k];
            mm[k] = b1 * mm[k] + (1.0f - b1) * gi;
            vv[k] = b2 * vv[k] + (1.0f - b2) * gi * gi;
            float denom = sqrtf(vv[k]) * inv_sqrt_bc2 + eps;
            pp[k] = pp[k] - step_size * (mm[k] / denom);
        }
        reinterpret_cast<float4*>(m)[i] = m4;
        reinterpret_cast<float4*>(v)[i] = v4;
        reinterpret_cast<float4*>(p)[i] = p4;
        if (sg.n > 0) {
#pragma unroll
            for (int k = 0; k < 4; ++k) adam_stage(sg, 4 * (long long)i + k, pp[k]);
        }
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {  // tail (the flat stores are padded to multiples of 4; kept for generality)
        long i = (n4 << 2) + threadIdx.x;
        float pi = p[i];
        float gi = g[i] * grad_scale + wd * pi;
        float mi = b1 * m[i] + (1.0f - b1) * gi;
        float vi = b2 * v[i] + (1.0f - b2) * gi * gi;
        m[i] = mi;
        v[i] = vi;
        pi -= step_size * (mi / (sqrtf(vi) * inv_sqrt_bc2 + eps));
        p[i] = pi;
        if (sg.n > 0) adam_stage(sg, i, pi);
    }
    if (!ticket) return;
    __syncthreads();  // every thread of this CTA has read *step
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(ticket, 1) == (int)gridDim.x - 1) {
            *step = t;
            *ticket = 0;
        }
    }
}

// step: device counter of completed optimiser steps.  ticket == null: *step must already hold the 1-based index of this
// step (spv_adam_tick before the call); ticket != null (a zeroed device int): the launch uses *step + 1 and stores it back.
// Staging segments (nseg <= 8, host arrays): parameter block [seg_begin, seg_begin + seg_rows * seg_cols) viewed as
// [seg_rows, seg_cols] is mirrored as bf16 into seg_dst with row pitch seg_ld.
extern "C" int spv_adam(float* p, const float* g, float* m, float* v, long long n, float lr, float b1, float b2, float eps,
                        float wd, float grad_scale, int* step, int* ticket, int nseg, const long long* seg_begin,
                        const int* seg_rows, const int* seg_cols, void* const* seg_dst, void* const* seg_dst_lo,
                        const int* seg_f16, const long long* seg_ld, int max_blocks, void* stream) {
    if (!p || !g || !m || !v || !step || n <= 0 || nseg < 0 || nseg > ADAM_MAX_SEGS) return SPV_ERR_ARG;
    if (nseg > 0 && (!seg_begin || !seg_rows || !seg_cols || !seg_dst || !seg_ld)) return SPV_ERR_ARG;
    if ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
         reinterpret_cast<uintptr_t>(v)) & 15)
        return SPV_ERR_ARG;
    AdamSegs sg;
    sg.n = nseg;
    for (int s = 0; s < nseg; ++s) {
        if (seg_rows[s] <= 0 || seg_cols[s] <= 0 || !seg_dst[s] || seg_ld[s] < seg_cols[s] ||
            (long long)seg_rows[s] * seg_cols[s] >= (1ll << 24))
            return SPV_ERR_ARG;
        sg.begin[s] = seg_begin[s];
        sg.end[s] = seg_begin[s] + (long long)seg_rows[s] * seg_cols[s];
        sg.ld[s] = seg_ld[s];
        sg.dst[s] = reinterpret_cast<__nv_bfloat16*>(seg_dst[s]);
        sg.dst_lo[s] = seg_dst_lo ? reinterpret_cast<__nv_bfloat16*>(seg_dst_lo[s]) : nullptr;
        sg.f16[s] = seg_f16 ? seg_f16[s] : 0;
        sg.cols[s] = seg_cols[s];
        sg.inv_cols[s] = 1.0f / (float)seg_cols[s];
    }
    int blocks = (int)min((long long)148 * 8, (n / 4 + 255) / 256);
    if (max_blocks > 0 && blocks > max_blocks) blocks = max_blocks;  // a small grid leaves SMs to concurrently running kernels
    if (blocks < 1) blocks = 1;
    adam_kernel<<<blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(p, g, m, v, n, lr, b1, b2, eps, wd, grad_scale, step,
                                                                            ticket, sg);
    SPV_CHECK_LAUNCH();
    return SPV_OK;
}

// ---------------------------------------------------------------------------------------
// batch covariates (n_batch > 1): the one-hot batch code is appended to the input of the encoders' first layer and of the four
// decoder nets (reference nn/utils.py:9-13, nn/networks.py:110-118, scvi FCLayers inject_covariates).
// ---------------------------------------------------------------------------------------
__global__ void one_hot_kernel(const int* __restrict__ code, float* __restrict__ out, long ld, int B, int nb) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * nb) return;
    const int b = i / nb, k = i - b * nb;
    out[(long)b * ld + k] = code[b] == k ? 1.0f : 0.0f;
}

// out[b, 0:nb] = one_hot(code[b])
extern "C" int spv_one_hot(const int* code, float* out, long long ld, int B, int nb, void* stream) {
    if (!code || !out || B <= 0 || nb <= 0 || ld < nb) return SPV_ERR_ARG;
    one_hot_kernel<<<(B * nb + 255) / 256, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(code, out, ld, B, nb);
    SPV_CHECK_LAUNCH();
    return SPV_OK;
}

// zzb[b] = [zz[b, 0:P] | oh | zz[b, P:P+S] | oh]: the inputs of the two factor regressors side by side, each with its covariate
// columns (their weights are [G, P + nb] and [G, S + nb]); oh_tail[b, 0:nb] = oh (the covariate columns behind [hm | zz] in the
// mixing net's input)
__global__ void cov_expand_kernel(const float* __restrict__ zz, long ld_zz, const int* __restrict__ code, float* __restrict__ zzb,
                                  long ld_zzb, float* __restrict__ oh_tail, long ld_oh, int B, int P, int S, int nb) {
    const int W = P + S + 2 * nb;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * W) return;
    const int b = i / W, c = i - b * W;
    float v;
    if (c < P) v = zz[(long)b * ld_zz + c];
    else if (c < P + nb) v = code[b] == c - P ? 1.0f : 0.0f;
    else if (c < P + nb + S) v = zz[(long)b * ld_zz + c - nb];
    else {
        v = code[b] == c - (P + nb + S) ? 1.0f : 0.0f;
        if (oh_tail) oh_tail[(long)b * ld_oh + c - (P + nb + S)] = v;
    }
    zzb[(long)b * ld_zzb + c] = v;
}

extern "C" int spv_cov_expand(const float* zz, long long ld_zz, const int* code, float* zzb, long long ld_zzb, float* oh_tail,
                              long long ld_oh, int B, int P, int S, int nb, void* stream) {
    if (!zz || !code || !zzb || B <= 0 || P <= 0 || S <= 0 || nb <= 0) return SPV_ERR_ARG;
    const int total = B * (P + S + 2 * nb);
    cov_expand_kernel<<<(total + 255) / 256, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(zz, ld_zz, code, zzb, ld_zzb, oh_tail, ld_oh,
                                                                                              B, P, S, nb);
    SPV_CHECK_LAUNCH();
    return SPV_OK;
}

// the reverse for gradients: dzz[b] = [dzzb[b, 0:P] | dzzb[b, P+nb : P+nb+S]] (the covariate columns carry no gradient)
__global__ void cov_compact_kernel(const float* __restrict__ dzzb, long ld_in, float* __restrict__ dzz, long ld_out, int B, int P, int S,
                                   int nb, const float* __restrict__ add, long ld_add) {
    const int W = P + S;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * W) return;
    const int b = i / W, c = i - b * W;
    dzz[(long)b * ld_out + c] = dzzb[(long)b * ld_in + (c < P ? c : c + nb)] + (add ? add[(long)b * ld_add + c] : 0.0f);
}

extern "C" int spv_cov_compact(const float* dzzb, long long ld_in, float* dzz, long long ld_out, int B, int P, int S, int nb,
                               const float* add, long long ld_add, void* stream) {
    if (!dzzb || !dzz || B <= 0 || P <= 0 || S <= 0 || nb <= 0) return SPV_ERR_ARG;
    const int total = B * (P + S);
    cov_compact_kernel<<<(total + 255) / 256, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(dzzb, ld_in, dzz, ld_out, B, P, S, nb,
                                                                                               add, ld_add);
    SPV_CHECK_LAUNCH();
    return SPV_OK;
}
