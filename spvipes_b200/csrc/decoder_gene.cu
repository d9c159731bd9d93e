// Per-gene stages of the decoder: closed-form BatchNorm statistics of the two "factor regressor" branches
// (u = z W^T, BatchNorm1d over the minibatch, eps 1e-3, momentum 0.01: scvi FCLayers; reference nn/networks.py:314-320),
// their fold into an affine map, the constants of the NB term (theta = exp(px_r), reference module/spVIPESmodule.py:758),
// and the matching backward.  Because u is linear in z, the batch mean / variance of u[:, g] are W[g] . mean(z) and
// W[g]^T Cov(z) W[g]: no [B, G] pass is needed, only the [KZ, KZ] covariance of the latent minibatch.
//
// Work split: one warp per gene (lanes over the latent dimension), 64 genes per CTA.
#include <cuda_bf16.h>
#include "common.cuh"
#include "decoder_common.cuh"
#include "../../include/spvipes_b200.h"

#define GENES_PER_CTA 64        // gene_bwd: 64 genes share one set of per-CTA partial sums
#define FOLD_GENES_PER_CTA 8    // fold: one gene per warp
#define GENE_BWD_THREADS 1024

// ---------------------------------------------------------------------------------------
// partial (un-normalised, centred) second moments of zz over a chunk of 64 rows
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) zcov_kernel(const float* __restrict__ zz, long ld, int B, int KZ,
                                                   const float* __restrict__ zsum, float* cov_part,
                                                   float* __restrict__ zmean_out, float* __restrict__ zcov_out) {
    extern __shared__ float tile[];  // [64][KZ + 1]
    const int r0 = blockIdx.x * 64;
    const int ldt = KZ + 1;
    const float invB = 1.0f / (float)B;
    for (int i = threadIdx.x; i < 64 * KZ; i += blockDim.x) {
        int r = i / KZ, k = i % KZ;
        float v = 0.0f;
        if (r0 + r < B) v = zz[(long)(r0 + r) * ld + k] - zsum[k] * invB;
        tile[r * ldt + k] = v;
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < KZ * KZ; idx += blockDim.x) {
        int i = idx / KZ, j = idx % KZ;
        float s = 0.0f;
#pragma unroll 8
        for (int r = 0; r < 64; ++r) s = fmaf(tile[r * ldt + i], tile[r * ldt + j], s);
        cov_part[(long)blockIdx.x * KZ * KZ + idx] = s;
    }
    // the last CTA to finish sums the partials in chunk order (deterministic) into the final mean / covariance
    __shared__ int is_last;
    int* counter = reinterpret_cast<int*>(cov_part + (long)gridDim.x * KZ * KZ);  // one spare slot, zero between launches
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        int ticket = atomicAdd(counter, 1);
        is_last = ticket == (int)gridDim.x - 1;
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    for (int k = threadIdx.x; k < KZ; k += blockDim.x) zmean_out[k] = zsum[k] * invB;
    for (int idx = threadIdx.x; idx < KZ * KZ; idx += blockDim.x) {
        float s = 0.0f;
        for (int c = 0; c < (int)gridDim.x; ++c) s += __ldcg(cov_part + (long)c * KZ * KZ + idx);
        zcov_out[idx] = s * invB;
    }
    if (threadIdx.x == 0) *counter = 0;
}

struct FoldP {
    const float *Wp, *Ws, *gp, *bp, *gs, *bs, *px_r;
    float *rm_p, *rv_p, *rm_s, *rv_s;
    const float *zsum, *cov_part;
    float *wfold, *genec, *zmean, *zcov;
    // optional: rows [Gp, 3 Gp) of the stacked bf16 operand [3 Gp, ld_wz] (private block then shared block); the folded weights
    // go into the latent columns HD .. HD + P + S of their block, every other entry of those rows stays zero
    __nv_bfloat16* wfold_bf16;
    long ld_wz;
    int Gp, HD;
    int G, P, S, B, ncov, training;
    float eps, momentum;
};

__global__ void __launch_bounds__(256) fold_kernel(FoldP p) {
    extern __shared__ float sh[];  // mean[KZ] | cov[KZ*KZ]
    const int KZ = p.P + p.S;
    float* smean = sh;
    float* scov = sh + KZ;
    if (p.training) {  // final mean / covariance of the latent minibatch (zcov_kernel)
        for (int k = threadIdx.x; k < KZ; k += blockDim.x) smean[k] = p.zmean[k];
        for (int i = threadIdx.x; i < KZ * KZ; i += blockDim.x) scov[i] = p.zcov[i];
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long G = p.G;
    {  // one gene per warp: every gene's chain of dependent loads runs concurrently
        const int g = blockIdx.x * FOLD_GENES_PER_CTA + warp;
        if (g >= p.G) return;
#pragma unroll
        for (int br = 0; br < 2; ++br) {
            const int K = br == 0 ? p.P : p.S;
            const int off = br == 0 ? 0 : p.P;
            const float* W = (br == 0 ? p.Wp : p.Ws) + (long)g * K;
            float* rm = br == 0 ? p.rm_p : p.rm_s;
            float* rv = br == 0 ? p.rv_p : p.rv_s;
            float mean, var;
            if (p.training) {
                float pm = 0.0f, pv = 0.0f;
                for (int k = lane; k < K; k += 32) {
                    float wk = W[k];
                    pm = fmaf(smean[off + k], wk, pm);
                    float t = 0.0f;
                    for (int l = 0; l < K; ++l) t = fmaf(scov[(off + k) * KZ + off + l], __ldg(W + l), t);
                    pv = fmaf(wk, t, pv);
                }
                mean = warp_sum(pm);
                var = fmaxf(warp_sum(pv), 0.0f);
                if (lane == 0) {
                    float unb = var * ((float)p.B / (float)max(p.B - 1, 1));
                    rm[g] = (1.0f - p.momentum) * rm[g] + p.momentum * mean;
                    rv[g] = (1.0f - p.momentum) * rv[g] + p.momentum * unb;
                }
            } else {
                mean = rm[g];
                var = rv[g];
            }
            float invstd = 1.0f / sqrtf(var + p.eps);
            float a = (br == 0 ? p.gp : p.gs)[g] * invstd;
            float c = (br == 0 ? p.bp : p.bs)[g] - mean * a;
            for (int k = lane; k < K; k += 32) p.wfold[(long)g * KZ + off + k] = a * W[k];
            if (p.wfold_bf16)  // bf16 copy inside the stacked tensor-core operand: block br, row g, latent columns HD + off ..
                for (int k = lane; k < K; k += 32)
                    p.wfold_bf16[((long)br * p.Gp + g) * p.ld_wz + p.HD + off + k] = __float2bfloat16(a * W[k]);
            if (lane == 0) {
                p.genec[(br == 0 ? GC_CP : GC_CS) * G + g] = c;
                p.genec[(br == 0 ? GC_AP : GC_AS) * G + g] = a;
                p.genec[(br == 0 ? GC_ISTD_P : GC_ISTD_S) * G + g] = invstd;
                p.genec[(br == 0 ? GC_MEAN_P : GC_MEAN_S) * G + g] = mean;
            }
        }
        if (lane == 0) {
            float th = expf(p.px_r[g]);  // reference module/spVIPESmodule.py:758
            p.genec[GC_THETA * G + g] = th;
            p.genec[GC_LTE * G + g] = logf(th + NB_EPS);
            p.genec[GC_LGT * G + g] = lgammaf(th);
            p.genec[GC_DGT * G + g] = digammaf_pos(th);
        }
    }
}

// ptrs: Wp, Ws, gamma_p, beta_p, gamma_s, beta_s, px_r, rm_p, rv_p, rm_s, rv_s, zz, zsum, cov_part, wfold, genec, zmean, zcov
extern "C" int spv_dec_fold(const void* const* ptrs, long long ld_zz, int B, int G, int P, int S, int training, float eps,
                            float momentum, void* wz_bf16, long long ld_wz, int Gp, int HD, void* stream) {
    if (!ptrs || B <= 0 || G <= 0 || P <= 0 || S <= 0 || P + S > 96) return SPV_ERR_ARG;
    for (int i = 0; i < 18; ++i)
        if (!ptrs[i]) return SPV_ERR_ARG;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int KZ = P + S;
    const float* zz = (const float*)ptrs[11];
    const float* zsum = (const float*)ptrs[12];
    float* cov_part = (float*)ptrs[13];
    const int ncov = (B + 63) / 64;
    if (training) {
        size_t sm1 = (size_t)64 * (KZ + 1) * sizeof(float);
        zcov_kernel<<<ncov, 256, sm1, st>>>(zz, ld_zz, B, KZ, zsum, cov_part, (float*)ptrs[16], (float*)ptrs[17]);
        SPV_CHECK_LAUNCH();
    }
    FoldP p;
    p.Wp = (const float*)ptrs[0]; p.Ws = (const float*)ptrs[1]; p.gp = (const float*)ptrs[2]; p.bp = (const float*)ptrs[3];
    p.gs = (const float*)ptrs[4]; p.bs = (const float*)ptrs[5]; p.px_r = (const float*)ptrs[6];
    p.rm_p = (float*)ptrs[7]; p.rv_p = (float*)ptrs[8]; p.rm_s = (float*)ptrs[9]; p.rv_s = (float*)ptrs[10];
    p.zsum = zsum; p.cov_part = cov_part; p.wfold = (float*)ptrs[14]; p.genec = (float*)ptrs[15];
    p.zmean = (float*)ptrs[16]; p.zcov = (float*)ptrs[17];
    p.wfold_bf16 = reinterpret_cast<__nv_bfloat16*>(wz_bf16);
    p.ld_wz = ld_wz; p.Gp = Gp; p.HD = HD;
    p.G = G; p.P = P; p.S = S; p.B = B; p.ncov = training ? ncov : 0; p.training = training; p.eps = eps; p.momentum = momentum;
    size_t sm2 = (size_t)(KZ + KZ * KZ) * sizeof(float);
    if (sm2 > 48 * 1024) cudaFuncSetAttribute(fold_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm2);
    fold_kernel<<<(G + FOLD_GENES_PER_CTA - 1) / FOLD_GENES_PER_CTA, 32 * FOLD_GENES_PER_CTA, sm2, st>>>(p);
    SPV_CHECK_LAUNCH();
    return SPV_OK;
}

// ---------------------------------------------------------------------------------------
// per-gene backward of the two folded BatchNorm+Linear branches (closed form):
//   Q = dy^T z (from spv_gemm), sdy = colsum(dy):  S2 = (Q.W - sdy mean_u) invstd = dgamma,  dbeta = sdy,
//   dW = a (Q - sdy zbar - S2 invstd Cov W)
// and the per-CTA partial sums of the two operands of the d z correction (BatchNorm couples all cells of a gene):
//   v1[c]   = sum_g (a sdy / B) W[g, c]                 -> vpart[cta, c]
//   M[k, l] = sum_g (a S2 invstd / B) W[g, k] W[g, l]   -> mpart[cta, k * KZ + l]   (block diagonal: private, shared)
// also d px_r = theta * colsum(dtheta), d bm = colsum(dpi).
// ---------------------------------------------------------------------------------------
struct GeneBwdP {
    const float *Wp, *Ws, *Qp, *Qs, *genec, *colsum, *zmean, *zcov;
    float *dWp, *dWs, *dgp, *dbp, *dgs, *dbs, *dpx_r, *dbm, *vpart, *mpart;
    int G, P, S, B;
    long ldq;  // row pitch of Qp / Qs (0: packed, P resp. S)
};

__global__ void __launch_bounds__(GENE_BWD_THREADS) gene_bwd_kernel(GeneBwdP p) {
    extern __shared__ float sh[];
    const int KZ = p.P + p.S;
    float* smean = sh;                         // [KZ]
    float* scov = smean + KZ;                  // [KZ * KZ]
    float* sW = scov + KZ * KZ;                // [64][KZ]  (private | shared weights of this CTA's genes)
    float* scv = sW + GENES_PER_CTA * KZ;      // [2][64]
    float* scm = scv + 2 * GENES_PER_CTA;      // [2][64]
    const int g0 = blockIdx.x * GENES_PER_CTA;
    for (int k = threadIdx.x; k < KZ; k += blockDim.x) smean[k] = p.zmean[k];
    for (int i = threadIdx.x; i < KZ * KZ; i += blockDim.x) scov[i] = p.zcov[i];
    for (int i = threadIdx.x; i < GENES_PER_CTA * KZ; i += blockDim.x) {
        int gl = i / KZ, k = i - gl * KZ, g = g0 + gl;
        float v = 0.0f;
        if (g < p.G) v = k < p.P ? p.Wp[(long)g * p.P + k] : p.Ws[(long)g * p.S + (k - p.P)];
        sW[i] = v;
    }
    for (int i = threadIdx.x; i < 2 * GENES_PER_CTA; i += blockDim.x) { scv[i] = 0.0f; scm[i] = 0.0f; }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long G = p.G;
    const float invB = 1.0f / (float)p.B;
    for (int gl = warp; gl < GENES_PER_CTA; gl += (int)(blockDim.x >> 5)) {
        const int g = g0 + gl;
        if (g >= p.G) break;
#pragma unroll
        for (int br = 0; br < 2; ++br) {
            const int K = br == 0 ? p.P : p.S;
            const int off = br == 0 ? 0 : p.P;
            const float* W = sW + gl * KZ + off;
            const float* Q = (br == 0 ? p.Qp : p.Qs) + (long)g * (p.ldq > 0 ? p.ldq : K);
            float* dW = (br == 0 ? p.dWp : p.dWs) + (long)g * K;
            const float a = p.genec[(br == 0 ? GC_AP : GC_AS) * G + g];
            const float invstd = p.genec[(br == 0 ? GC_ISTD_P : GC_ISTD_S) * G + g];
            const float mean_u = p.genec[(br == 0 ? GC_MEAN_P : GC_MEAN_S) * G + g];
            const float sdy = p.colsum[(long)br * G + g];
            float qw = 0.0f;
            for (int k = lane; k < K; k += 32) qw = fmaf(Q[k], W[k], qw);
            qw = warp_sum(qw);
            const float S2 = (qw - sdy * mean_u) * invstd;
            if (lane == 0) {
                (br == 0 ? p.dgp : p.dgs)[g] = S2;
                (br == 0 ? p.dbp : p.dbs)[g] = sdy;
                scv[br * GENES_PER_CTA + gl] = a * sdy * invB;
                scm[br * GENES_PER_CTA + gl] = a * S2 * invB * invstd;
            }
            for (int k = lane; k < K; k += 32) {
                float cw = 0.0f;
                for (int l = 0; l < K; ++l) cw = fmaf(scov[(off + k) * KZ + off + l], W[l], cw);
                dW[k] = a * (Q[k] - sdy * smean[off + k] - S2 * invstd * cw);
            }
        }
        if (lane == 0) {
            p.dpx_r[g] = p.genec[GC_THETA * G + g] * p.colsum[3 * G + g];
            p.dbm[g] = p.colsum[2 * G + g];
        }
    }
    __syncthreads();
    // per-CTA partials of v1 and M
    for (int c = threadIdx.x; c < KZ; c += blockDim.x) {
        const float* cv = scv + (c < p.P ? 0 : GENES_PER_CTA);
        float s = 0.0f;
        for (int gl = 0; gl < GENES_PER_CTA; ++gl) s = fmaf(cv[gl], sW[gl * KZ + c], s);
        p.vpart[(long)blockIdx.x * KZ + c] = s;
    }
    const int nPP = p.P * p.P, nSS = p.S * p.S;
    for (int idx = threadIdx.x; idx < nPP + nSS; idx += blockDim.x) {
        int k, l;
        const float* cm;
        if (idx < nPP) { k = idx / p.P; l = idx - k * p.P; cm = scm; }
        else { int j = idx - nPP; k = p.P + j / p.S; l = p.P + j % p.S; cm = scm + GENES_PER_CTA; }
        float s = 0.0f;
        for (int gl = 0; gl < GENES_PER_CTA; ++gl) s = fmaf(cm[gl] * sW[gl * KZ + k], sW[gl * KZ + l], s);
        p.mpart[(long)blockIdx.x * KZ * KZ + k * KZ + l] = s;
    }
}

// ptrs: Wp, Ws, Qp, Qs, genec, colsum, zmean, zcov, dWp, dWs, dgamma_p, dbeta_p, dgamma_s, dbeta_s, dpx_r, dbm,
//       vpart [ceil(G/64), KZ], mpart [ceil(G/64), KZ*KZ]
extern "C" int spv_dec_gene_bwd(const void* const* ptrs, long long ldq, int B, int G, int P, int S, void* stream) {
    if (!ptrs || B <= 0 || G <= 0 || P <= 0 || S <= 0 || P + S > 96) return SPV_ERR_ARG;
    for (int i = 0; i < 18; ++i)
        if (!ptrs[i]) return SPV_ERR_ARG;
    GeneBwdP p;
    p.Wp = (const float*)ptrs[0]; p.Ws = (const float*)ptrs[1]; p.Qp = (const float*)ptrs[2]; p.Qs = (const float*)ptrs[3];
    p.genec = (const float*)ptrs[4]; p.colsum = (const float*)ptrs[5]; p.zmean = (const float*)ptrs[6];
    p.zcov = (const float*)ptrs[7]; p.dWp = (float*)ptrs[8]; p.dWs = (float*)ptrs[9]; p.dgp = (float*)ptrs[10];
    p.dbp = (float*)ptrs[11]; p.dgs = (float*)ptrs[12]; p.dbs = (float*)ptrs[13]; p.dpx_r = (float*)ptrs[14];
    p.dbm = (float*)ptrs[15]; p.vpart = (float*)ptrs[16]; p.mpart = (float*)ptrs[17];
    p.G = G; p.P = P; p.S = S; p.B = B; p.ldq = ldq;
    const int KZ = P + S;
    size_t smem = (size_t)(KZ + KZ * KZ + GENES_PER_CTA * KZ + 4 * GENES_PER_CTA) * sizeof(float);
    if (smem > 48 * 1024) cudaFuncSetAttribute(gene_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    gene_bwd_kernel<<<(G + GENES_PER_CTA - 1) / GENES_PER_CTA, GENE_BWD_THREADS, smem, reinterpret_cast<cudaStream_t>(stream)>>>(p);
    SPV_CHECK_LAUNCH();
    return SPV_OK;
}

// dzz[b, c] = dmix[b, c] + dzraw[b, c] - v1[c] - sum_l M[c, l] (zz[b, l] - zbar[l])   (l within c's branch block)
//   dmix: the zz columns of d Amix (mixture GEMM);  dzraw: further addends (dy W' of the softmax branches when they are
//   not part of dmix already, and dah Wh of the mixing net's hidden layer), may be null;
//   v1 / M: sums over the `nparts` per-CTA partials written by spv_dec_gene_bwd (fixed order: deterministic).
// One latent column per CTA (blockIdx.x) and 256 rows (blockIdx.y): the CTA only needs row c of M, i.e. at most
// max(P, S) + 1 sums over the partials, taken by 8 part lanes x 32 entries and combined in lane order.
#define DZC_THREADS 256
__global__ void __launch_bounds__(DZC_THREADS) dzz_combine_kernel(const float* __restrict__ dmix, long ld_dmix,
                                                                  const float* __restrict__ dzraw, const float* __restrict__ vpart,
                                                                  const float* __restrict__ mpart, int nparts,
                                                                  const float* __restrict__ zz, long ld_zz,
                                                                  const float* __restrict__ zmean, float* __restrict__ dzz, int B,
                                                                  int P, int S) {
    __shared__ float red[8][100];
    __shared__ float srow[100];  // [0, n): M[c, lo + .],  [n]: v1[c]
    const int KZ = P + S;
    const int c = blockIdx.x;
    const int lo = c < P ? 0 : P, n = c < P ? P : S;
    const int lane = threadIdx.x & 31, q = threadIdx.x >> 5;
    for (int e0 = 0; e0 <= n; e0 += 32) {
        const int e = e0 + lane;
        float s = 0.0f;
        if (e <= n) {
            const float* src = e < n ? mpart + (long)c * KZ + lo + e : vpart + c;
            const long pitch = e < n ? (long)KZ * KZ : KZ;
#pragma unroll 4
            for (int t = q; t < nparts; t += 8) s += src[(long)t * pitch];
            red[q][e] = s;
        }
    }
    __syncthreads();
    if (threadIdx.x <= n) {
        float s = 0.0f;
#pragma unroll
        for (int i = 0; i < 8; ++i) s += red[i][threadIdx.x];
        srow[threadIdx.x] = s;
    }
    __syncthreads();
    const int b = blockIdx.y * DZC_THREADS + threadIdx.x;
    if (b >= B) return;
    const float* zrow = zz + (long)b * ld_zz + lo;
    float corr = 0.0f;
    for (int l = 0; l < n; ++l) corr = fmaf(srow[l], zrow[l] - zmean[lo + l], corr);
    dzz[(long)b * KZ + c] = dmix[(long)b * ld_dmix + c] + (dzraw ? dzraw[(long)b * KZ + c] : 0.0f) - srow[n] - corr;
}

extern "C" int spv_dec_dzz_combine(const float* dmix, long long ld_dmix, const float* dzraw, const float* vpart, const float* mpart,
                                   int nparts, const float* zz, long long ld_zz, const float* zmean, float* dzz, int B, int P,
                                   int S, void* stream) {
    if (!dmix || !vpart || !mpart || !zz || !zmean || !dzz || B <= 0 || P <= 0 || S <= 0 || nparts <= 0 || P > 96 || S > 96)
        return SPV_ERR_ARG;
    dim3 grid(P + S, (B + DZC_THREADS - 1) / DZC_THREADS);
    dzz_combine_kernel<<<grid, DZC_THREADS, 0, reinterpret_cast<cudaStream_t>(stream)>>>(dmix, ld_dmix, dzraw, vpart, mpart, nparts,
                                                                                         zz, ld_zz, zmean, dzz, B, P, S);
    SPV_CHECK_LAUNCH();
    return SPV_OK;
}
