// Decoder + negative-binomial-mixture likelihood, fp32 SIMT path.
//   rho_p = exp(lib) softmax_g(BN_p(z_p Wp^T)),  rho_s = exp(lib) softmax_g(BN_s(z_s Ws^T)),
//   pi = [relu(BN_h(zz Wh^T + bh)) | zz] Wm^T + bm,  rec_b = -sum_g log_mixture_nb(log1p(x), rho_p, rho_s, theta, pi)
// Reference: nn/networks.py:314-325 (LinearDecoderSPVIPE.forward), module/spVIPESmodule.py:751-759 (generative),
// :817-824 (loss), scvi-tools 0.20.0 log_mixture_nb (eps 1e-8, shared theta) and FCLayers BatchNorm1d(eps 1e-3, mom 0.01).
//
// No [B, G] softmax / rate tensor is materialised: the per-gene BatchNorm statistics of z W^T are obtained in closed
// form from mean(z) and Cov(z) (spv_dec_fold), the gene-axis softmax normaliser by a first tile sweep
// (pass STATS), the likelihood and the row sums the backward needs by a second sweep (pass NB).
#include <cuda_bf16.h>
#include "gemm_simt.cuh"
#include "../../include/spvipes_b200.h"

// genec rows (SoA, stride G)
enum { GC_CP = 0, GC_CS, GC_AP, GC_AS, GC_ISTD_P, GC_ISTD_S, GC_MEAN_P, GC_MEAN_S, GC_THETA, GC_LTE, GC_LGT, GC_DGT, GC_N };

// ---------------------------------------------------------------------------------------
// partial (un-normalised, centred) second moments of zz over a chunk of 64 rows
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) zcov_kernel(const float* __restrict__ zz, long ld, int B, int KZ,
                                                   const float* __restrict__ zsum, float* __restrict__ cov_part) {
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
}

// ---------------------------------------------------------------------------------------
// per gene: closed-form BatchNorm statistics of u = z W^T, folded affine, running-stat update,
// and the constants of the NB term
// ---------------------------------------------------------------------------------------
struct FoldP {
    const float *Wp, *Ws, *gp, *bp, *gs, *bs, *px_r;
    float *rm_p, *rv_p, *rm_s, *rv_s;
    const float *zsum, *cov_part;
    float *wfold, *genec, *zmean, *zcov;
    int G, P, S, B, ncov, training;
    float eps, momentum;
};

__global__ void __launch_bounds__(256) fold_kernel(FoldP p) {
    extern __shared__ float sh[];  // mean[KZ] | cov[KZ*KZ]
    const int KZ = p.P + p.S;
    float* smean = sh;
    float* scov = sh + KZ;
    const float invB = 1.0f / (float)p.B;
    for (int k = threadIdx.x; k < KZ; k += blockDim.x) smean[k] = p.zsum[k] * invB;
    for (int i = threadIdx.x; i < KZ * KZ; i += blockDim.x) {
        float s = 0.0f;
        for (int c = 0; c < p.ncov; ++c) s += p.cov_part[(long)c * KZ * KZ + i];
        scov[i] = s * invB;
    }
    __syncthreads();
    if (blockIdx.x == 0) {
        for (int k = threadIdx.x; k < KZ; k += blockDim.x) p.zmean[k] = smean[k];
        for (int i = threadIdx.x; i < KZ * KZ; i += blockDim.x) p.zcov[i] = scov[i];
    }
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= p.G) return;
    const int G = p.G;
#pragma unroll
    for (int br = 0; br < 2; ++br) {
        const int K = br == 0 ? p.P : p.S;
        const int off = br == 0 ? 0 : p.P;
        const float* W = (br == 0 ? p.Wp : p.Ws) + (long)g * K;
        float* rm = br == 0 ? p.rm_p : p.rm_s;
        float* rv = br == 0 ? p.rv_p : p.rv_s;
        float mean, var;
        if (p.training) {
            mean = 0.0f;
            var = 0.0f;
            for (int k = 0; k < K; ++k) {
                float wk = W[k];
                mean = fmaf(smean[off + k], wk, mean);
                float t = 0.0f;
                for (int l = 0; l < K; ++l) t = fmaf(scov[(off + k) * KZ + off + l], W[l], t);
                var = fmaf(wk, t, var);
            }
            var = fmaxf(var, 0.0f);
            float unb = var * ((float)p.B / (float)max(p.B - 1, 1));
            rm[g] = (1.0f - p.momentum) * rm[g] + p.momentum * mean;
            rv[g] = (1.0f - p.momentum) * rv[g] + p.momentum * unb;
        } else {
            mean = rm[g];
            var = rv[g];
        }
        float invstd = 1.0f / sqrtf(var + p.eps);
        float a = (br == 0 ? p.gp : p.gs)[g] * invstd;
        float c = (br == 0 ? p.bp : p.bs)[g] - mean * a;
        for (int k = 0; k < K; ++k) p.wfold[(long)g * KZ + off + k] = a * W[k];
        p.genec[(br == 0 ? GC_CP : GC_CS) * (long)G + g] = c;
        p.genec[(br == 0 ? GC_AP : GC_AS) * (long)G + g] = a;
        p.genec[(br == 0 ? GC_ISTD_P : GC_ISTD_S) * (long)G + g] = invstd;
        p.genec[(br == 0 ? GC_MEAN_P : GC_MEAN_S) * (long)G + g] = mean;
    }
    float th = expf(p.px_r[g]);  // reference module/spVIPESmodule.py:758
    p.genec[GC_THETA * (long)G + g] = th;
    p.genec[GC_LTE * (long)G + g] = logf(th + NB_EPS);
    p.genec[GC_LGT * (long)G + g] = lgammaf(th);
    p.genec[GC_DGT * (long)G + g] = digammaf_pos(th);
}

// ptrs: Wp, Ws, gamma_p, beta_p, gamma_s, beta_s, px_r, rm_p, rv_p, rm_s, rv_s, zz, zsum, cov_part, wfold, genec, zmean, zcov
extern "C" int spv_dec_fold(const void* const* ptrs, long long ld_zz, int B, int G, int P, int S, int training, float eps,
                            float momentum, void* stream) {
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
        zcov_kernel<<<ncov, 256, sm1, st>>>(zz, ld_zz, B, KZ, zsum, cov_part);
        SPV_CHECK_LAUNCH();
    }
    FoldP p;
    p.Wp = (const float*)ptrs[0]; p.Ws = (const float*)ptrs[1]; p.gp = (const float*)ptrs[2]; p.bp = (const float*)ptrs[3];
    p.gs = (const float*)ptrs[4]; p.bs = (const float*)ptrs[5]; p.px_r = (const float*)ptrs[6];
    p.rm_p = (float*)ptrs[7]; p.rv_p = (float*)ptrs[8]; p.rm_s = (float*)ptrs[9]; p.rv_s = (float*)ptrs[10];
    p.zsum = zsum; p.cov_part = cov_part; p.wfold = (float*)ptrs[14]; p.genec = (float*)ptrs[15];
    p.zmean = (float*)ptrs[16]; p.zcov = (float*)ptrs[17];
    p.G = G; p.P = P; p.S = S; p.B = B; p.ncov = training ? ncov : 0; p.training = training; p.eps = eps; p.momentum = momentum;
    size_t sm2 = (size_t)(KZ + KZ * KZ) * sizeof(float);
    if (sm2 > 48 * 1024) cudaFuncSetAttribute(fold_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm2);
    fold_kernel<<<(G + 255) / 256, 256, sm2, st>>>(p);
    SPV_CHECK_LAUNCH();
    return SPV_OK;
}

// ---------------------------------------------------------------------------------------
// the NB-mixture element (scvi log_mixture_nb, shared theta).  t = log1p(count) (quirk Q3).
// ---------------------------------------------------------------------------------------
struct NbFwd { float ll, ep, es; };

__device__ __forceinline__ NbFwd nb_forward(float t, float lp, float ls, float pi, float th, float lte, float lgt,
                                            float Rp, float Rs) {
    float rp = expf(lp + Rp), rs = expf(ls + Rs);  // exp(lib) * softmax
    float d1 = th + rp + NB_EPS, d2 = th + rs + NB_EPS;
    float l1 = logf(d1), l2 = logf(d2);
    float a = th * (lte - l1), b = th * (lte - l2);
    float gp = 0.0f, gs = 0.0f;
    if (t != 0.0f) {
        float lg = lgammaf(t + th) - lgt - lgammaf(t + 1.0f);
        a += t * (logf(rp + NB_EPS) - l1) + lg;
        b += t * (logf(rs + NB_EPS) - l2) + lg;
        gp = t / (rp + NB_EPS);
        gs = t / (rs + NB_EPS);
    }
    b -= pi;
    float mx = fmaxf(a, b);
    float ea = expf(a - mx), eb = expf(b - mx);
    float se = ea + eb;
    float lse = mx + logf(se);
    NbFwd o;
    o.ll = lse - softplusf(-pi);
    float wa = ea / se, wb = eb / se;
    o.ep = wa * (gp - (th + t) / d1) * rp;  // d ll / d rho_p * rho_p
    o.es = wb * (gs - (th + t) / d2) * rs;
    return o;
}

struct NbBwd { float dyp, dys, dpi, dth; };

// scale = upstream d loss / d ll  (= -grad_scale / B)
__device__ __forceinline__ NbBwd nb_backward(float t, float lp, float ls, float pi, float th, float lte, float dgt, float Rp,
                                             float Rs, float Dp, float Ds, float inv_elib, float scale) {
    float rp = expf(lp + Rp), rs = expf(ls + Rs);
    float d1 = th + rp + NB_EPS, d2 = th + rs + NB_EPS;
    float l1 = logf(d1), l2 = logf(d2);
    float diff = th * (l2 - l1) + pi;  // log_nb_p - (log_nb_s - pi); the lgamma terms cancel
    float gp = 0.0f, gs = 0.0f, dg = 0.0f;
    if (t != 0.0f) {
        diff += t * (logf(rp + NB_EPS) - l1 - logf(rs + NB_EPS) + l2);
        gp = t / (rp + NB_EPS);
        gs = t / (rs + NB_EPS);
        dg = digammaf_pos(t + th) - dgt;
    }
    float wa = 1.0f / (1.0f + expf(-diff));
    float wb = 1.0f / (1.0f + expf(diff));
    float q1 = (th + t) / d1, q2 = (th + t) / d2;
    float ep = wa * (gp - q1) * rp;
    float es = wb * (gs - q2) * rs;
    NbBwd o;
    o.dyp = scale * (ep - rp * inv_elib * Dp);  // softmax backward: rho (g - D / exp(lib))
    o.dys = scale * (es - rs * inv_elib * Ds);
    o.dpi = scale * (1.0f / (1.0f + expf(pi)) - wb);
    o.dth = scale * (wa * (lte - l1 - q1) + wb * (lte - l2 - q2) + th / (th + NB_EPS) + dg);
    return o;
}

// ---------------------------------------------------------------------------------------
// tile sweeps over the [B, G] index space
// ---------------------------------------------------------------------------------------
struct DecP {
    const void* X; long ldx; const int* rows;           // counts (row-gathered minibatch)
    const float* amix; long ld_amix;                     // [B, HD + P + S] = [hm | z_private_arg | z_shared_arg]
    const float* wfold;                                  // [G, P + S]
    const float* wm; const float* bm;                    // [G, HD + P + S], [G]
    const float* genec;                                  // [GC_N, G]
    const float* lib;                                    // [B]
    float* part_stats;                                   // [nTG, B, 4]
    float* rowc;                                         // [B, 4] = Rp, Rs, Dp, Ds
    float* pi;                                           // [B, G]
    float* part_nb;                                      // [nTG, B, 3]
    float* dyp; float* dys; float* dpi;                  // [B, G]
    float* colpart;                                      // [nTB, 4, G]
    int B, G, HD, P, S;
    float scale;
    int pi_ready;                                        // PASS_NB: pi already holds the mixture logits (tensor-core GEMM)
    __nv_bfloat16* dpi_bf16; long ld_dpi_bf16;           // PASS_BWD: write d pi as bf16 (operand of the tensor-core GEMMs)
};

enum { PASS_STATS = 0, PASS_NB = 1, PASS_BWD = 2 };

template <int PASS, int SRC>
__global__ void __launch_bounds__(GT_THREADS) dec_tile_kernel(DecP p) {
    __shared__ GemmSmem sm;
    __shared__ float red[16][GT_BN + 1];
    const int n0 = blockIdx.x * GT_BN, m0 = blockIdx.y * GT_BM;
    const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
    const int KMIX = p.HD + p.P + p.S, KZ = p.P + p.S;
    const int G = p.G, B = p.B;
    float lp[4][4], ls[4][4], pi[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) lp[i][j] = ls[i][j] = pi[i][j] = 0.0f;
    const float* azz = p.amix + p.HD;
    tile_mainloop<SPV_SRC_F32, false, SPV_SRC_F32, true>(lp, azz, p.ld_amix, nullptr, p.wfold, KZ, nullptr, B, G, 0, p.P, m0, n0, sm);
    tile_mainloop<SPV_SRC_F32, false, SPV_SRC_F32, true>(ls, azz + p.P, p.ld_amix, nullptr, p.wfold + p.P, KZ, nullptr, B, G, 0, p.S, m0, n0, sm);
    if (PASS == PASS_NB && !p.pi_ready)
        tile_mainloop<SPV_SRC_F32, false, SPV_SRC_F32, true>(pi, p.amix, p.ld_amix, nullptr, p.wm, KMIX, nullptr, B, G, 0, KMIX, m0, n0, sm);

    // per-gene constants of this thread's 4 genes
    float cp[4], cs[4], th[4], lte[4], lgx[4], bm[4];
    bool nok[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        int n = n0 + tx * 4 + j;
        nok[j] = n < G;
        int nn = nok[j] ? n : 0;
        cp[j] = p.genec[GC_CP * (long)G + nn];
        cs[j] = p.genec[GC_CS * (long)G + nn];
        if (PASS != PASS_STATS) {
            th[j] = p.genec[GC_THETA * (long)G + nn];
            lte[j] = p.genec[GC_LTE * (long)G + nn];
            lgx[j] = p.genec[(PASS == PASS_NB ? GC_LGT : GC_DGT) * (long)G + nn];
            bm[j] = PASS == PASS_NB ? p.bm[nn] : 0.0f;
        }
    }

    if (PASS == PASS_STATS) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            int m = m0 + ty * 4 + i;
            float mp = -INFINITY, ms = -INFINITY;
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (nok[j]) {
                    lp[i][j] += cp[j];
                    ls[i][j] += cs[j];
                    mp = fmaxf(mp, lp[i][j]);
                    ms = fmaxf(ms, ls[i][j]);
                }
            mp = half_warp_max(mp);
            ms = half_warp_max(ms);
            float sp = 0.0f, ss = 0.0f;
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (nok[j]) {
                    sp += expf(lp[i][j] - mp);
                    ss += expf(ls[i][j] - ms);
                }
            sp = half_warp_sum(sp);
            ss = half_warp_sum(ss);
            if (tx == 0 && m < B) {
                float* o = p.part_stats + ((long)blockIdx.x * B + m) * 4;
                o[0] = mp; o[1] = sp; o[2] = ms; o[3] = ss;
            }
        }
        return;
    }

    float csum[4][4];  // [quantity][gene] column partial sums (PASS_BWD)
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int j = 0; j < 4; ++j) csum[q][j] = 0.0f;

#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int m = m0 + ty * 4 + i;
        const bool mok = m < B;
        const int mm = mok ? m : 0;
        const float Rp = p.rowc[(long)mm * 4 + 0], Rs = p.rowc[(long)mm * 4 + 1];
        const long xr = (p.rows ? (long)p.rows[mm] : (long)mm) * p.ldx;
        if (PASS == PASS_NB) {
            float sll = 0.0f, sep = 0.0f, ses = 0.0f;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                int n = n0 + tx * 4 + j;
                if (mok && nok[j]) {
                    float t = load_src<SRC>(p.X, xr + n);
                    float piv = p.pi_ready ? p.pi[(long)m * G + n] : pi[i][j] + bm[j];
                    NbFwd o = nb_forward(t, lp[i][j] + cp[j], ls[i][j] + cs[j], piv, th[j], lte[j], lgx[j], Rp, Rs);
                    sll += o.ll; sep += o.ep; ses += o.es;
                    if (!p.pi_ready) p.pi[(long)m * G + n] = piv;
                }
            }
            sll = half_warp_sum(sll);
            sep = half_warp_sum(sep);
            ses = half_warp_sum(ses);
            if (tx == 0 && mok) {
                float* o = p.part_nb + ((long)blockIdx.x * B + m) * 3;
                o[0] = sll; o[1] = sep; o[2] = ses;
            }
        } else {  // PASS_BWD
            const float Dp = p.rowc[(long)mm * 4 + 2], Ds = p.rowc[(long)mm * 4 + 3];
            const float inv_elib = expf(-p.lib[mm]);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                int n = n0 + tx * 4 + j;
                if (mok && nok[j]) {
                    float t = load_src<SRC>(p.X, xr + n);
                    float piv = p.pi[(long)m * G + n];
                    NbBwd o = nb_backward(t, lp[i][j] + cp[j], ls[i][j] + cs[j], piv, th[j], lte[j], lgx[j], Rp, Rs, Dp, Ds,
                                          inv_elib, p.scale);
                    p.dyp[(long)m * G + n] = o.dyp;
                    p.dys[(long)m * G + n] = o.dys;
                    if (p.dpi_bf16) p.dpi_bf16[(long)m * p.ld_dpi_bf16 + n] = __float2bfloat16(o.dpi);
                    else p.dpi[(long)m * G + n] = o.dpi;
                    csum[0][j] += o.dyp; csum[1][j] += o.dys; csum[2][j] += o.dpi; csum[3][j] += o.dth;
                }
            }
        }
    }
    if (PASS == PASS_BWD) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            __syncthreads();
#pragma unroll
            for (int j = 0; j < 4; ++j) red[ty][tx * 4 + j] = csum[q][j];
            __syncthreads();
            if (threadIdx.x < GT_BN) {
                int n = n0 + threadIdx.x;
                float s = 0.0f;
#pragma unroll
                for (int r = 0; r < 16; ++r) s += red[r][threadIdx.x];
                if (n < G) p.colpart[((long)blockIdx.y * 4 + q) * G + n] = s;
            }
        }
    }
}

// combine the per-gene-tile softmax partials: Rp = lib - logsumexp_g(y_p), Rs likewise
__global__ void rowstat_kernel(const float* __restrict__ part, int nTG, int B, const float* __restrict__ lib,
                               float* __restrict__ rowc) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    float Mp = -INFINITY, Ms = -INFINITY;
    for (int t = 0; t < nTG; ++t) {
        const float* o = part + ((long)t * B + b) * 4;
        Mp = fmaxf(Mp, o[0]);
        Ms = fmaxf(Ms, o[2]);
    }
    float Sp = 0.0f, Ss = 0.0f;
    for (int t = 0; t < nTG; ++t) {
        const float* o = part + ((long)t * B + b) * 4;
        Sp += o[1] * expf(o[0] - Mp);
        Ss += o[3] * expf(o[2] - Ms);
    }
    float l = lib[b];
    rowc[(long)b * 4 + 0] = l - (Mp + logf(Sp));
    rowc[(long)b * 4 + 1] = l - (Ms + logf(Ss));
}

__global__ void rownb_kernel(const float* __restrict__ part, int nTG, int B, float* __restrict__ rowc, float* __restrict__ rec) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    float ll = 0.0f, dp = 0.0f, ds = 0.0f;
    for (int t = 0; t < nTG; ++t) {
        const float* o = part + ((long)t * B + b) * 3;
        ll += o[0]; dp += o[1]; ds += o[2];
    }
    rec[b] = -ll;  // reference :823-824
    rowc[(long)b * 4 + 2] = dp;
    rowc[(long)b * 4 + 3] = ds;
}

static void fill_decp(DecP& p, const void* const* ptrs, long long ldx, long long ld_amix, int B, int G, int HD, int P, int S,
                      float scale) {
    p.X = ptrs[0]; p.rows = (const int*)ptrs[1]; p.amix = (const float*)ptrs[2]; p.wfold = (const float*)ptrs[3];
    p.wm = (const float*)ptrs[4]; p.bm = (const float*)ptrs[5]; p.genec = (const float*)ptrs[6]; p.lib = (const float*)ptrs[7];
    p.part_stats = (float*)ptrs[8]; p.rowc = (float*)ptrs[9]; p.pi = (float*)ptrs[10]; p.part_nb = (float*)ptrs[11];
    p.dyp = (float*)ptrs[12]; p.dys = (float*)ptrs[13]; p.dpi = (float*)ptrs[14]; p.colpart = (float*)ptrs[15];
    p.ldx = ldx; p.ld_amix = ld_amix; p.B = B; p.G = G; p.HD = HD; p.P = P; p.S = S; p.scale = scale;
    p.pi_ready = 0; p.dpi_bf16 = nullptr; p.ld_dpi_bf16 = 0;
}

// ptrs (SPV_DEC_NPTR = 17): X, rows, amix, wfold, wm, bm, genec, lib, part_stats, rowc, pi, part_nb, dyp, dys, dpi,
// colpart, rec.   Forward: rec[b], rowc and pi are produced.
extern "C" int spv_dec_nb_fwd(int src, const void* const* ptrs, long long ldx, long long ld_amix, int B, int G, int HD, int P,
                              int S, int phases, void* stream) {
    if (!ptrs || (phases & 3) == 0 || B <= 0 || G <= 0 || HD < 0 || P <= 0 || S <= 0) return SPV_ERR_ARG;
    const int need[] = {0, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 16};
    for (int i : need)
        if (!ptrs[i]) return SPV_ERR_ARG;
    DecP p;
    fill_decp(p, ptrs, ldx, ld_amix, B, G, HD, P, S, 0.0f);
    p.pi_ready = (phases & 4) ? 1 : 0;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    dim3 grid((G + GT_BN - 1) / GT_BN, (B + GT_BM - 1) / GT_BM);
    const int nTG = grid.x;
    if (src != SPV_SRC_U16_LOG1P && src != SPV_SRC_F32_LOG1P) return SPV_ERR_ARG;
    if (phases & 1) {  // gene-axis softmax normalisers
        if (src == SPV_SRC_U16_LOG1P) dec_tile_kernel<PASS_STATS, SPV_SRC_U16_LOG1P><<<grid, GT_THREADS, 0, st>>>(p);
        else dec_tile_kernel<PASS_STATS, SPV_SRC_F32_LOG1P><<<grid, GT_THREADS, 0, st>>>(p);
        SPV_CHECK_LAUNCH();
        rowstat_kernel<<<(B + 127) / 128, 128, 0, st>>>(p.part_stats, nTG, B, p.lib, p.rowc);
        SPV_CHECK_LAUNCH();
    }
    if (phases & 2) {  // mixture GEMM + NB log-likelihood
        if (src == SPV_SRC_U16_LOG1P) dec_tile_kernel<PASS_NB, SPV_SRC_U16_LOG1P><<<grid, GT_THREADS, 0, st>>>(p);
        else dec_tile_kernel<PASS_NB, SPV_SRC_F32_LOG1P><<<grid, GT_THREADS, 0, st>>>(p);
        SPV_CHECK_LAUNCH();
        rownb_kernel<<<(B + 127) / 128, 128, 0, st>>>(p.part_nb, nTG, B, p.rowc, (float*)ptrs[16]);
        SPV_CHECK_LAUNCH();
    }
    return SPV_OK;
}

__global__ void colpart_reduce_kernel(const float* __restrict__ colpart, int nTB, int G, float* __restrict__ colsum) {
    long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (i >= 4L * G) return;
    float s = 0.0f;
    for (int t = 0; t < nTB; ++t) s += colpart[(long)t * 4 * G + i];
    colsum[i] = s;
}

// Backward sweep: writes dyp, dys, dpi [B, G] (gradients w.r.t. the two BatchNorm outputs and the mixture logits)
// and colsum [4, G] = column sums of dyp, dys, dpi and d loss / d theta.   scale = -grad_scale / B.
extern "C" int spv_dec_nb_bwd(int src, const void* const* ptrs, long long ldx, long long ld_amix, int B, int G, int HD, int P,
                              int S, float scale, float* colsum, void* dpi_bf16, long long ld_dpi_bf16, void* stream) {
    if (!ptrs || !colsum || B <= 0 || G <= 0 || HD < 0 || P <= 0 || S <= 0) return SPV_ERR_ARG;
    const int need[] = {0, 2, 3, 6, 7, 9, 10, 12, 13, 15};
    for (int i : need)
        if (!ptrs[i]) return SPV_ERR_ARG;
    if (!ptrs[14] && !dpi_bf16) return SPV_ERR_ARG;
    DecP p;
    fill_decp(p, ptrs, ldx, ld_amix, B, G, HD, P, S, scale);
    p.dpi_bf16 = reinterpret_cast<__nv_bfloat16*>(dpi_bf16);
    p.ld_dpi_bf16 = ld_dpi_bf16;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    dim3 grid((G + GT_BN - 1) / GT_BN, (B + GT_BM - 1) / GT_BM);
    if (src == SPV_SRC_U16_LOG1P) dec_tile_kernel<PASS_BWD, SPV_SRC_U16_LOG1P><<<grid, GT_THREADS, 0, st>>>(p);
    else if (src == SPV_SRC_F32_LOG1P) dec_tile_kernel<PASS_BWD, SPV_SRC_F32_LOG1P><<<grid, GT_THREADS, 0, st>>>(p);
    else return SPV_ERR_ARG;
    SPV_CHECK_LAUNCH();
    colpart_reduce_kernel<<<(4 * G + 255) / 256, 256, 0, st>>>(p.colpart, grid.y, G, colsum);
    SPV_CHECK_LAUNCH();
    return SPV_OK;
}

// ---------------------------------------------------------------------------------------
// per-gene backward of the two folded BatchNorm+Linear branches (closed form):
//   Q = dy^T z (from spv_gemm), sdy = colsum(dy):  S2 = (Q.W - sdy mean_u) invstd = dgamma,  dbeta = sdy,
//   dW = a (Q - sdy zbar - S2 invstd Cov W);  rows of the correction operands for d z:
//   wv[g, :] = a sdy / B * W[g, :],   wmx[g, :] = a S2 / B * invstd * W[g, :]
// also d px_r = theta * colsum(dtheta), d bm = colsum(dpi).
// ---------------------------------------------------------------------------------------
struct GeneBwdP {
    const float *Wp, *Ws, *Qp, *Qs, *genec, *colsum, *zmean, *zcov;
    float *dWp, *dWs, *dgp, *dbp, *dgs, *dbs, *dpx_r, *dbm, *wv, *wmx;
    int G, P, S, B;
};

__global__ void __launch_bounds__(256) gene_bwd_kernel(GeneBwdP p) {
    extern __shared__ float sh[];
    const int KZ = p.P + p.S;
    float* smean = sh;
    float* scov = sh + KZ;
    for (int k = threadIdx.x; k < KZ; k += blockDim.x) smean[k] = p.zmean[k];
    for (int i = threadIdx.x; i < KZ * KZ; i += blockDim.x) scov[i] = p.zcov[i];
    __syncthreads();
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= p.G) return;
    const long G = p.G;
    const float invB = 1.0f / (float)p.B;
#pragma unroll
    for (int br = 0; br < 2; ++br) {
        const int K = br == 0 ? p.P : p.S;
        const int off = br == 0 ? 0 : p.P;
        const float* W = (br == 0 ? p.Wp : p.Ws) + (long)g * K;
        const float* Q = (br == 0 ? p.Qp : p.Qs) + (long)g * K;
        float* dW = (br == 0 ? p.dWp : p.dWs) + (long)g * K;
        const float a = p.genec[(br == 0 ? GC_AP : GC_AS) * G + g];
        const float invstd = p.genec[(br == 0 ? GC_ISTD_P : GC_ISTD_S) * G + g];
        const float mean_u = p.genec[(br == 0 ? GC_MEAN_P : GC_MEAN_S) * G + g];
        const float sdy = p.colsum[(long)br * G + g];
        float qw = 0.0f;
        for (int k = 0; k < K; ++k) qw = fmaf(Q[k], W[k], qw);
        const float S2 = (qw - sdy * mean_u) * invstd;
        (br == 0 ? p.dgp : p.dgs)[g] = S2;
        (br == 0 ? p.dbp : p.dbs)[g] = sdy;
        const float cv = a * sdy * invB, cm = a * S2 * invB * invstd;
        for (int k = 0; k < K; ++k) {
            float cw = 0.0f;
            for (int l = 0; l < K; ++l) cw = fmaf(scov[(off + k) * KZ + off + l], W[l], cw);
            dW[k] = a * (Q[k] - sdy * smean[off + k] - S2 * invstd * cw);
            p.wv[(long)g * KZ + off + k] = cv * W[k];
            p.wmx[(long)g * KZ + off + k] = cm * W[k];
        }
    }
    p.dpx_r[g] = p.genec[GC_THETA * G + g] * p.colsum[3 * G + g];
    p.dbm[g] = p.colsum[2 * G + g];
}

// ptrs: Wp, Ws, Qp, Qs, genec, colsum, zmean, zcov, dWp, dWs, dgamma_p, dbeta_p, dgamma_s, dbeta_s, dpx_r, dbm, wv, wmx
extern "C" int spv_dec_gene_bwd(const void* const* ptrs, int B, int G, int P, int S, void* stream) {
    if (!ptrs || B <= 0 || G <= 0 || P <= 0 || S <= 0 || P + S > 96) return SPV_ERR_ARG;
    for (int i = 0; i < 18; ++i)
        if (!ptrs[i]) return SPV_ERR_ARG;
    GeneBwdP p;
    p.Wp = (const float*)ptrs[0]; p.Ws = (const float*)ptrs[1]; p.Qp = (const float*)ptrs[2]; p.Qs = (const float*)ptrs[3];
    p.genec = (const float*)ptrs[4]; p.colsum = (const float*)ptrs[5]; p.zmean = (const float*)ptrs[6];
    p.zcov = (const float*)ptrs[7]; p.dWp = (float*)ptrs[8]; p.dWs = (float*)ptrs[9]; p.dgp = (float*)ptrs[10];
    p.dbp = (float*)ptrs[11]; p.dgs = (float*)ptrs[12]; p.dbs = (float*)ptrs[13]; p.dpx_r = (float*)ptrs[14];
    p.dbm = (float*)ptrs[15]; p.wv = (float*)ptrs[16]; p.wmx = (float*)ptrs[17];
    p.G = G; p.P = P; p.S = S; p.B = B;
    const int KZ = P + S;
    size_t smem = (size_t)(KZ + KZ * KZ) * sizeof(float);
    if (smem > 48 * 1024) cudaFuncSetAttribute(gene_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    gene_bwd_kernel<<<(G + 255) / 256, 256, smem, reinterpret_cast<cudaStream_t>(stream)>>>(p);
    SPV_CHECK_LAUNCH();
    return SPV_OK;
}

// dzz[b, c] = dmix[b, c] + dzraw[b, c] - v1[c] - sum_l M[c, l] (zz[b, l] - zbar[l])   (l within c's branch block)
//   dmix: the zz columns of d Amix (mixture GEMM);  dzraw = dy W' (both softmax branches);
//   v1 = colsum over genes of wv;  M = wmx^T W (block diagonal: [P, P] and [S, S], stored in a [KZ, KZ] matrix)
__global__ void dzz_combine_kernel(const float* __restrict__ dmix, long ld_dmix, const float* __restrict__ dzraw,
                                   const float* __restrict__ v1, const float* __restrict__ M, const float* __restrict__ zz,
                                   long ld_zz, const float* __restrict__ zmean, float* __restrict__ dzz, int B, int P, int S) {
    const int KZ = P + S;
    long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (i >= (long)B * KZ) return;
    int c = (int)(i % KZ);
    long b = i / KZ;
    int lo = c < P ? 0 : P, hi = c < P ? P : KZ;
    float corr = 0.0f;
    for (int l = lo; l < hi; ++l) corr = fmaf(M[c * KZ + l], zz[b * ld_zz + l] - zmean[l], corr);
    dzz[b * KZ + c] = dmix[b * ld_dmix + c] + dzraw[b * KZ + c] - v1[c] - corr;
}

extern "C" int spv_dec_dzz_combine(const float* dmix, long long ld_dmix, const float* dzraw, const float* v1, const float* M,
                                   const float* zz, long long ld_zz, const float* zmean, float* dzz, int B, int P, int S,
                                   void* stream) {
    if (!dmix || !dzraw || !v1 || !M || !zz || !zmean || !dzz || B <= 0 || P <= 0 || S <= 0) return SPV_ERR_ARG;
    long total = (long)B * (P + S);
    dzz_combine_kernel<<<(int)((total + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        dmix, ld_dmix, dzraw, v1, M, zz, ld_zz, zmean, dzz, B, P, S);
    SPV_CHECK_LAUNCH();
    return SPV_OK;
}
