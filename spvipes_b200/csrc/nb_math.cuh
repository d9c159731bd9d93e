// Element math of the NB-mixture likelihood for the tensor-core path (fast SFU intrinsics: ex2.approx / lg2.approx /
// rcp.approx; parity gate of this path is 1e-2 relative on the loss terms, BASELINE.json north_star).
// scvi-tools 0.20.0 log_mixture_nb, shared-theta branch, eps = 1e-8; reference call sites module/spVIPESmodule.py:759, 823-824.
#pragma once
#include "common.cuh"

// single-MUFU transcendental forms (flush-to-zero, no denormal fix-up code): operands here are never denormal
__device__ __forceinline__ float fast_lg2(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float fast_ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float fast_rcp(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float fast_log(float x) { return fast_lg2(x) * 0.69314718056f; }
__device__ __forceinline__ float fast_exp(float x) { return fast_ex2(x * 1.44269504089f); }
__device__ __forceinline__ float fast_div(float a, float b) { return a * fast_rcp(b); }

// lgamma(x) for x > 0: shift by 8 (lgamma(x) = lgamma(x + 8) - log(x (x+1) ... (x+7))) then Stirling at y = x + 8 >= 8
// (truncation error < 1/(1260 y^5) < 3e-8).  Branch-free: 2 lg2 + 1 rcp on the SFU, the rest FMA.
__device__ __forceinline__ float lgamma_pos_fast(float x) {
    float p = x * (x + 1.0f);
    p *= (x + 2.0f) * (x + 3.0f);
    p *= (x + 4.0f) * (x + 5.0f);
    p *= (x + 6.0f) * (x + 7.0f);
    float y = x + 8.0f;
    float iy = fast_rcp(y);
    float s = iy * (0.083333333f - iy * iy * 0.0027777778f);
    return (y - 0.5f) * fast_log(y) - y + 0.91893853f + s - fast_log(p);
}

// digamma(x) for x > 0: psi(x) = psi(x + 6) - sum_{i<6} 1/(x+i), the sum as P'(x)/P(x) of P = prod (x+i) (one division),
// then the asymptotic series at y = x + 6 (truncation error < 1/(240 y^8)).
__device__ __forceinline__ float digamma_pos_fast(float x) {
    float pr = x, dp = 1.0f;
#pragma unroll
    for (int i = 1; i < 6; ++i) {
        float xi = x + (float)i;
        dp = fmaf(dp, xi, pr);
        pr *= xi;
    }
    float y = x + 6.0f;
    float iy = fast_rcp(y), iy2 = iy * iy;
    float s = iy2 * (0.083333333f - iy2 * (0.0083333333f - iy2 * 0.003968254f));
    return fast_log(y) - 0.5f * iy - s - fast_div(dp, pr);
}

struct NbGrad { float dyp, dys, dpi, dth; };

// gradients of loss w.r.t. the two softmax logits, the mixture logit and theta for one (cell, gene) element.
// Dp / Ds: this cell's row sums of d ll / d rho * rho (from the forward), inv_elib = exp(-lib), scale = d loss / d ll.
__device__ __forceinline__ NbGrad nb_backward_fast(float t, float lp, float ls, float pi, float th, float lte, float dgt, float Rp,
                                                   float Rs, float Dp, float Ds, float inv_elib, float scale) {
    float rp = fast_exp(lp + Rp), rs = fast_exp(ls + Rs);
    float d1 = th + rp + NB_EPS, d2 = th + rs + NB_EPS;
    float l1 = fast_log(d1), l2 = fast_log(d2);
    float diff = th * (l2 - l1) + pi;  // log_nb_p - (log_nb_s - pi); the lgamma terms cancel
    float gp = 0.0f, gs = 0.0f, dg = 0.0f;
    if (t != 0.0f) {
        diff += t * (fast_log(rp + NB_EPS) - l1 - fast_log(rs + NB_EPS) + l2);
        gp = fast_div(t, rp + NB_EPS);
        gs = fast_div(t, rs + NB_EPS);
        dg = digamma_pos_fast(t + th) - dgt;
    }
    float e = fast_exp(-fabsf(diff));
    float wmin = fast_div(e, 1.0f + e);
    float wa = diff >= 0.0f ? 1.0f - wmin : wmin, wb = 1.0f - wa;
    float q1 = fast_div(th + t, d1), q2 = fast_div(th + t, d2);
    float ep = wa * (gp - q1) * rp, es = wb * (gs - q2) * rs;
    float epi = fast_exp(-fabsf(pi));
    float sneg = fast_div(pi >= 0.0f ? epi : 1.0f, 1.0f + epi);  // sigmoid(-pi)
    NbGrad o;
    o.dyp = scale * (ep - rp * inv_elib * Dp);
    o.dys = scale * (es - rs * inv_elib * Ds);
    o.dpi = scale * (sneg - wb);
    o.dth = scale * (wa * (lte - l1 - q1) + wb * (lte - l2 - q2) + fast_div(th, th + NB_EPS) + dg);
    return o;
}

struct NbOut { float ll, ep, es; };

// t = log1p(count); lp / ls = BatchNorm'd softmax logits of the private / shared branch; Rp / Rs = lib - logsumexp_g(logits);
// th = exp(px_r), lte = log(th + eps), lgt = lgamma(th).   Returns log-likelihood and d ll / d rho * rho for both branches.
__device__ __forceinline__ NbOut nb_forward_fast(float t, float lp, float ls, float pi, float th, float lte, float lgt, float Rp,
                                                 float Rs) {
    float rp = fast_exp(lp + Rp), rs = fast_exp(ls + Rs);
    float d1 = th + rp + NB_EPS, d2 = th + rs + NB_EPS;
    float l1 = fast_log(d1), l2 = fast_log(d2);
    float a = th * (lte - l1), b = th * (lte - l2);
    float gp = 0.0f, gs = 0.0f;
    if (t != 0.0f) {
        float lg = lgamma_pos_fast(t + th) - lgt - lgamma_pos_fast(t + 1.0f);
        a += t * (fast_log(rp + NB_EPS) - l1) + lg;
        b += t * (fast_log(rs + NB_EPS) - l2) + lg;
        gp = fast_div(t, rp + NB_EPS);
        gs = fast_div(t, rs + NB_EPS);
    }
    b -= pi;
    // logsumexp(a, b) = max + log(1 + exp(-|a - b|));  softplus(-pi) = max(-pi, 0) + log(1 + exp(-|pi|))
    float df = a - b;
    float e = fast_exp(-fabsf(df));
    float lse = fmaxf(a, b) + fast_log(1.0f + e);
    float sp = fmaxf(-pi, 0.0f) + fast_log(1.0f + fast_exp(-fabsf(pi)));
    NbOut o;
    o.ll = lse - sp;
    float wmin = fast_div(e, 1.0f + e);  // weight of the smaller of (a, b)
    float wa = df >= 0.0f ? 1.0f - wmin : wmin;
    float wb = 1.0f - wa;
    float q1 = fast_div(th + t, d1), q2 = fast_div(th + t, d2);
    o.ep = wa * (gp - q1) * rp;
    o.es = wb * (gs - q2) * rs;
    return o;
}
