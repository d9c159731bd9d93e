// Element math of the NB-mixture likelihood for the tensor-core path (fast SFU intrinsics: ex2.approx / lg2.approx /
// rcp.approx; parity gate of this path is 1e-2 relative on the loss terms, BASELINE.json north_star).
// scvi-tools 0.20.0 log_mixture_nb, shared-theta branch, eps = 1e-8; reference call sites module/spVIPESmodule.py:759, 823-824.
#pragma once
#include "common.cuh"

// single-MUFU transcendental forms (flush-to-zero, no denormal fix-up code): operands here are never denormal
#if defined(PTC_EXP) && PTC_EXP == 2  // timing experiment only: no SFU instructions (wrong numerics)
__device__ __forceinline__ float fast_lg2(float x) { return fmaf(x, 0.25f, -1.0f); }
__device__ __forceinline__ float fast_ex2(float x) { return fmaf(x, 0.01f, 1.0f); }
__device__ __forceinline__ float fast_rcp(float x) { return fmaf(x, -0.01f, 1.0f); }
#else
__device__ __forceinline__ float fast_lg2(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float fast_ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float fast_rcp(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
#endif
__device__ __forceinline__ float fast_log(float x) { return fast_lg2(x) * 0.69314718056f; }
__device__ __forceinline__ float fast_exp(float x) { return fast_ex2(x * 1.44269504089f); }

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

// outputs of the element functions below
struct NbGrad { float dyp, dys, dpi, dth; };
struct NbOut { float ll, ep, es; };

// (t, lgamma(t + 1)) of a raw count computed directly (counts beyond the table: rare)
__device__ __forceinline__ float2 nb_count_terms_exact(uint32_t c) {
    const float t = c == 0u ? 0.0f : fast_log(1.0f + (float)c);
    return make_float2(t, lgamma_pos_fast(t + 1.0f));
}
// fills lut[0..255] = (log1p(c), lgamma(log1p(c) + 1)); call with 256 consecutive thread indices i
__device__ __forceinline__ void nb_fill_count_lut(float2* lut, int i) {
    const float t = i == 0 ? 0.0f : log1pf((float)i);
    lut[i] = make_float2(t, i == 0 ? 0.0f : lgammaf(t + 1.0f));
}

// ================================================================================================================
// v3 element math: everything in the base-2 domain with per-gene / per-row constants prepared by the caller.
// Instruction counts are what matters here (profiles/r1_nb_persistent_notes.md): 132 warp instructions per 32 elements in the
// forward kernel, 16 of them MUFU at 8 issue cycles each, ~171 cycles measured.  The rewrite minimised the FP32 work (~60
// forward, ~65 backward; was ~91), takes direct reciprocals, gets log(rho + eps) from the logit (rho = 2^x was just
// computed; the plain lg2 form is the EXACT variant, used for the rare element with rho ~ eps) and gates nothing on t != 0:
// every t-dependent term vanishes at t = 0 by itself.  Algebra checked against scvi's formula in float64: max |err| 7e-7
// (Stirling at y >= 4).
//   per gene : cpl = cp log2e, csl = cs log2e, bm, th, thE = th + eps,
//              forward K0 = th log(th + eps) - lgamma(th) + 0.5 log(2 pi);  backward K1 = log(th + eps) + th / (th + eps) - digamma(th)
//   per row  : Rpl = Rp log2e, Rsl = Rs log2e;  backward DpI = exp(-lib) Dp, DsI = exp(-lib) Ds
//   per count: t = log1p(c), lgt1 = lgamma(t + 1)
// ================================================================================================================
#define NB_LOG2E 1.4426950408889634f
#define NB_LN2 0.6931471805599453f

struct NbGene { float cpl, csl, bm, th, thE, K; };

// v4 (r2): the same algebra with 11 (forward) / 10 (backward) SFU operations per element instead of 16 / 15.  At the C5 shape
// the forward sweep ran at 61 % of the SFU pipe's peak (bench.py roofline.other), so SFU operations are what to cut:
//   * independent reciprocals share one rcp:  1/d1, 1/d2, 1/y from rcp(d1 d2 y);  1/o1, 1/o2 from rcp(o1 o2)  (2 extra FMULs per
//     shared factor; the products stay far inside fp32's range: d <= rho + theta, y = t + theta + 4, o <= 2)
//   * log2(rho + eps) = x (the logit, rho = 2^x) and rho / (rho + eps) = 1 unless rho < 1e-6, where the EXACT variant runs
//     (the first-order eps / rho term of v3 cost two reciprocals; it only matters for elements with a positive count under an
//     expected count below 1e-4, whose probability is of the order of that expected count).
template <bool EXACT>
__device__ __forceinline__ NbOut nb_forward_v3(float t, float lgt1, float accp, float accs, float accpi, const NbGene& g, float Rpl,
                                               float Rsl, bool& rare) {
    const float xp = fmaf(accp, NB_LOG2E, g.cpl + Rpl), xs = fmaf(accs, NB_LOG2E, g.csl + Rsl);
    const float pi = accpi + g.bm;
    const float rp = fast_ex2(xp), rs = fast_ex2(xs);
    const float d1 = rp + g.thE, d2 = rs + g.thE;
    // lgamma(x), x = t + th: shift by 4, P = x (x+1) (x+2) (x+3) = q (q + 2) with q = x (x + 3); Stirling at y = x + 4
    const float x = t + g.th;
    const float q = x * (x + 3.0f), P = q * (q + 2.0f), y = x + 4.0f;
    const float d12 = d1 * d2;
    const float r3 = fast_rcp(d12 * y);
    const float id1 = r3 * (d2 * y), id2 = r3 * (d1 * y), iy = r3 * d12, iy2 = iy * iy;
    const float L1 = fast_lg2(d1), L2 = fast_lg2(d2);
    float Lap, Las, fp, fs;  // log2(rho + eps) and rho / (rho + eps)
    if (EXACT) {
        const float ap = rp + NB_EPS, as = rs + NB_EPS;
        Lap = fast_lg2(ap); Las = fast_lg2(as);
        fp = rp * fast_rcp(ap); fs = rs * fast_rcp(as);
    } else {
        rare = rare || fminf(rp, rs) < 1e-6f;
        Lap = xp; Las = xs; fp = 1.0f; fs = 1.0f;
    }
    const float lgv = fmaf(fmaf(y - 0.5f, fast_lg2(y), -fast_lg2(P)), NB_LN2, fmaf(iy, fmaf(iy2, -0.0027777778f, 0.083333333f), -y));
    // a = K0 + lgv - lgt1 - ln2 m1,  b = K0 + lgv - lgt1 - ln2 m2 - pi,  m = x log2(th + rho + eps) - t log2(rho + eps)
    const float m1 = fmaf(-t, Lap, x * L1), m2 = fmaf(-t, Las, x * L2);
    const float df = fmaf(m2 - m1, NB_LN2, pi);  // a - b
    const float e = fast_ex2(-NB_LOG2E * fabsf(df)), epi = fast_ex2(-NB_LOG2E * fabsf(pi));
    const float o1 = 1.0f + e, o2 = 1.0f + epi;
    const float r2 = fast_rcp(o1 * o2);
    const float i1 = r2 * o2, i2 = r2 * o1;
    NbOut o;
    // logsumexp(a, b) - softplus(-pi) = a + max(0, -df) - max(-pi, 0) + log((1 + e) / (1 + epi))
    o.ll = (g.K + lgv - lgt1) + fmaf(m1, -NB_LN2, fmaxf(-df, 0.0f)) - fmaxf(-pi, 0.0f) + NB_LN2 * fast_lg2(o1 * i2);
    const float wmin = e * i1;
    const float wa = df >= 0.0f ? 1.0f - wmin : wmin, wb = 1.0f - wa;
    o.ep = wa * fmaf(t, fp, -(x * id1) * rp);
    o.es = wb * fmaf(t, fs, -(x * id2) * rs);
    return o;
}

template <bool EXACT>
__device__ __forceinline__ NbGrad nb_backward_v3(float t, float accp, float accs, float accpi, const NbGene& g, float Rpl, float Rsl,
                                                 float DpI, float DsI, float scale, bool& rare) {
    const float xp = fmaf(accp, NB_LOG2E, g.cpl + Rpl), xs = fmaf(accs, NB_LOG2E, g.csl + Rsl);
    const float pi = accpi + g.bm;
    const float rp = fast_ex2(xp), rs = fast_ex2(xs);
    const float d1 = rp + g.thE, d2 = rs + g.thE;
    const float x = t + g.th;
    // digamma(x): psi(x) = psi(x + 4) - P'(x) / P(x), P = q (q + 2), q = x (x + 3), P' = (2 q + 2)(2 x + 3); series at y = x + 4
    const float q = x * (x + 3.0f), P = q * (q + 2.0f), y = x + 4.0f;
    const float ra = fast_rcp(d1 * d2), rb = fast_rcp(y * P);
    const float id1 = ra * d2, id2 = ra * d1, iy = rb * P, iP = rb * y, iy2 = iy * iy;
    const float L1 = fast_lg2(d1), L2 = fast_lg2(d2);
    float Lap, Las, fp, fs;
    if (EXACT) {
        const float ap = rp + NB_EPS, as = rs + NB_EPS;
        Lap = fast_lg2(ap); Las = fast_lg2(as);
        fp = rp * fast_rcp(ap); fs = rs * fast_rcp(as);
    } else {
        rare = rare || fminf(rp, rs) < 1e-6f;
        Lap = xp; Las = xs; fp = 1.0f; fs = 1.0f;
    }
    const float m1 = fmaf(-t, Lap, x * L1), m2 = fmaf(-t, Las, x * L2);
    const float df = fmaf(m2 - m1, NB_LN2, pi);
    const float num = fmaf(q, 2.0f, 2.0f) * fmaf(x, 2.0f, 3.0f);
    const float ser = iy2 * fmaf(iy2, fmaf(iy2, 0.003968254f, -0.0083333333f), 0.083333333f);
    const float psi = fmaf(fast_lg2(y), NB_LN2, fmaf(-0.5f, iy, -ser)) - num * iP;
    const float e = fast_ex2(-NB_LOG2E * fabsf(df)), epi = fast_ex2(-NB_LOG2E * fabsf(pi));
    const float o1 = 1.0f + e, o2 = 1.0f + epi;
    const float r2 = fast_rcp(o1 * o2);
    const float i1 = r2 * o2, i2 = r2 * o1;
    const float wmin = e * i1;
    const float wa = df >= 0.0f ? 1.0f - wmin : wmin, wb = 1.0f - wa;
    const float q1 = x * id1, q2 = x * id2;
    const float ep = wa * fmaf(t, fp, -q1 * rp), es = wb * fmaf(t, fs, -q2 * rs);
    const float sneg = (pi >= 0.0f ? epi : 1.0f) * i2;  // sigmoid(-pi)
    NbGrad o;
    o.dyp = scale * fmaf(-rp, DpI, ep);
    o.dys = scale * fmaf(-rs, DsI, es);
    o.dpi = scale * (sneg - wb);
    o.dth = scale * ((g.K + psi) - wa * fmaf(L1, NB_LN2, q1) - wb * fmaf(L2, NB_LN2, q2));
    return o;
}
