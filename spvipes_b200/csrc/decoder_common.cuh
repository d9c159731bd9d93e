// per-gene constant table shared by the decoder kernels: genec[GC_N][G] (SoA, stride G)
#pragma once
enum {
    GC_CP = 0,   // folded BatchNorm shift of the private branch: beta - mean * a
    GC_CS,       // ... shared branch
    GC_AP,       // folded scale gamma / sqrt(var + eps), private
    GC_AS,       // ... shared
    GC_ISTD_P,   // 1 / sqrt(var + eps)
    GC_ISTD_S,
    GC_MEAN_P,   // batch (or running) mean of z W^T
    GC_MEAN_S,
    GC_THETA,    // exp(px_r)
    GC_LTE,      // log(theta + eps)
    GC_LGT,      // lgamma(theta)
    GC_DGT,      // digamma(theta)
    // ready-made constants of the base-2 element math of the tensor-core likelihood kernels (nb_math.cuh, v3)
    GC_CPL,      // GC_CP * log2(e)
    GC_CSL,      // GC_CS * log2(e)
    GC_THE,      // theta + eps
    GC_K0,       // theta log(theta + eps) - lgamma(theta) + 0.5 log(2 pi)          (forward)
    GC_K1,       // log(theta + eps) + theta / (theta + eps) - digamma(theta)       (backward)
    // shifts of the CENTRED form y = W' (z - m) + c' used by the tensor-core kernels (m = batch mean of z when training, so
    // c' = beta exactly; m = 0 in eval mode), times log2(e)
    GC_CPLC,
    GC_CSLC,
    GC_N
};
