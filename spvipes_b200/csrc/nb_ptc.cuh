// Persistent variant of the fused decoder-GEMM + NB-likelihood kernels (see nb_ptc.cu): shared declarations.
#pragma once
#include "tc_common.cuh"

namespace ptc {

constexpr int BM = 128;        // cells per row tile (TMEM lanes)
constexpr int BN = 16;         // genes per unit: fine-grained so that a static split over the SMs is balanced to ~6 %
constexpr int BK = 64;         // bf16 elements per 128-byte swizzle row
constexpr int KB_MAX = 5;      // k-blocks of the resident A tile: K = HD + 64 <= 320
constexpr int MAX_CTAS = 148;

struct Range {
    int u0, u1;  // units [u0, u1) of this CTA; unit u = (row tile u / nG, gene tile u % nG)
};
__host__ __device__ inline Range cta_range(int cta, int n_ctas, int units) {
    Range r;
    r.u0 = (int)((long long)cta * units / n_ctas);
    r.u1 = (int)((long long)(cta + 1) * units / n_ctas);
    return r;
}

bool enabled();  // opt-in: SPV_NB_PERSISTENT=1

// floats of the row-partial buffer: [cta][segment 0/1][epilogue group 0/1][128 rows][3]
inline long long part_floats() { return (long long)MAX_CTAS * 2 * 2 * BM * 3; }

// the persistent kernels need: the whole K extent of a row tile resident in shared memory, at most two row tiles per CTA
inline bool eligible(int B, int G, int K, int n_ctas) {
    const int nTB = (B + BM - 1) / BM, nG = (G + BN - 1) / BN;
    const long long units = (long long)nTB * nG;
    const int per_cta = (int)((units + n_ctas - 1) / n_ctas);
    return enabled() && (K + BK - 1) / BK <= KB_MAX && per_cta <= nG && units < (1ll << 30);
}

int sm_count();  // cached cudaDevAttrMultiProcessorCount of the current device, clamped to MAX_CTAS

struct FwdParams {
    const void* X; long ldx; const int* rows;
    const float* bm;      // [G]
    const float* genec;   // [GC_N, G]
    const float* rowc;    // [B, 4]: Rp, Rs
    float* pi;            // [B, G] or null
    float* part;          // part_floats()
    long long* trace;     // diagnostic (spv_debug_trace): per CTA, group and unit four globaltimer stamps, or null
    int B, G, num_kb, kb_z, Gp, nG, units;
};
int fwd_launch(int src, const CUtensorMap& mapA, const CUtensorMap& mapB, const CUtensorMap& mapZ, const FwdParams& p,
               cudaStream_t st);
int fwd_rowreduce(const float* part, int G, int B, float* rowc, float* rec, cudaStream_t st);

struct BwdParams {
    const void* X; long ldx; const int* rows;
    const float* bm;
    const float* genec;
    const float* rowc;    // [B, 4]: Rp, Rs, Dp, Ds
    const float* lib;     // [B]
    __nv_bfloat16* d3;    // [B, ld_d3] = [dpi | dyp | dys], blocks of Gp columns
    long ld_d3;
    float* colpart;       // [nTB, 4, G]
    float scale;
    int B, G, num_kb, kb_z, Gp, nG, units;
};
int bwd_launch(int src, const CUtensorMap& mapA, const CUtensorMap& mapB, const CUtensorMap& mapZ, const BwdParams& p,
               cudaStream_t st);

}  // namespace ptc
