// Generic fp32 SIMT GEMM (+ split-K reduce) behind spv_gemm.  Replaces the cuBLAS calls the
// reference makes through nn.Linear / autograd (reference nn/networks.py:119-125, 314-325).
#include <cuda_bf16.h>
#include "gemm_simt.cuh"
#include "../../include/spvipes_b200.h"

struct GemmParams {
    const void* A;
    const void* B;
    float* C;
    const float* bias;
    const int* rowsA;
    const int* rowsB;
    float* ws;
    long lda, ldb, ldc, sA, sB, sC, sBias;
    int M, N, K, batch, splits, kchunk, relu, accumulate;
    // optional fused epilogue (whole-K kernel only; spv_gemm_fused), applied after bias / ReLU, (m, n') with n' = batch * sC + n:
    const float* gate_y;     // multiply by (gate_y[m, n'] > 0 ? (gate_mask ? gate_mask[m, n'] : gate_scale) : 0): ReLU (+ dropout) backward
    const float* gate_mask;
    long ld_gate, ld_mask;
    float gate_scale;
    float drop_p;            // > 0 (or drop_mask): dropout forward with the Philox keep mask of spv_dropout (seed, stream id, *step,
    const float* drop_mask;  //   element index m * drop_ld + n') or an explicit multiplier matrix
    unsigned long long drop_seed;
    unsigned int drop_stream;
    const int* drop_step;
    long drop_ld;
    __nv_bfloat16* c_bf16;   // also store the result as bf16 at c_bf16[m * ld_cbf16 + n']
    __nv_bfloat16* c_bf16_lo;  // and (optional) its bf16 residual, same layout (split-operand tensor-core GEMM)
    long ld_cbf16;
};

template <int SRC_A, bool TA, int SRC_B, bool TB>
__global__ void __launch_bounds__(GT_THREADS) gemm_kernel(GemmParams p) {
    __shared__ GemmSmem sm;
    const int z = blockIdx.z;
    const int b = z / p.splits, sp = z % p.splits;
    const int m0 = blockIdx.y * GT_BM, n0 = blockIdx.x * GT_BN;
    const int kBegin = sp * p.kchunk;
    const int kEnd = min(p.K, kBegin + p.kchunk);
    const char* A = reinterpret_cast<const char*>(p.A) + (size_t)b * p.sA * (SRC_A == SPV_SRC_U16_LOG1P ? 2 : 4);
    const char* B = reinterpret_cast<const char*>(p.B) + (size_t)b * p.sB * (SRC_B == SPV_SRC_U16_LOG1P ? 2 : 4);
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;
    tile_mainloop<SRC_A, TA, SRC_B, TB>(acc, A, p.lda, p.rowsA, B, p.ldb, p.rowsB, p.M, p.N, kBegin, kEnd, m0, n0, sm);
    const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
    if (p.splits > 1) {
        float* ws = p.ws + (size_t)z * p.M * p.N;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            int m = m0 + ty * 4 + i;
            if (m >= p.M) continue;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                int n = n0 + tx * 4 + j;
                if (n < p.N) ws[(size_t)m * p.N + n] = acc[i][j];
            }
        }
        return;
    }
    float* C = p.C + (size_t)b * p.sC;
    const float* bias = p.bias ? p.bias + (size_t)b * p.sBias : nullptr;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int m = m0 + ty * 4 + i;
        if (m >= p.M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int n = n0 + tx * 4 + j;
            if (n >= p.N) continue;
            float v = acc[i][j];
            float* c = C + (size_t)m * p.ldc + n;
            if (p.accumulate == 2) v += *c;  // pre-activation addend already in C (covariate term of the first encoder layer)
            if (bias) v += bias[n];
            if (p.relu) v = fmaxf(v, 0.0f);
            *c = p.accumulate == 1 ? (*c + v) : v;
        }
    }
}

__global__ void splitk_reduce_kernel(GemmParams p) {
    long total = (long)p.batch * p.M * p.N;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        int n = (int)(i % p.N);
        long r = i / p.N;
        int m = (int)(r % p.M);
        int b = (int)(r / p.M);
        float v = 0.0f;
        for (int s = 0; s < p.splits; ++s) v += p.ws[((size_t)(b * p.splits + s) * p.M + m) * p.N + n];  // fixed order: deterministic
        float* c = p.C + (size_t)b * p.sC + (size_t)m * p.ldc + n;
        if (p.accumulate == 2) v += *c;
        if (p.bias) v += p.bias[(size_t)b * p.sBias + n];
        if (p.relu) v = fmaxf(v, 0.0f);
        *c = p.accumulate == 1 ? (*c + v) : v;
    }
}

// ---------------------------------------------------------------------------------------
// Whole-K variant for the small dense layers of the step (K <= 256: encoder fc2 and heads, the mixing net's hidden layer
// and their input gradients).  These are 512-row problems of a few MFLOP, so what matters is (a) how many SMs share
// the work and (b) how few dependent memory round trips a CTA makes: 8 x 64 output tiles (64 .. 256 CTAs), both operand
// tiles brought into shared memory with all global loads of a thread in flight together (at most two rounds), 128 threads,
// 1 x 4 outputs per thread.  fp32, A stored [M][K]; B stored [N][K] (TB) or [K][N].
// (ncu, r1: the 32 x 64-tile / 256-thread first version spent 2550 instructions per warp on 16 SMs, 9-12 us per launch.)
// ---------------------------------------------------------------------------------------
#define SK_BM 8
#define SK_BN 64
#define SK_MAXK 256
#define SK_THREADS 128
#define SK_ROUND 16  // float4 loads in flight per thread and round

// One round of staging of a [rows][cols4 * 4] float tile (zero filled outside rows_valid x cols_valid): U float4 per
// thread starting at element `base`, loaded into registers by load() and written to shared memory by store(), so that the
// caller can put the loads of several tiles in flight before the first store.
template <int U>
struct SkRound {
    float4 v[U];
    __device__ __forceinline__ void load(int base, const float* __restrict__ src, long ld, int rows, int rows_valid, int cols4,
                                         int cols_valid) {
        const bool vec = ((reinterpret_cast<uintptr_t>(src) & 15) == 0) && ((ld & 3) == 0);
        const int dq = SK_THREADS / cols4, dr = SK_THREADS - dq * cols4;  // (r, c4) advance per SK_THREADS elements
        const int i0 = base + (int)threadIdx.x;
        int rr = i0 / cols4, cc = i0 - rr * cols4;
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int c = cc * 4;
            const bool in = rr < rows_valid && rr < rows;
            const float* g = src + (long)(in ? rr : 0) * ld + c;
            float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
            if (vec && c + 4 <= cols_valid) {
                if (in) x = __ldg(reinterpret_cast<const float4*>(g));
            } else {
                if (in && c < cols_valid) x.x = __ldg(g);
                if (in && c + 1 < cols_valid) x.y = __ldg(g + 1);
                if (in && c + 2 < cols_valid) x.z = __ldg(g + 2);
                if (in && c + 3 < cols_valid) x.w = __ldg(g + 3);
            }
            v[u] = x;
            rr += dq; cc += dr;
            if (cc >= cols4) { cc -= cols4; ++rr; }
        }
    }
    __device__ __forceinline__ void store(int base, float* dst, int ldd, int rows, int cols4) const {
        const int dq = SK_THREADS / cols4, dr = SK_THREADS - dq * cols4;
        const int i0 = base + (int)threadIdx.x;
        int r = i0 / cols4, c4 = i0 - r * cols4;
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (r < rows) *reinterpret_cast<float4*>(dst + r * ldd + c4 * 4) = v[u];
            r += dq; c4 += dr;
            if (c4 >= cols4) { c4 -= cols4; ++r; }
        }
    }
};

// RPT rows per thread: 8 x 64 tiles (RPT 1, the 512-row problems this kernel was written for) or 32 x 64 tiles (RPT 4, minibatches of
// 1024+ rows: the weight tile is staged once per 32 rows instead of once per 8, and a thread's four rows share every weight
// load - at 2048 rows and K = 256 the 8-row form moved 147 MB through L2 for 0.5 GFLOP).  The accumulation order per output element
// is the same in both, so they agree bitwise.
template <bool TB, int RPT>
__global__ void __launch_bounds__(SK_THREADS) gemm_smallk_kernel(GemmParams p) {
    constexpr int BM = SK_BM * RPT;
    extern __shared__ __align__(16) float sk_smem[];
    const int K4 = (p.K + 3) / 4, Kp = K4 * 4;
    const int lda_s = Kp + 4;                   // A tile [BM][Kp + 4]
    const int ldb_s = TB ? Kp + 4 : SK_BN + 4;  // B tile [SK_BN][Kp + 4] (TB) or [Kp][SK_BN + 4]
    float* As = sk_smem;
    float* Bs = sk_smem + BM * lda_s;
    const int b = blockIdx.z;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * SK_BN;
    const int nvalid = min(SK_BN, p.N - n0);
    const float* A = reinterpret_cast<const float*>(p.A) + (size_t)b * p.sA + (size_t)m0 * p.lda;
    const float* B = reinterpret_cast<const float*>(p.B) + (size_t)b * p.sB;
    {
        // B tile: rows x cols4 float4;  A tile: SK_BM x K4 float4 (<= 4 per thread).  First round: A and B loads together.
        const float* Bsrc = TB ? B + (size_t)n0 * p.ldb : B + n0;
        const int b_rows = TB ? SK_BN : Kp, b_rows_valid = TB ? nvalid : p.K;
        const int b_cols4 = TB ? K4 : SK_BN / 4, b_cols_valid = TB ? p.K : nvalid;
        const int b_total = b_rows * b_cols4;
        SkRound<BM * SK_MAXK / 4 / SK_THREADS> ra;
        SkRound<SK_ROUND> rb;
        ra.load(0, A, p.lda, BM, min(BM, p.M - m0), K4, p.K);
        rb.load(0, Bsrc, p.ldb, b_rows, b_rows_valid, b_cols4, b_cols_valid);
        ra.store(0, As, lda_s, BM, K4);
        rb.store(0, Bs, ldb_s, b_rows, b_cols4);
        for (int base = SK_ROUND * SK_THREADS; base < b_total; base += SK_ROUND * SK_THREADS) {
            rb.load(base, Bsrc, p.ldb, b_rows, b_rows_valid, b_cols4, b_cols_valid);
            rb.store(base, Bs, ldb_s, b_rows, b_cols4);
        }
    }
    __syncthreads();
    const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;  // row ty; columns tx + 16 j (TB) or 4 tx + j
    float acc[RPT][4];
#pragma unroll
    for (int r = 0; r < RPT; ++r) acc[r][0] = acc[r][1] = acc[r][2] = acc[r][3] = 0.f;
    const float* a0 = As + ty * lda_s;  // rows ty + 8 r
#pragma unroll 2
    for (int k = 0; k < Kp; k += 4) {
        float wv[4][4];  // [kk][j]
        if (TB) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float4 w4 = *reinterpret_cast<const float4*>(Bs + (tx + 16 * j) * ldb_s + k);
                wv[0][j] = w4.x; wv[1][j] = w4.y; wv[2][j] = w4.z; wv[3][j] = w4.w;
            }
        } else {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                const float4 w4 = *reinterpret_cast<const float4*>(Bs + (k + kk) * ldb_s + tx * 4);
                wv[kk][0] = w4.x; wv[kk][1] = w4.y; wv[kk][2] = w4.z; wv[kk][3] = w4.w;
            }
        }
#pragma unroll
        for (int r = 0; r < RPT; ++r) {
            const float4 x0 = *reinterpret_cast<const float4*>(a0 + r * SK_BM * lda_s + k);
            const float xa[4] = {x0.x, x0.y, x0.z, x0.w};
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[r][j] = fmaf(xa[kk], wv[kk][j], acc[r][j]);
        }
    }
    float* C = p.C + (size_t)b * p.sC;
    const float* bias = p.bias ? p.bias + (size_t)b * p.sBias : nullptr;
#pragma unroll
    for (int r = 0; r < RPT; ++r) {
        const int m = m0 + ty + SK_BM * r;
        if (m >= p.M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + (TB ? tx + 16 * j : tx * 4 + j);
            if (n >= p.N) continue;
            float v = acc[r][j];
            float* c = C + (size_t)m * p.ldc + n;
            if (p.accumulate == 2) v += *c;
            if (bias) v += __ldg(bias + n);
            if (p.relu) v = fmaxf(v, 0.0f);
            if (p.accumulate == 1) v += *c;
            const long nn = (long)b * p.sC + n;  // column inside the full output matrix
            if (p.gate_y) {
                const float gm = p.gate_mask ? __ldg(p.gate_mask + (size_t)m * p.ld_mask + nn) : p.gate_scale;
                v = __ldg(p.gate_y + (size_t)m * p.ld_gate + nn) > 0.0f ? v * gm : 0.0f;
            }
            if (p.drop_mask) {
                v *= __ldg(p.drop_mask + (size_t)m * p.drop_ld + nn);
            } else if (p.drop_p > 0.0f) {
                const unsigned int stp = p.drop_step ? (unsigned int)*p.drop_step : 0u;
                const float keep = 1.0f - p.drop_p;
                v = philox_uniform(p.drop_seed, p.drop_stream, stp, (unsigned long long)((size_t)m * p.drop_ld + nn)) <= keep ? v * (1.0f / keep) : 0.0f;
            }
            *c = v;
            if (p.c_bf16) {
                const __nv_bfloat16 hi = __float2bfloat16(v);
                p.c_bf16[(size_t)m * p.ld_cbf16 + nn] = hi;
                if (p.c_bf16_lo) p.c_bf16_lo[(size_t)m * p.ld_cbf16 + nn] = __float2bfloat16(v - __bfloat162float(hi));
            }
        }
    }
}

template <bool TB, int RPT>
static int launch_smallk_rpt(const GemmParams& p, cudaStream_t st) {
    constexpr int BM = SK_BM * RPT;
    const int Kp = (p.K + 3) / 4 * 4;
    const size_t smem = sizeof(float) * ((size_t)BM * (Kp + 4) + (TB ? (size_t)SK_BN * (Kp + 4) : (size_t)Kp * (SK_BN + 4)));
    static size_t configured[64] = {};  // per device: the attribute belongs to the function ON a device
    int dev = 0;
    cudaGetDevice(&dev);
    if (smem > 48 * 1024 && smem > configured[dev & 63]) {
        if (cudaFuncSetAttribute(gemm_smallk_kernel<TB, RPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
            return SPV_ERR_LAUNCH;
        configured[dev & 63] = smem;
    }
    dim3 grid((p.N + SK_BN - 1) / SK_BN, (p.M + BM - 1) / BM, p.batch);
    gemm_smallk_kernel<TB, RPT><<<grid, SK_THREADS, smem, st>>>(p);
    SPV_CHECK_LAUNCH();
    return SPV_OK;
}

template <bool TB>
static int launch_smallk(const GemmParams& p, cudaStream_t st) {
    return p.M >= 1024 ? launch_smallk_rpt<TB, 4>(p, st) : launch_smallk_rpt<TB, 1>(p, st);
}

template <int SRC_A, bool TA, int SRC_B, bool TB>
static int launch(const GemmParams& p, cudaStream_t st) {
    dim3 grid((p.N + GT_BN - 1) / GT_BN, (p.M + GT_BM - 1) / GT_BM, p.batch * p.splits);
    gemm_kernel<SRC_A, TA, SRC_B, TB><<<grid, GT_THREADS, 0, st>>>(p);
    SPV_CHECK_LAUNCH();
    if (p.splits > 1) {
        long total = (long)p.batch * p.M * p.N;
        int blocks = (int)min((long)148 * 8, (total + 255) / 256);
        splitk_reduce_kernel<<<blocks, 256, 0, st>>>(p);
        SPV_CHECK_LAUNCH();
    }
    return SPV_OK;
}

extern "C" int spv_gemm(int srcA, int transA, int srcB, int transB, const void* A, long long lda, const int* rowsA,
                        const void* B, long long ldb, const int* rowsB, float* C, long long ldc, int M, int N, int K,
                        int batch, long long sA, long long sB, long long sC, const float* bias, long long sBias, int relu,
                        int accumulate, int splits, float* ws, void* stream) {
    return spv_gemm_fused(srcA, transA, srcB, transB, A, lda, rowsA, B, ldb, rowsB, C, ldc, M, N, K, batch, sA, sB, sC, bias, sBias,
                          relu, accumulate, splits, ws, nullptr, 0, nullptr, 0, 1.0f, 0.0f, nullptr, 0ull, 0u, nullptr, 0, nullptr, nullptr, 0,
                          stream);
}

// spv_gemm + fused epilogue stages (see GemmParams).  The fused stages need the whole-K kernel: fp32 operands, A not
// transposed, no row gather, K <= 256, splits == 1; otherwise SPV_ERR_ARG when any of them is requested.
extern "C" int spv_gemm_fused(int srcA, int transA, int srcB, int transB, const void* A, long long lda, const int* rowsA,
                              const void* B, long long ldb, const int* rowsB, float* C, long long ldc, int M, int N, int K,
                              int batch, long long sA, long long sB, long long sC, const float* bias, long long sBias, int relu,
                              int accumulate, int splits, float* ws, const float* gate_y, long long ld_gate,
                              const float* gate_mask, long long ld_mask, float gate_scale, float drop_p, const float* drop_mask,
                              unsigned long long drop_seed, unsigned int drop_stream, const int* drop_step, long long drop_ld,
                              void* c_bf16, void* c_bf16_lo, long long ld_cbf16, void* stream) {
    if (M <= 0 || N <= 0 || K < 0 || batch <= 0 || !A || !B || !C) return SPV_ERR_ARG;
    if (drop_p < 0.0f || drop_p >= 1.0f) return SPV_ERR_ARG;
    if (splits < 1) splits = 1;
    if (splits > 1 && !ws) return SPV_ERR_ARG;
    GemmParams p;
    p.gate_y = gate_y; p.gate_mask = gate_mask; p.ld_gate = ld_gate; p.ld_mask = ld_mask; p.gate_scale = gate_scale;
    p.drop_p = drop_p; p.drop_mask = drop_mask; p.drop_seed = drop_seed; p.drop_stream = drop_stream; p.drop_step = drop_step;
    p.drop_ld = drop_ld; p.c_bf16 = reinterpret_cast<__nv_bfloat16*>(c_bf16); p.ld_cbf16 = ld_cbf16;
    p.c_bf16_lo = reinterpret_cast<__nv_bfloat16*>(c_bf16_lo);
    const bool fused = gate_y || drop_mask || drop_p > 0.0f || c_bf16;
    const bool smallk_ok = srcA == SPV_SRC_F32 && srcB == SPV_SRC_F32 && !transA && !rowsA && !rowsB && splits == 1 && K > 0 && K <= SK_MAXK;
    if (fused && !smallk_ok) return SPV_ERR_ARG;
    p.A = A; p.B = B; p.C = C; p.bias = bias; p.rowsA = rowsA; p.rowsB = rowsB; p.ws = ws;
    p.lda = lda; p.ldb = ldb; p.ldc = ldc; p.sA = sA; p.sB = sB; p.sC = sC; p.sBias = sBias;
    p.M = M; p.N = N; p.K = K; p.batch = batch; p.relu = relu; p.accumulate = accumulate;
    int kchunk = (K + splits - 1) / splits;
    kchunk = ((kchunk + GT_BK - 1) / GT_BK) * GT_BK;
    if (kchunk < GT_BK) kchunk = GT_BK;
    splits = K > 0 ? (K + kchunk - 1) / kchunk : 1;
    p.splits = splits; p.kchunk = kchunk;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const bool ta = transA != 0, tb = transB != 0;
    if (srcA == SPV_SRC_F32 && srcB == SPV_SRC_F32 && !ta && !rowsA && !rowsB && splits == 1 && K > 0 && K <= SK_MAXK)
        return tb ? launch_smallk<true>(p, st) : launch_smallk<false>(p, st);
    if (srcA == SPV_SRC_F32 && srcB == SPV_SRC_F32) {
        if (!ta && tb) return launch<SPV_SRC_F32, false, SPV_SRC_F32, true>(p, st);
        if (!ta && !tb) return launch<SPV_SRC_F32, false, SPV_SRC_F32, false>(p, st);
        if (ta && !tb) return launch<SPV_SRC_F32, true, SPV_SRC_F32, false>(p, st);
        return SPV_ERR_ARG;
    }
    if (srcB == SPV_SRC_F32 && !ta && tb) {  // encoder fc1 forward: counts are the A operand
        if (srcA == SPV_SRC_U16_LOG1P) return launch<SPV_SRC_U16_LOG1P, false, SPV_SRC_F32, true>(p, st);
        if (srcA == SPV_SRC_F32_LOG1P) return launch<SPV_SRC_F32_LOG1P, false, SPV_SRC_F32, true>(p, st);
    }
    if (srcA == SPV_SRC_F32 && ta && !tb) {  // encoder fc1 weight gradient: counts are the B operand
        if (srcB == SPV_SRC_U16_LOG1P) return launch<SPV_SRC_F32, true, SPV_SRC_U16_LOG1P, false>(p, st);
        if (srcB == SPV_SRC_F32_LOG1P) return launch<SPV_SRC_F32, true, SPV_SRC_F32_LOG1P, false>(p, st);
    }
    return SPV_ERR_ARG;
}
