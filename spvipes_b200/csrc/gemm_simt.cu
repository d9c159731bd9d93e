// Generic fp32 SIMT GEMM (+ split-K reduce) behind spv_gemm.  Replaces the cuBLAS calls the
// reference makes through nn.Linear / autograd (reference nn/networks.py:119-125, 314-325).
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
            if (bias) v += bias[n];
            if (p.relu) v = fmaxf(v, 0.0f);
            float* c = C + (size_t)m * p.ldc + n;
            *c = p.accumulate ? (*c + v) : v;
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
        if (p.bias) v += p.bias[(size_t)b * p.sBias + n];
        if (p.relu) v = fmaxf(v, 0.0f);
        float* c = p.C + (size_t)b * p.sC + (size_t)m * p.ldc + n;
        *c = p.accumulate ? (*c + v) : v;
    }
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
    if (M <= 0 || N <= 0 || K < 0 || batch <= 0 || !A || !B || !C) return SPV_ERR_ARG;
    if (splits < 1) splits = 1;
    if (splits > 1 && !ws) return SPV_ERR_ARG;
    GemmParams p;
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
