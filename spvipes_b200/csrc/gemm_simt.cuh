// fp32 SIMT tile GEMM used by the "fp32 mode" of the hot path (north_star: 1e-4 parity mode).
// One 64x64 output tile per CTA, 256 threads, 4x4 outputs per thread, BK = 16.
// The mainloop is a device function so the fused decoder kernels (decoder.cu) reuse it.
#pragma once
#include "common.cuh"

#define GT_BM 64
#define GT_BN 64
#define GT_BK 16
#define GT_LD 68  // padded smem row (floats); 272 B keeps float4 alignment
#define GT_THREADS 256

struct GemmSmem {
    float As[GT_BK][GT_LD];
    float Bs[GT_BK][GT_LD];
};

// C[m, n] += sum_{k in [kBegin, kEnd)} A(m, k) * B(k, n) for the tile at (m0, n0).
//   TA == false: A stored [M][K] (element (m,k) at A[row(m) * lda + k]),   rowsA indexes m
//   TA == true : A stored [K][M] (element (m,k) at A[row(k) * lda + m]),   rowsA indexes k
//   TB == true : B stored [N][K] (element (k,n) at B[row(n) * ldb + k])   ("NT", y = x W^T)
//   TB == false: B stored [K][N] (element (k,n) at B[row(k) * ldb + n])   ("NN")
template <int SRC_A, bool TA, int SRC_B, bool TB>
__device__ __forceinline__ void tile_mainloop(float (&acc)[4][4], const void* __restrict__ A, long lda,
                                              const int* __restrict__ rowsA, const void* __restrict__ B, long ldb,
                                              const int* __restrict__ rowsB, int M, int N, int kBegin, int kEnd, int m0,
                                              int n0, GemmSmem& sm) {
    const int t = threadIdx.x;
    const int ty = t >> 4, tx = t & 15;
    float ra[4], rb[4];

    auto fetch = [&](int k0) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            // ---- A ----
            int mm, kk;
            if (!TA) { kk = t & 15; mm = (t >> 4) + 16 * i; } else { mm = t & 63; kk = (t >> 6) + 4 * i; }
            int m = m0 + mm, k = k0 + kk;
            float v = 0.0f;
            if (m < M && k < kEnd) {
                long r = TA ? (rowsA ? (long)rowsA[k] : (long)k) : (rowsA ? (long)rowsA[m] : (long)m);
                long c = TA ? (long)m : (long)k;
                v = load_src<SRC_A>(A, r * lda + c);
            }
            ra[i] = v;
            // ---- B ----
            int nn;
            if (TB) { kk = t & 15; nn = (t >> 4) + 16 * i; } else { nn = t & 63; kk = (t >> 6) + 4 * i; }
            int n = n0 + nn;
            k = k0 + kk;
            v = 0.0f;
            if (n < N && k < kEnd) {
                long r = TB ? (rowsB ? (long)rowsB[n] : (long)n) : (rowsB ? (long)rowsB[k] : (long)k);
                long c = TB ? (long)k : (long)n;
                v = load_src<SRC_B>(B, r * ldb + c);
            }
            rb[i] = v;
        }
    };
    auto stash = [&]() {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            int mm, kk;
            if (!TA) { kk = t & 15; mm = (t >> 4) + 16 * i; } else { mm = t & 63; kk = (t >> 6) + 4 * i; }
            sm.As[kk][mm] = ra[i];
            int nn;
            if (TB) { kk = t & 15; nn = (t >> 4) + 16 * i; } else { nn = t & 63; kk = (t >> 6) + 4 * i; }
            sm.Bs[kk][nn] = rb[i];
        }
    };

    if (kBegin >= kEnd) return;
    fetch(kBegin);
    __syncthreads();  // protect smem reuse across successive mainloop calls
    stash();
    __syncthreads();
    for (int k0 = kBegin; k0 < kEnd; k0 += GT_BK) {
        const bool more = (k0 + GT_BK) < kEnd;
        if (more) fetch(k0 + GT_BK);
#pragma unroll
        for (int kk = 0; kk < GT_BK; ++kk) {
            float4 a4 = *reinterpret_cast<const float4*>(&sm.As[kk][ty * 4]);
            float4 b4 = *reinterpret_cast<const float4*>(&sm.Bs[kk][tx * 4]);
            float a[4] = {a4.x, a4.y, a4.z, a4.w};
            float b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
        if (more) {
            stash();
            __syncthreads();
        }
    }
}
