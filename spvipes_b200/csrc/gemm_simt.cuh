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
//
// The global loads of a k-tile are branch-free (clamped addresses + select) so that all eight of a thread's loads are in
// flight together; with per-element branches ptxas reuses load destination registers as address temporaries and the
// loads serialise (ncu: long-scoreboard stalls, ~7600 cycles per k-tile; profiles/r1_simt_gemm_serialised.md).
template <int SRC_A, bool TA, int SRC_B, bool TB>
__device__ __forceinline__ void tile_mainloop(float (&acc)[4][4], const void* __restrict__ A, long lda,
                                              const int* __restrict__ rowsA, const void* __restrict__ B, long ldb,
                                              const int* __restrict__ rowsB, int M, int N, int kBegin, int kEnd, int m0,
                                              int n0, GemmSmem& sm) {
    const int t = threadIdx.x;
    const int ty = t >> 4, tx = t & 15;
    if (kBegin >= kEnd) return;  // uniform over the CTA
    unsigned int ra[4], rb[4];  // raw bits; the log1p transform is applied when the tile is stored to shared memory
    bool va[4], vb[4];

    // fixed (k-independent) part of each element's address
    long fixA[4], fixB[4];
    bool okA[4], okB[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        if (!TA) {  // element (m = (t>>4)+16i, k = k0 + (t&15)): row fixed
            int m = m0 + (t >> 4) + 16 * i;
            okA[i] = m < M;
            int mc = okA[i] ? m : (M - 1);
            fixA[i] = (rowsA ? (long)__ldg(rowsA + mc) : (long)mc) * lda;
        } else {  // element (m = t&63, k = k0 + (t>>6)+4i): column fixed
            int m = m0 + (t & 63);
            okA[i] = m < M;
            fixA[i] = okA[i] ? m : (M - 1);
        }
        if (TB) {
            int n = n0 + (t >> 4) + 16 * i;
            okB[i] = n < N;
            int nc = okB[i] ? n : (N - 1);
            fixB[i] = (rowsB ? (long)__ldg(rowsB + nc) : (long)nc) * ldb;
        } else {
            int n = n0 + (t & 63);
            okB[i] = n < N;
            fixB[i] = okB[i] ? n : (N - 1);
        }
    }

    auto fetch = [&](int k0) {
        long offA[4], offB[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            int kA = k0 + (TA ? (t >> 6) + 4 * i : (t & 15));
            va[i] = okA[i] && kA < kEnd;
            int kc = kA < kEnd ? kA : (kEnd - 1);
            offA[i] = TA ? (rowsA ? (long)__ldg(rowsA + kc) : (long)kc) * lda + fixA[i] : fixA[i] + kc;
            int kB = k0 + (TB ? (t & 15) : (t >> 6) + 4 * i);
            vb[i] = okB[i] && kB < kEnd;
            kc = kB < kEnd ? kB : (kEnd - 1);
            offB[i] = TB ? fixB[i] + kc : (rowsB ? (long)__ldg(rowsB + kc) : (long)kc) * ldb + fixB[i];
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            ra[i] = load_raw<SRC_A>(A, offA[i]);
            rb[i] = load_raw<SRC_B>(B, offB[i]);
        }
    };
    auto stash = [&]() {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            int mm, kk;
            if (!TA) { kk = t & 15; mm = (t >> 4) + 16 * i; } else { mm = t & 63; kk = (t >> 6) + 4 * i; }
            sm.As[kk][mm] = va[i] ? xform_raw<SRC_A>(ra[i]) : 0.0f;
            int nn;
            if (TB) { kk = t & 15; nn = (t >> 4) + 16 * i; } else { nn = t & 63; kk = (t >> 6) + 4 * i; }
            sm.Bs[kk][nn] = vb[i] ? xform_raw<SRC_B>(rb[i]) : 0.0f;
        }
    };

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
