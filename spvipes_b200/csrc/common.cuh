// Shared device helpers for the spVIPES B200 hot path (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#define SPV_OK 0
#define SPV_ERR_ARG -1
#define SPV_ERR_LAUNCH -2
#define SPV_ERR_ARCH -3

// every kernel launch site goes through this: counts the launch (spv_launch_count) and checks the launch status
extern unsigned long long g_spv_launches;
#define SPV_CHECK_LAUNCH()                              \
    do {                                                \
        ++g_spv_launches;                               \
        cudaError_t e__ = cudaGetLastError();           \
        if (e__ != cudaSuccess) return SPV_ERR_LAUNCH;  \
    } while (0)

// how a matrix operand is stored / transformed on load
enum SpvSrc : int {
    SPV_SRC_F32 = 0,       // plain float
    SPV_SRC_U16_LOG1P = 1, // uint16 counts, element = log(1 + x)   (reference module/spVIPESmodule.py:432-433)
    SPV_SRC_F32_LOG1P = 2, // float counts,  element = log(1 + x)
};

// PoE variants (reference module/spVIPESmodule.py:484-509 dispatcher)
enum SpvPoeMode : int { SPV_POE_LABEL = 0, SPV_POE_PAIRED = 1, SPV_POE_CLUSTER = 2 };
#define SPV_PARTNER_PAD (-1)
#define SPV_PARTNER_ABSENT (-2)

#define NB_EPS 1e-8f

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
// sum over the 16 lanes of a half warp (lanes with the same lane>>4)
__device__ __forceinline__ float half_warp_sum(float v) {
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float half_warp_max(float v) {
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// First round (U elements per thread) of copying a small global array to shared memory, split into load() and store() so
// that the loads of several arrays are in flight before the first store; rest() copies what is beyond U * blockDim.
template <int U>
struct StageArr {
    float v[U];
    __device__ __forceinline__ void load(const float* __restrict__ src, int n) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int i = u * (int)blockDim.x + (int)threadIdx.x;
            v[u] = (src && i < n) ? __ldg(src + i) : 0.0f;
        }
    }
    __device__ __forceinline__ void store(float* dst, int n) const {
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int i = u * (int)blockDim.x + (int)threadIdx.x;
            if (i < n) dst[i] = v[u];
        }
    }
    static __device__ __forceinline__ void rest(float* dst, const float* __restrict__ src, int n) {
        for (int i = U * (int)blockDim.x + (int)threadIdx.x; i < n; i += (int)blockDim.x) dst[i] = src ? __ldg(src + i) : 0.0f;
    }
};

template <int SRC>
__device__ __forceinline__ float load_src(const void* p, long i) {
    if (SRC == SPV_SRC_F32) return __ldg(reinterpret_cast<const float*>(p) + i);
    if (SRC == SPV_SRC_U16_LOG1P) {
        unsigned short c = __ldg(reinterpret_cast<const unsigned short*>(p) + i);
        return c == 0 ? 0.0f : log1pf((float)c);
    }
    float c = __ldg(reinterpret_cast<const float*>(p) + i);
    return c == 0.0f ? 0.0f : logf(1.0f + c);  // torch.log(1 + x), reference :433
}

// raw load / transform split, so that a batch of loads can be issued before any transform code runs
template <int SRC>
__device__ __forceinline__ unsigned int load_raw(const void* p, long i) {
    if (SRC == SPV_SRC_U16_LOG1P) return (unsigned int)__ldg(reinterpret_cast<const unsigned short*>(p) + i);
    return __ldg(reinterpret_cast<const unsigned int*>(p) + i);
}
template <int SRC>
__device__ __forceinline__ float xform_raw(unsigned int r) {
    if (SRC == SPV_SRC_F32) return __uint_as_float(r);
    if (SRC == SPV_SRC_U16_LOG1P) return r == 0u ? 0.0f : log1pf((float)r);
    float c = __uint_as_float(r);
    return c == 0.0f ? 0.0f : logf(1.0f + c);
}

// ---------------------------------------------------------------------------------------
// Philox4x32-10 counter RNG (for in-kernel reparameterisation noise and dropout keep masks)
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
        uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
        ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
        key.x += W0;
        key.y += W1;
    }
    return ctr;
}
__device__ __forceinline__ float u32_to_unit(uint32_t u) {  // (0,1]
    return ((float)(u >> 8) + 1.0f) * (1.0f / 16777216.0f);
}
// one standard normal for (stream, index); Box-Muller on two Philox words
__device__ __forceinline__ float philox_normal(unsigned long long seed, uint32_t stream, uint32_t step, unsigned long long idx) {
    uint4 r = philox4x32_10(make_uint4((uint32_t)idx, (uint32_t)(idx >> 32), stream, step),
                            make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
    float u1 = u32_to_unit(r.x), u2 = u32_to_unit(r.y);
    return sqrtf(-2.0f * logf(u1)) * cospif(2.0f * u2);
}
__device__ __forceinline__ float philox_uniform(unsigned long long seed, uint32_t stream, uint32_t step, unsigned long long idx) {
    uint4 r = philox4x32_10(make_uint4((uint32_t)idx, (uint32_t)(idx >> 32), stream, step),
                            make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
    return u32_to_unit(r.x);
}

// digamma for x > 0 (fp32): recurrence up to x >= 6, then the asymptotic series
__device__ __forceinline__ float digammaf_pos(float x) {
    float r = 0.0f;
    while (x < 6.0f) {
        r -= 1.0f / x;
        x += 1.0f;
    }
    float ix = 1.0f / x, ix2 = ix * ix;
    float s = ix2 * (1.0f / 12.0f - ix2 * (1.0f / 120.0f - ix2 * (1.0f / 252.0f)));
    return r + logf(x) - 0.5f * ix - s;
}

// fp32 -> fp16 operand with saturation: a value beyond fp16's range (an exploding latent of an outlier cell: the model does not
// clamp its log-variances) stays finite, as it does in the fp32 / bf16 formats, instead of turning the step into NaN
#include <cuda_fp16.h>
__device__ __forceinline__ __half to_half_sat(float x) { return __float2half_rn(fminf(fmaxf(x, -65504.0f), 65504.0f)); }

// torch F.softplus(x) (beta 1, threshold 20)
__device__ __forceinline__ float softplusf(float x) { return x > 20.0f ? x : log1pf(expf(x)); }
