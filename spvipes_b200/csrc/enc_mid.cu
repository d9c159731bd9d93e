// Middle of the encoders in one launch per direction (reference nn/networks.py:119-125, Encoder.forward):
//   forward : h2 = dropout(relu(h1 W2^T + b2)),  r = h2 Whead^T + bhead          (fc2 -> ReLU -> dropout -> mu / logvar heads)
//   backward: dh2 = (dr Whead) * gate(h2),  dh1 = (dh2 W2) * gate(h1)  (+ bf16 copy of dh1 for the tensor-core dW1 GEMM)
// for both encoders (private, shared) of a group: blockIdx.y = encoder.  These layers are a few MFLOP on 512 rows; as
// separate GEMM / elementwise launches they cost 5-10 us each in dependent memory round trips.  Here a CTA owns 8 rows of
// one encoder, brings W2, the head weights and its activations into shared memory with cp.async (everything in flight at
// once, no registers), and keeps the intermediate (h2 resp. dh2) in shared memory between the two contractions.
// fp32 throughout.  Limits: n_hidden <= 128, n_hidden % 4 == 0, head width (2P resp. 2S) <= 128.
#include <cuda_bf16.h>
#include "common.cuh"
#include "../../include/spvipes_b200.h"

namespace {

constexpr int EM_ROWS = 8, EM_THREADS = 128;

__device__ __forceinline__ void cp16(float* dst_smem, const float* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_wait_all() {
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}
// rows x cols floats (cols % 4 == 0, 16-byte aligned rows) from global (row pitch ld) into shared (row pitch lds)
__device__ __forceinline__ void stage_rows(float* dst, int lds, const float* src, long ld, int rows, int cols, int rows_valid) {
    const int c4 = cols >> 2;
    for (int i = threadIdx.x; i < rows * c4; i += EM_THREADS) {
        const int r = i / c4, c = (i - r * c4) << 2;
        if (r < rows_valid) cp16(dst + r * lds + c, src + (long)r * ld + c);
        else *reinterpret_cast<float4*>(dst + r * lds + c) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

struct EncMidFwd {
    const float* h1; long ld_h1;        // [B, 2H]
    const float *W2, *b2;               // [2H, H], [2H]   (private rows then shared rows)
    const float *Whp, *Whs, *bhd;       // [2P, H], [2S, H], [2P + 2S]
    float* h2; long ld_h2;              // [B, 2H]
    float* r; long ld_r;                // [B, 2P + 2S]
    const float* drop_mask; long ld_mask;  // explicit multiplier [B, 2H] or null
    float drop_p; unsigned long long seed; unsigned int stream_id; const int* step;
    int B, H, P2, S2;
};

__global__ void __launch_bounds__(EM_THREADS) enc_mid_fwd_kernel(EncMidFwd p) {
    extern __shared__ __align__(16) float em_sh[];
    const int H = p.H, ldw = H + 4;
    const int e = blockIdx.y;                       // 0 private, 1 shared
    const int NH = e == 0 ? p.P2 : p.S2;
    const int m0 = blockIdx.x * EM_ROWS, rows_valid = min(EM_ROWS, p.B - m0);
    float* sW2 = em_sh;                             // [H][H + 4]
    float* sWh = sW2 + H * ldw;                     // [NH][H + 4]
    float* sx = sWh + 128 * ldw;                    // [8][H + 4]   h1 rows, then reused for h2
    float* sh2 = sx + EM_ROWS * ldw;                // [8][H + 4]
    stage_rows(sW2, ldw, p.W2 + (long)e * H * H, H, H, H, H);
    stage_rows(sWh, ldw, e == 0 ? p.Whp : p.Whs, H, NH, H, NH);
    stage_rows(sx, ldw, p.h1 + (long)m0 * p.ld_h1 + e * H, p.ld_h1, EM_ROWS, H, rows_valid);
    cp_wait_all();
    __syncthreads();
    const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;  // row ty, columns tx + 16 j
    const int m = m0 + ty;
    {   // fc2: up to 8 outputs per thread
        float acc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = 0.0f;
        const float* xr = sx + ty * ldw;
        for (int k = 0; k < H; k += 4) {
            const float4 x4 = *reinterpret_cast<const float4*>(xr + k);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                if (tx + 16 * j < H) {
                    const float4 w4 = *reinterpret_cast<const float4*>(sW2 + (tx + 16 * j) * ldw + k);
                    acc[j] = fmaf(x4.x, w4.x, fmaf(x4.y, w4.y, fmaf(x4.z, w4.z, fmaf(x4.w, w4.w, acc[j]))));
                }
            }
        }
        const unsigned int stp = p.step ? (unsigned int)*p.step : 0u;
        const float keep = 1.0f - p.drop_p, inv_keep = 1.0f / keep;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int n = tx + 16 * j;
            if (n < H) {
                float v = fmaxf(acc[j] + __ldg(p.b2 + e * H + n), 0.0f);
                const long col = (long)e * H + n;
                if (m < p.B) {
                    if (p.drop_mask) v *= __ldg(p.drop_mask + (long)m * p.ld_mask + col);
                    else if (p.drop_p > 0.0f)
                        v = philox_uniform(p.seed, p.stream_id, stp, (unsigned long long)((long)m * (2 * H) + col)) <= keep ? v * inv_keep : 0.0f;
                    p.h2[(long)m * p.ld_h2 + col] = v;
                }
                sh2[ty * ldw + n] = v;
            }
        }
    }
    __syncthreads();
    {   // heads: NH <= 128 outputs per row
        float acc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = 0.0f;
        const float* xr = sh2 + ty * ldw;
        for (int k = 0; k < H; k += 4) {
            const float4 x4 = *reinterpret_cast<const float4*>(xr + k);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                if (tx + 16 * j < NH) {
                    const float4 w4 = *reinterpret_cast<const float4*>(sWh + (tx + 16 * j) * ldw + k);
                    acc[j] = fmaf(x4.x, w4.x, fmaf(x4.y, w4.y, fmaf(x4.z, w4.z, fmaf(x4.w, w4.w, acc[j]))));
                }
            }
        }
        const int c0 = e == 0 ? 0 : p.P2;
        if (m < p.B) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int n = tx + 16 * j;
                if (n < NH) p.r[(long)m * p.ld_r + c0 + n] = acc[j] + __ldg(p.bhd + c0 + n);
            }
        }
    }
}

struct EncMidBwd {
    const float* dr; long ld_dr;        // [B, 2P + 2S]
    const float *Whp, *Whs, *W2;
    const float* h2; long ld_h2;        // forward outputs (gates)
    const float* h1; long ld_h1;
    const float* drop_mask; long ld_mask; float drop_scale;  // dropout backward multiplier where h2 > 0
    float* dh2; long ld_dh2;            // [B, 2H]
    float* dh1; long ld_dh1;            // [B, 2H]
    __nv_bfloat16* dh1_bf16; long ld_dh1b;  // optional
    int B, H, P2, S2;
};

__global__ void __launch_bounds__(EM_THREADS) enc_mid_bwd_kernel(EncMidBwd p) {
    extern __shared__ __align__(16) float em_sh[];
    const int H = p.H, ldw = H + 4;
    const int e = blockIdx.y;
    const int NH = e == 0 ? p.P2 : p.S2, c0 = e == 0 ? 0 : p.P2;
    const int m0 = blockIdx.x * EM_ROWS;
    float* sW2 = em_sh;                             // [H][H + 4]   rows = fc2 outputs o, columns = inputs i
    float* sWh = sW2 + H * ldw;                     // [NH][H + 4]  rows = head outputs c, columns = fc2 outputs o
    float* sdr = sWh + 128 * ldw;                   // [8][132]
    float* sdh2 = sdr + EM_ROWS * 132;              // [8][H + 4]
    stage_rows(sW2, ldw, p.W2 + (long)e * H * H, H, H, H, H);
    stage_rows(sWh, ldw, e == 0 ? p.Whp : p.Whs, H, NH, H, NH);
    for (int i = threadIdx.x; i < EM_ROWS * NH; i += EM_THREADS) {  // dr tile (row pitch of dr is not 16-byte aligned in general)
        const int r = i / NH, c = i - r * NH;
        sdr[r * 132 + c] = m0 + r < p.B ? __ldg(p.dr + (long)(m0 + r) * p.ld_dr + c0 + c) : 0.0f;
    }
    cp_wait_all();
    __syncthreads();
    const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;  // row ty, columns 4 tx + {0..3} + 64 j
    const int m = m0 + ty;
    const bool mok = m < p.B;
    {   // dh2[m, o] = sum_c dr[m, c] Wh[c, o], gated by the ReLU / dropout of h2
        float acc[2][4];
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int q = 0; q < 4; ++q) acc[j][q] = 0.0f;
        for (int c = 0; c < NH; ++c) {
            const float d = sdr[ty * 132 + c];
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int o = 4 * tx + 64 * j;
                if (o < H) {
                    const float4 w4 = *reinterpret_cast<const float4*>(sWh + c * ldw + o);
                    acc[j][0] = fmaf(d, w4.x, acc[j][0]); acc[j][1] = fmaf(d, w4.y, acc[j][1]);
                    acc[j][2] = fmaf(d, w4.z, acc[j][2]); acc[j][3] = fmaf(d, w4.w, acc[j][3]);
                }
            }
        }
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int o = 4 * tx + 64 * j;
            if (o < H) {
                float v[4] = {acc[j][0], acc[j][1], acc[j][2], acc[j][3]};
                const long col = (long)e * H + o;
                if (mok) {
                    const float4 y = *reinterpret_cast<const float4*>(p.h2 + (long)m * p.ld_h2 + col);
                    const float yy[4] = {y.x, y.y, y.z, y.w};
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const float gm = p.drop_mask ? __ldg(p.drop_mask + (long)m * p.ld_mask + col + q) : p.drop_scale;
                        v[q] = yy[q] > 0.0f ? v[q] * gm : 0.0f;
                    }
                    *reinterpret_cast<float4*>(p.dh2 + (long)m * p.ld_dh2 + col) = make_float4(v[0], v[1], v[2], v[3]);
                } else {
                    v[0] = v[1] = v[2] = v[3] = 0.0f;
                }
                *reinterpret_cast<float4*>(sdh2 + ty * ldw + o) = make_float4(v[0], v[1], v[2], v[3]);
            }
        }
    }
    __syncthreads();
    {   // dh1[m, i] = sum_o dh2[m, o] W2[o, i], gated by the ReLU of h1
        float acc[2][4];
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int q = 0; q < 4; ++q) acc[j][q] = 0.0f;
        for (int o = 0; o < H; ++o) {
            const float d = sdh2[ty * ldw + o];
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int i = 4 * tx + 64 * j;
                if (i < H) {
                    const float4 w4 = *reinterpret_cast<const float4*>(sW2 + o * ldw + i);
                    acc[j][0] = fmaf(d, w4.x, acc[j][0]); acc[j][1] = fmaf(d, w4.y, acc[j][1]);
                    acc[j][2] = fmaf(d, w4.z, acc[j][2]); acc[j][3] = fmaf(d, w4.w, acc[j][3]);
                }
            }
        }
        if (mok) {
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int i = 4 * tx + 64 * j;
                if (i < H) {
                    const long col = (long)e * H + i;
                    const float4 y = *reinterpret_cast<const float4*>(p.h1 + (long)m * p.ld_h1 + col);
                    float4 v;
                    v.x = y.x > 0.0f ? acc[j][0] : 0.0f; v.y = y.y > 0.0f ? acc[j][1] : 0.0f;
                    v.z = y.z > 0.0f ? acc[j][2] : 0.0f; v.w = y.w > 0.0f ? acc[j][3] : 0.0f;
                    *reinterpret_cast<float4*>(p.dh1 + (long)m * p.ld_dh1 + col) = v;
                    if (p.dh1_bf16) {
                        __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
                        *reinterpret_cast<uint2*>(p.dh1_bf16 + (long)m * p.ld_dh1b + col) =
                            make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
                    }
                }
            }
        }
    }
}

size_t em_smem(int H) { return sizeof(float) * ((size_t)(H + 128) * (H + 4) + 2 * EM_ROWS * (H + 4 > 132 ? H + 4 : 132)); }
bool em_ok(int H, int P2, int S2) { return H > 0 && H <= 128 && (H & 3) == 0 && P2 > 0 && S2 > 0 && P2 <= 128 && S2 <= 128; }
bool aligned16(const void* p, long ld) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0 && (ld & 3) == 0; }

}  // namespace

// returns 1 when spv_enc_mid_fwd / _bwd support these sizes (else the caller uses the separate GEMM launches)
extern "C" int spv_enc_mid_supported(int H, int P, int S) { return em_ok(H, 2 * P, 2 * S) ? 1 : 0; }

extern "C" int spv_enc_mid_fwd(const float* h1, long long ld_h1, const float* W2, const float* b2, const float* Whp,
                               const float* Whs, const float* bhd, float* h2, long long ld_h2, float* r, long long ld_r,
                               const float* drop_mask, long long ld_mask, float drop_p, unsigned long long seed,
                               unsigned int stream_id, const int* step, int B, int H, int P, int S, void* stream) {
    if (!h1 || !W2 || !b2 || !Whp || !Whs || !bhd || !h2 || !r || B <= 0 || !em_ok(H, 2 * P, 2 * S)) return SPV_ERR_ARG;
    if (drop_p < 0.0f || drop_p >= 1.0f) return SPV_ERR_ARG;
    if (!aligned16(h1, ld_h1) || !aligned16(W2, H) || !aligned16(Whp, H) || !aligned16(Whs, H)) return SPV_ERR_ARG;
    EncMidFwd p;
    p.h1 = h1; p.ld_h1 = ld_h1; p.W2 = W2; p.b2 = b2; p.Whp = Whp; p.Whs = Whs; p.bhd = bhd; p.h2 = h2; p.ld_h2 = ld_h2;
    p.r = r; p.ld_r = ld_r; p.drop_mask = drop_mask; p.ld_mask = ld_mask; p.drop_p = drop_mask ? 0.0f : drop_p; p.seed = seed;
    p.stream_id = stream_id; p.step = step; p.B = B; p.H = H; p.P2 = 2 * P; p.S2 = 2 * S;
    const size_t smem = em_smem(H);
    static size_t configured[64] = {};  // per device: the attribute belongs to the function ON a device
    int dev = 0;
    cudaGetDevice(&dev);
    if (smem > configured[dev & 63]) {
        if (cudaFuncSetAttribute(enc_mid_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return SPV_ERR_LAUNCH;
        configured[dev & 63] = smem;
    }
    enc_mid_fwd_kernel<<<dim3((B + EM_ROWS - 1) / EM_ROWS, 2), EM_THREADS, smem, reinterpret_cast<cudaStream_t>(stream)>>>(p);
    SPV_CHECK_LAUNCH();
    return SPV_OK;
}

extern "C" int spv_enc_mid_bwd(const float* dr, long long ld_dr, const float* Whp, const float* Whs, const float* W2,
                               const float* h2, long long ld_h2, const float* h1, long long ld_h1, const float* drop_mask,
                               long long ld_mask, float drop_scale, float* dh2, long long ld_dh2, float* dh1,
                               long long ld_dh1, void* dh1_bf16, long long ld_dh1b, int B, int H, int P, int S, void* stream) {
    if (!dr || !Whp || !Whs || !W2 || !h2 || !h1 || !dh2 || !dh1 || B <= 0 || !em_ok(H, 2 * P, 2 * S)) return SPV_ERR_ARG;
    if (!aligned16(W2, H) || !aligned16(Whp, H) || !aligned16(Whs, H) || !aligned16(h2, ld_h2) || !aligned16(h1, ld_h1) ||
        !aligned16(dh2, ld_dh2) || !aligned16(dh1, ld_dh1))
        return SPV_ERR_ARG;
    if (dh1_bf16 && ((reinterpret_cast<uintptr_t>(dh1_bf16) & 7) || (ld_dh1b & 3))) return SPV_ERR_ARG;
    EncMidBwd p;
    p.dr = dr; p.ld_dr = ld_dr; p.Whp = Whp; p.Whs = Whs; p.W2 = W2; p.h2 = h2; p.ld_h2 = ld_h2; p.h1 = h1; p.ld_h1 = ld_h1;
    p.drop_mask = drop_mask; p.ld_mask = ld_mask; p.drop_scale = drop_scale; p.dh2 = dh2; p.ld_dh2 = ld_dh2; p.dh1 = dh1;
    p.ld_dh1 = ld_dh1; p.dh1_bf16 = reinterpret_cast<__nv_bfloat16*>(dh1_bf16); p.ld_dh1b = ld_dh1b;
    p.B = B; p.H = H; p.P2 = 2 * P; p.S2 = 2 * S;
    const size_t smem = em_smem(H);
    static size_t configured[64] = {};  // per device: the attribute belongs to the function ON a device
    int dev = 0;
    cudaGetDevice(&dev);
    if (smem > configured[dev & 63]) {
        if (cudaFuncSetAttribute(enc_mid_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return SPV_ERR_LAUNCH;
        configured[dev & 63] = smem;
    }
    enc_mid_bwd_kernel<<<dim3((B + EM_ROWS - 1) / EM_ROWS, 2), EM_THREADS, smem, reinterpret_cast<cudaStream_t>(stream)>>>(p);
    SPV_CHECK_LAUNCH();
    return SPV_OK;
}
