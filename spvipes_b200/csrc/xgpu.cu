// Gradient all-reduce of the data-parallel step as ONE kernel per parameter range, launched from inside the step's CUDA graph
// (one all-reduce of the gradients per step; the reference is single-device).
//
// The flat gradient buffer of every rank lives in symmetric memory (same offset on every GPU, mapped into every peer's address
// space over NVLink 5 / NVSwitch; allocated and exchanged by the host side, spvipes_b200/parallel.py).  Rank r owns slice r of
// the range.  Each CTA:
//   1. start handshake with the same-numbered CTA of every peer (release / acquire flags at system scope in a symmetric flag
//      buffer): a peer that answers has launched this kernel, i.e. its backward kernels - stream-ordered before it - are done;
//   2. reduces its part of the slice and writes the sums back into EVERY rank's buffer:
//        multicast (NVLS) path: multimem.ld_reduce (the switch adds the W copies) + multimem.st (the switch replicates);
//        peer path (no multicast object): W peer loads summed in rank order + W peer stores;
//   3. end handshake: once every peer's CTA b has signalled, all writes into this rank's part of chunk b have landed, and the
//      peers have finished reading this rank's gradients, so the next backward may overwrite them.
// Flags carry a monotonically increasing epoch (device counter per channel, advanced by the last CTA), so nothing is ever
// reset and a replayed CUDA graph needs no host-side argument update.  Every element is reduced by exactly one rank and
// broadcast, so all ranks hold bitwise identical sums.
#include "common.cuh"
#include "../../include/spvipes_b200.h"

#define XG_MAX_WORLD 16
#define XG_MAX_BLOCKS 64

struct XgPeers {
    float* buf[XG_MAX_WORLD];  // peer-mapped base address of every rank's gradient buffer (buf[rank] = the local one)
    int* flags[XG_MAX_WORLD];  // peer-mapped base address of every rank's flag buffer
};

__device__ __forceinline__ void flag_put(int* addr, int v) {
    asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ int flag_get(const int* addr) {
    int v;
    asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(addr) : "memory");
    return v;
}
__device__ __forceinline__ float4 mc_ld_reduce(const float* mc) {
    float4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(mc)
                 : "memory");
    return v;
}
__device__ __forceinline__ void mc_st(float* mc, float4 v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}
__device__ __forceinline__ float4 peer_ld(const float* p) {
    float4 v;
    asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void peer_st(float* p, float4 v) {
    asm volatile("st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// flag slot of (channel, CTA, source rank) in a rank's flag buffer
__device__ __forceinline__ int xg_slot(int channel, int cta, int src) { return (channel * XG_MAX_BLOCKS + cta) * XG_MAX_WORLD + src; }

// every thread of the CTA has finished its part (bar.sync), then thread k tells peer k (release, cumulative over the CTA's
// writes) and waits for peer k's matching signal
// A peer that does not answer within XG_TIMEOUT_NS (a rank that died or never launched the step) must not hang the GPU: the
// wait gives up, records the failure in *err (sticky, read by the host: spvipes_b200.parallel.NvlinkGradSync.check) and the
// kernel runs to completion with whatever the buffers hold.
#define XG_TIMEOUT_NS 20000000000ull
__device__ __forceinline__ unsigned long long xg_now() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void xg_handshake(const XgPeers& pr, int channel, int rank, int world, int epoch, int* err) {
    __syncthreads();
    if ((int)threadIdx.x < world) {
        const int k = threadIdx.x;
        flag_put(pr.flags[k] + xg_slot(channel, blockIdx.x, rank), epoch);
        const int* mine = pr.flags[rank] + xg_slot(channel, blockIdx.x, k);
        unsigned long long t0 = 0;
        unsigned int spins = 0;
        while (flag_get(mine) - epoch < 0) {
            if ((++spins & 1023u) == 0) {
                const unsigned long long t = xg_now();
                if (t0 == 0) t0 = t;
                else if (t - t0 > XG_TIMEOUT_NS) {
                    atomicExch(err, 1 + k);
                    break;
                }
            }
        }
    }
    __syncthreads();
}

template <bool MULTICAST>
__global__ void __launch_bounds__(512) xg_allreduce_kernel(const __grid_constant__ XgPeers pr, float* __restrict__ mc, long off4, long n4,
                                                           int rank, int world, int channel, int* epoch_ctr, int* ticket, int* err) {
    const int e0 = *epoch_ctr;  // stable during the launch: the last CTA advances it after everybody has read it
    xg_handshake(pr, channel, rank, world, e0 + 1, err);
    const long per = (n4 + world - 1) / world;
    const long lo = min(n4, (long)rank * per), hi = min(n4, lo + per);
    const long chunk = (hi - lo + gridDim.x - 1) / gridDim.x;
    const long c_lo = min(hi, lo + (long)blockIdx.x * chunk), c_hi = min(hi, c_lo + chunk);
    if (MULTICAST) {
        float* base = mc + 4 * off4;
        long i = c_lo + threadIdx.x;
        for (; i + 3 * (long)blockDim.x < c_hi; i += 4 * (long)blockDim.x) {  // four reductions in flight per thread
            float4 a = mc_ld_reduce(base + 4 * i), b = mc_ld_reduce(base + 4 * (i + blockDim.x));
            float4 c = mc_ld_reduce(base + 4 * (i + 2 * (long)blockDim.x)), d = mc_ld_reduce(base + 4 * (i + 3 * (long)blockDim.x));
            mc_st(base + 4 * i, a);
            mc_st(base + 4 * (i + blockDim.x), b);
            mc_st(base + 4 * (i + 2 * (long)blockDim.x), c);
            mc_st(base + 4 * (i + 3 * (long)blockDim.x), d);
        }
        for (; i < c_hi; i += blockDim.x) mc_st(base + 4 * i, mc_ld_reduce(base + 4 * i));
    } else {
        for (long i = c_lo + threadIdx.x; i < c_hi; i += blockDim.x) {
            float4 s = peer_ld(pr.buf[0] + 4 * (off4 + i));
            for (int k = 1; k < world; ++k) {  // fixed rank order: the sum does not depend on which rank computes it
                const float4 v = peer_ld(pr.buf[k] + 4 * (off4 + i));
                s.x += v.x, s.y += v.y, s.z += v.z, s.w += v.w;
            }
            for (int k = 0; k < world; ++k) peer_st(pr.buf[k] + 4 * (off4 + i), s);
        }
    }
    __threadfence_system();
    xg_handshake(pr, channel, rank, world, e0 + 2, err);
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(ticket, 1) == (int)gridDim.x - 1) {
            *epoch_ctr = e0 + 2;
            *ticket = 0;
        }
    }
}

// In-place sum over ranks of floats [offset, offset + n) of the symmetric gradient buffer.  peer_bufs / peer_flags: HOST arrays of
// `world` device addresses (every rank's buffer / flag buffer as mapped into this process; entry `rank` is the local one);
// mc_buf: this rank's multicast address of the same buffer, or NULL for the peer load/store path.  channel 0..3: launches that
// may be in flight at the same time (the decoder range beside the encoder backward) use different channels.  state: 12 ints on THIS
// device, zeroed once (per channel: epoch counter, ticket; state[8]: 1 + rank of a peer that timed out, 0 = healthy).  The flag buffer holds spv_xgpu_flag_ints() ints, zeroed on
// every rank before the first launch anywhere.  offset and n multiples of 4, buffers 16-byte aligned.
extern "C" int spv_xgpu_allreduce(void* const* peer_bufs, void* const* peer_flags, void* mc_buf, long long offset, long long n,
                                  int rank, int world, int channel, int* state, int blocks, void* stream) {
    if (!peer_bufs || !peer_flags || !state || world < 1 || world > XG_MAX_WORLD || rank < 0 || rank >= world || channel < 0 ||
        channel >= 4 || n <= 0 || (n & 3) || (offset & 3) || offset < 0)
        return SPV_ERR_ARG;
    XgPeers pr;
    for (int k = 0; k < world; ++k) {
        if (!peer_bufs[k] || !peer_flags[k] || (reinterpret_cast<uintptr_t>(peer_bufs[k]) & 15)) return SPV_ERR_ARG;
        pr.buf[k] = reinterpret_cast<float*>(peer_bufs[k]);
        pr.flags[k] = reinterpret_cast<int*>(peer_flags[k]);
    }
    if (reinterpret_cast<uintptr_t>(mc_buf) & 15) return SPV_ERR_ARG;
    const long n4 = n / 4;
    const long per = (n4 + world - 1) / world;
    if (blocks <= 0) blocks = 32;
    if (blocks > XG_MAX_BLOCKS) blocks = XG_MAX_BLOCKS;
    const long want = (per + 4 * 512 - 1) / (4 * 512);
    if (blocks > want) blocks = (int)(want < 1 ? 1 : want);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (mc_buf)
        xg_allreduce_kernel<true><<<blocks, 512, 0, st>>>(pr, reinterpret_cast<float*>(mc_buf), offset / 4, n4, rank, world, channel,
                                                          state + 2 * channel, state + 2 * channel + 1, state + 8);
    else
        xg_allreduce_kernel<false><<<blocks, 512, 0, st>>>(pr, nullptr, offset / 4, n4, rank, world, channel, state + 2 * channel,
                                                           state + 2 * channel + 1, state + 8);
    SPV_CHECK_LAUNCH();
    return SPV_OK;
}

// number of ints of the symmetric flag buffer
extern "C" int spv_xgpu_flag_ints(void) { return 4 * XG_MAX_BLOCKS * XG_MAX_WORLD; }
