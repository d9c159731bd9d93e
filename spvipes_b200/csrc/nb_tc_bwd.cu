// Backward sweep of the fused decoder + NB-mixture likelihood on the tensor-core path.
// Same tiling and pipeline as nb_tc.cu: the three logit tiles (pi, lp, ls) are RECOMPUTED on tcgen05 from the bf16 operands
// (nothing [B, G]-sized was saved by the forward); the epilogue turns them into
//   D3T [3 Gp, B] fp16, gene-major: dpi (operand of the weight / input gradient GEMMs of the mixture layer), dyp, dys
//   (gradients w.r.t. the two folded BatchNorm outputs = softmax logits),
//   colpart [nTB, 4, G]: per-128-row-tile column sums of dyp, dys, dpi and d loss / d theta.
// Reference: autograd of nn/networks.py:314-325 + scvi log_mixture_nb (module/spVIPESmodule.py:823-824).
#include <cuda_fp16.h>
#include "tc_common.cuh"
#include "nb_math.cuh"
#include "decoder_common.cuh"
#include "../../include/spvipes_b200.h"

namespace {

constexpr int BM = 128, BN = 64, BK = 64, STAGES = 2;
constexpr int WCOLS = BN / 2;                 // gene columns per epilogue warp (two warps per TMEM lane quarter)
constexpr int GATHER_ROWS = 64 / BN;          // rows of the count tile one warp gathers per load: 32 lanes cover BN / 2 words each
constexpr int EPI_WARPS = 8, EPI_THREADS = 32 * EPI_WARPS;
constexpr int THREADS = 64 + EPI_THREADS;
constexpr int A_BYTES = BM * BK * 2, B_BYTES = BN * BK * 2, STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int ACC_COLS = BN, Z_COLS = 2 * BN;  // tensor memory is allocated in two steps (powers of two >= 32)
static_assert(ACC_COLS == 32 || ACC_COLS == 64, "tensor-memory allocations are powers of two");
constexpr int CNT_PITCH_W = BN / 2 + 2;  // 34 words per row: thread = row reads 64 bits (four codes), aligned and conflict-free
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 2 * B_BYTES + 1024 + BN * 16 + 4 * 4 * BN * 4 + BN * NB_TAB * 8 + 256;

static_assert(BM * CNT_PITCH_W * 4 <= STAGES * STAGE_BYTES, "count tile must fit in the operand stages");

struct NbTcBwdParams {
    const void* X; long ldx; const int* rows;
    const float* bm;
    const float* genec;
    const float* rowc;   // [B, 4]: Rp, Rs, Dp, Ds
    const float* lib;    // [B]
    const float2* tgb;   // [G, NB_TAB] backward count table (spv_dec_theta_tables)
    __nv_bfloat16* dpi; long ld_dpi;   // D3T [3 * Gp, ld_dpi >= B] = [dpi ; dyp ; dys], gene-major (fp16 values)
    float* colpart;                    // [nTB, 4, G]
    int B, G, K, Gp;
    float scale;
};

// row of the count tile that lane `lane` of epilogue warp e loads in its i-th gather: a warp covers GATHER_ROWS rows per load
__device__ __forceinline__ int cnt_row(int e, int lane, int i) {
    return (e + EPI_WARPS * i) * GATHER_ROWS + lane / (BN / 2);
}

// CTAs of this kernel currently resident per SM (a scheduling hint only, see the allocation below; balanced by every CTA)
__device__ int g_resident_bwd[256];

__device__ __forceinline__ unsigned long long mad_wide(uint32_t a, uint32_t b, unsigned long long c) {
    unsigned long long d;
    asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(d) : "r"(a), "r"(b), "l"(c));
    return d;
}
__device__ __forceinline__ void st_half(unsigned long long addr, float v) {
    const unsigned short h = __half_as_ushort(__float2half_rn(v));
    asm volatile("st.global.u16 [%0], %1;" ::"l"(addr), "h"(h) : "memory");
}

// one element off the fast path (a count outside the table, a logit below the fast logarithm's range, an edge tile): out of line
template <int SRC>
__device__ __noinline__ NbGrad nb_bwd_general(uint32_t code, float xp, float xs, float acc_pi, float4 gc, const uint8_t* tg_row,
                                              const void* X, long xidx, const float* dgt, float DpI, float DsI) {
    float2 tcn;
    if (code == NB_CODE_SLOW) tcn = nb_count_terms_bwd_slow(nb_load_raw<SRC>(X, xidx), gc.x, __ldg(dgt));
    else tcn = *reinterpret_cast<const float2*>(tg_row + code);
    const float pi = acc_pi + gc.w;
    if (tcn.x != 0.0f && fminf(xp, xs) < NB_X_RARE) return nb_backward_v5<true>(tcn.x, tcn.y, xp, xs, pi, gc.x, gc.y, gc.z, DpI, DsI);
    return nb_backward_v5<false>(tcn.x, tcn.y, xp, xs, pi, gc.x, gc.y, gc.z, DpI, DsI);
}

template <int SRC>
__global__ void __launch_bounds__(THREADS, 3) nb_tc_bwd_kernel(const __grid_constant__ CUtensorMap mapA,
                                                               const __grid_constant__ CUtensorMap mapB,
                                                               const __grid_constant__ CUtensorMap mapZ,
                                                               const __grid_constant__ CUtensorMap mapZc, NbTcBwdParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = tc::smem_u32(smem_raw);
    const uint32_t pad = (1024u - (raw & 1023u)) & 1023u;
    uint8_t* tiles = smem_raw + pad;
    uint8_t* z_tiles = tiles + STAGES * STAGE_BYTES;
    float4* s_gc = reinterpret_cast<float4*>(z_tiles + 2 * B_BYTES);  // [BN]: theta, theta + eps, K1c, bm
    float* s_col = reinterpret_cast<float*>(s_gc + BN);                 // [4 quantities][4 quarters][BN]
    uint8_t* s_tg = reinterpret_cast<uint8_t*>(s_col + 16 * BN);        // [BN][NB_TAB] float2: (log1p(c), digamma term) per gene
    uint64_t* full = reinterpret_cast<uint64_t*>(s_tg + BN * NB_TAB * 8);
    uint64_t* empty = full + STAGES;
    uint64_t* z_full = empty + STAGES;
    uint64_t* tmem_full = z_full + 1;
    uint64_t* tmem_ready = tmem_full + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_ready + 1);
    uint32_t* s_cnt = reinterpret_cast<uint32_t*>(tiles);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n0 = blockIdx.x * BN, m0 = blockIdx.y * BM;
    const int num_kb = (p.K + BK - 1) / BK;

    if (warp == 0 && lane == 0) {
        tc::tma_prefetch_desc(&mapA);
        tc::tma_prefetch_desc(&mapB);
        tc::tma_prefetch_desc(&mapZ);
        tc::tma_prefetch_desc(&mapZc);
        for (int s = 0; s < STAGES; ++s) {
            tc::mbar_init(&full[s], 1);
            tc::mbar_init(&empty[s], 1);
        }
        tc::mbar_init(z_full, 1);
        tc::mbar_init(tmem_full, 1);
        tc::mbar_init(tmem_ready, 1);
        tc::fence_barrier_init();
    }
    __syncthreads();
    // tensor memory is allocated by the MMA warp when it is about to issue, so that a third resident CTA runs its prologue
    // while two others own the SM's 512 columns (see nb_tc_fwd_kernel)
    uint32_t tmem_base = 0, tmem_z = 0;  // mixture-logit accumulator [BN columns]; the two branch-logit accumulators [2 BN]

    if (warp == 0) {
        if (tc::elect_one()) {
            tc::mbar_expect_tx(z_full, 2 * B_BYTES);
            tc::tma_load_2d(&mapZ, z_full, z_tiles, 0, n0);  // folded branch weights, fp16 [2 Gp, 64]
            tc::tma_load_2d(&mapZ, z_full, z_tiles + B_BYTES, 0, p.Gp + n0);
            for (int i = 0; i < num_kb; ++i) {
                const int s = i % STAGES;
                const uint32_t ph = (i / STAGES) & 1;
                tc::mbar_wait(&empty[s], ph ^ 1);
                uint8_t* a_dst = tiles + s * STAGE_BYTES;
                tc::mbar_expect_tx(&full[s], STAGE_BYTES);
                tc::tma_load_2d(&mapA, &full[s], a_dst, i * BK, m0);
                tc::tma_load_2d(&mapB, &full[s], a_dst + A_BYTES, i * BK, n0);
            }
            {  // the branch k-block: centred latents (fp16), A tile only, next slot of the ring
                const int s = num_kb % STAGES;
                const uint32_t ph = (num_kb / STAGES) & 1;
                tc::mbar_wait(&empty[s], ph ^ 1);
                tc::mbar_expect_tx(&full[s], A_BYTES);
                tc::tma_load_2d(&mapZc, &full[s], tiles + s * STAGE_BYTES, 0, m0);
            }
        }
    } else if (warp == 1) {
        // two-step allocation as in nb_tc_fwd_kernel: the mixture-logit accumulator is completed while the CTA still waits
        // for the 2 BN columns of the branch logits; one thread issues every MMA and commit
        // A CTA that has allocated without giving up its permit keeps the SM from launching further CTAs (measured: the second
        // and third CTA of an SM entered 5 and 9.5 us after the first), so the two-step path is only taken by a CTA that finds
        // two others resident - the SM is full then anyway; the first two allocate everything at once.
        unsigned int smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        int ahead = 0;
        if (lane == 0) ahead = atomicAdd(&g_resident_bwd[smid & 255], 1);
        const bool at_once = __shfl_sync(0xffffffffu, ahead, 0) < 2;
        tc::tmem_alloc_keep_permit(tmem_slot, ACC_COLS);
        if (at_once) tc::tmem_alloc(tmem_slot + 1, Z_COLS);
        tc::fence_before_sync();
        __syncwarp();
        tc::fence_after_sync();
        tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);
        constexpr uint32_t idesc = tc::idesc_f16(BM, BN);  // every operand of the fused decoder is fp16
        constexpr uint32_t idesc_z = tc::idesc_f16(BM, BN);
        const int s_z = num_kb % STAGES;  // ring slot of the branch k-block (the centred latents)
        if (lane == 0) {
            for (int i = 0; i < num_kb; ++i) {
                const int s = i % STAGES;
                const uint32_t ph = (i / STAGES) & 1;
                tc::mbar_wait(&full[s], ph);
                tc::fence_after_sync();
                const uint32_t a_base = tc::smem_u32(tiles + s * STAGE_BYTES);
                const uint32_t b_base = a_base + A_BYTES;
#pragma unroll
                for (int kk = 0; kk < BK / 16; ++kk)
                    tc::umma_bf16(tmem_base, tc::smem_desc(a_base + kk * 32, 16, 1024), tc::smem_desc(b_base + kk * 32, 16, 1024),
                                  idesc, (i > 0 || kk > 0) ? 1u : 0u);
                tc::umma_commit(&empty[s]);
            }
        }
        __syncwarp();
        if (!at_once) {
            tc::tmem_alloc(tmem_slot + 1, Z_COLS);  // whole warp; blocks while two other CTAs own their full sets
            tc::fence_before_sync();
            __syncwarp();
            tc::fence_after_sync();
        }
        tmem_z = *reinterpret_cast<volatile uint32_t*>(tmem_slot + 1);
        if (lane == 0) {
            tc::mbar_arrive(tmem_ready);  // release: the epilogue warps read the slots after acquiring this barrier
            tc::mbar_wait(z_full, 0);
            tc::mbar_wait(&full[s_z], (num_kb / STAGES) & 1);
            tc::fence_after_sync();
            const uint32_t a_base = tc::smem_u32(tiles + s_z * STAGE_BYTES);
            const uint32_t zp_base = tc::smem_u32(z_tiles), zs_base = zp_base + B_BYTES;
#pragma unroll
            for (int kk = 0; kk < BK / 16; ++kk) {
                tc::umma_bf16(tmem_z, tc::smem_desc(a_base + kk * 32, 16, 1024), tc::smem_desc(zp_base + kk * 32, 16, 1024), idesc_z,
                              kk > 0 ? 1u : 0u);
                tc::umma_bf16(tmem_z + BN, tc::smem_desc(a_base + kk * 32, 16, 1024), tc::smem_desc(zs_base + kk * 32, 16, 1024),
                              idesc_z, kk > 0 ? 1u : 0u);
            }
            tc::umma_commit(tmem_full);
        }
        __syncwarp();
    } else {
        // ================= epilogue: 8 warps =================
        const int et = threadIdx.x - 64;
        const long G = p.G;
        // Every global load of the prologue is issued before anything waits on one (as in nb_tc_fwd_kernel): the row indices
        // first, then the per-row and per-gene constants; the LUT is computed while they are in flight.
        const int e = warp - 2;
        const int q = warp & 3;
        const int half = e >> 2;
        const int rloc = q * 32 + lane;
        const int m = m0 + rloc;
        const bool mok = m < p.B;
        const int mm = mok ? m : 0;
        constexpr int NGATHER = BM / EPI_WARPS / GATHER_ROWS;
        int ridx[NGATHER];
#pragma unroll
        for (int i = 0; i < NGATHER; ++i) {
            const int gm = m0 + cnt_row(e, lane, i);
            ridx[i] = gm < p.B ? (p.rows ? __ldg(p.rows + gm) : gm) : -1;
        }
        const int my_row = p.rows ? __ldg(p.rows + mm) : mm;
        const float4 rc = __ldg(reinterpret_cast<const float4*>(p.rowc) + mm);  // Rp, Rs, Dp, Ds
        const float libm = __ldg(p.lib + mm);
        float4 gcv = make_float4(1.0f, 1.0f, 0.0f, 0.0f);
        static_assert(BN <= EPI_THREADS, "one thread per gene of the tile stages its constants");
        if (et < BN && n0 + et < p.G) {
            const int g = n0 + et;
            gcv.x = __ldg(p.genec + GC_THETA * G + g);
            gcv.y = __ldg(p.genec + GC_THE * G + g);
            gcv.z = __ldg(p.genec + GC_K1C * G + g);
            gcv.w = __ldg(p.bm + g);
        }
        float4 tgv[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int w16 = et + u * EPI_THREADS;
            const int g = n0 + (w16 >> 3);
            tgv[u] = g < p.G ? __ldg(reinterpret_cast<const float4*>(p.tgb + (long)n0 * NB_TAB) + w16) : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        }
        // coalesced row gather of the tile's counts into registers as count codes (overlaps the MMA phase)
        uint32_t cw[NGATHER];
        {
            const int g = n0 + 2 * (lane % (BN / 2));
#pragma unroll
            for (int i = 0; i < NGATHER; ++i) cw[i] = 0u;
            constexpr int BATCH = SRC == SPV_SRC_U16_LOG1P ? NGATHER : NGATHER / 2;  // every load of a batch in flight before any is used
            if (nb_pair_vec_ok<SRC>(p.X, p.ldx, g, p.G)) {
#pragma unroll
                for (int i0 = 0; i0 < NGATHER; i0 += BATCH) {
                    uint2 raw[BATCH];
#pragma unroll
                    for (int i = 0; i < BATCH; ++i) {
                        raw[i] = make_uint2(0u, 0u);
                        if (ridx[i0 + i] >= 0) raw[i] = nb_load_pair_vec<SRC>(p.X, (long)ridx[i0 + i] * p.ldx, g);
                    }
#pragma unroll
                    for (int i = 0; i < BATCH; ++i) cw[i0 + i] = nb_pair_codes<SRC>(raw[i]);
                }
            } else {  // odd pitch, unaligned base or the last gene of an odd-sized matrix: element loads
#pragma unroll
                for (int i = 0; i < NGATHER; ++i)
                    if (ridx[i] >= 0) cw[i] = nb_pair_codes<SRC>(nb_load_pair<SRC>(p.X, (long)ridx[i] * p.ldx, g, p.G));
            }
        }
        if (et < BN) s_gc[et] = gcv;
#pragma unroll
        for (int u = 0; u < 2; ++u) reinterpret_cast<float4*>(s_tg)[et + u * EPI_THREADS] = tgv[u];
        const float inv_elib = fast_exp(-libm);
        const float DpI = inv_elib * rc.z, DsI = inv_elib * rc.w;
        const long xrow = (long)my_row * p.ldx;
        tc::mbar_wait(tmem_ready, 0);
        tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);
        tmem_z = *reinterpret_cast<volatile uint32_t*>(tmem_slot + 1);
        tc::mbar_wait(tmem_full, 0);  // accumulators complete; the operand stages are free from here on
        tc::fence_after_sync();
#pragma unroll
        for (int i = 0; i < NGATHER; ++i) s_cnt[cnt_row(e, lane, i) * CNT_PITCH_W + lane % (BN / 2)] = cw[i];
        asm volatile("bar.sync 1, %0;" ::"r"(EPI_THREADS) : "memory");
        const uint32_t ldb = (uint32_t)p.ld_dpi * 2u;
        // running store addresses: this cell's column in the three blocks of D3T, at the first gene of the current four
        const unsigned long long blk_bytes = (unsigned long long)p.Gp * ldb;
        unsigned long long d_pi = reinterpret_cast<unsigned long long>(p.dpi) + (unsigned long long)(n0 + half * WCOLS) * ldb + 2ull * (unsigned)m;
        unsigned long long d_yp = d_pi + blk_bytes, d_ys = d_yp + blk_bytes;
        const bool full_tile = n0 + BN <= p.G && m0 + BM <= p.B;  // CTA-uniform
        const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16), lane_z = tmem_z + ((uint32_t)(q * 32) << 16);
        const uint32_t* cnt_row_p = s_cnt + rloc * CNT_PITCH_W + half * (WCOLS / 2);
#pragma unroll 1
        for (int j4 = 0; j4 < WCOLS; j4 += 4) {
            const int c0 = half * WCOLS + j4;
            uint32_t rpi[4], rlp[4], rls[4];
            tc::tmem_ld4(lane_addr + (uint32_t)c0, rpi);
            tc::tmem_ld4(lane_z + (uint32_t)c0, rlp);
            tc::tmem_ld4(lane_z + (uint32_t)(BN + c0), rls);
            const uint2 cc = *reinterpret_cast<const uint2*>(cnt_row_p + (j4 >> 1));  // four count codes
            tc::tmem_ld_wait();
            const uint32_t call = cc.x | cc.y;
            const float xm = fminf(fminf(fminf(__uint_as_float(rlp[0]), __uint_as_float(rls[0])), fminf(__uint_as_float(rlp[1]), __uint_as_float(rls[1]))),
                                   fminf(fminf(__uint_as_float(rlp[2]), __uint_as_float(rls[2])), fminf(__uint_as_float(rlp[3]), __uint_as_float(rls[3]))));
            // every count tabulated, every column and row valid, the fast logarithm holds (as in nb_tc_fwd_kernel)
            const bool plain = full_tile && (call & 0x80008000u) == 0u && !(call != 0u && xm < NB_X_RARE);
            float vyp[4], vys[4], vpi[4], vth[4];
            // D3T[gene, cell] (fp16, gradients of the log-likelihood: the consumers apply the signed scale): the 32 lanes of the
            // warp are 32 consecutive cells, so each of the three stores of a gene writes 64 contiguous bytes
            if (plain) {
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    const int gl = c0 + jj;
                    const uint32_t w = jj < 2 ? cc.x : cc.y;
                    const uint32_t code = (jj & 1) ? (w >> 16) : (w & 0xffffu);
                    const float2 tcn = *reinterpret_cast<const float2*>(s_tg + gl * (NB_TAB * 8) + code);
                    const float4 gc = s_gc[gl];
                    const NbGrad o = nb_backward_v5<false>(tcn.x, tcn.y, __uint_as_float(rlp[jj]), __uint_as_float(rls[jj]),
                                                           __uint_as_float(rpi[jj]) + gc.w, gc.x, gc.y, gc.z, DpI, DsI);
                    vyp[jj] = o.dyp; vys[jj] = o.dys; vpi[jj] = o.dpi; vth[jj] = o.dth;
                    st_half(d_pi + jj * ldb, o.dpi);
                    st_half(d_yp + jj * ldb, o.dyp);
                    st_half(d_ys + jj * ldb, o.dys);
                }
            } else {
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    const int gl = c0 + jj;
                    const uint32_t w = jj < 2 ? cc.x : cc.y;
                    const uint32_t code = (jj & 1) ? (w >> 16) : (w & 0xffffu);
                    vyp[jj] = vys[jj] = vpi[jj] = vth[jj] = 0.0f;
                    if (mok && n0 + gl < p.G) {
                        const NbGrad o = nb_bwd_general<SRC>(code, __uint_as_float(rlp[jj]), __uint_as_float(rls[jj]), __uint_as_float(rpi[jj]),
                                                             s_gc[gl], s_tg + gl * (NB_TAB * 8), p.X, xrow + n0 + gl,
                                                             p.genec + GC_DGT * G + n0 + gl, DpI, DsI);
                        vyp[jj] = o.dyp; vys[jj] = o.dys; vpi[jj] = o.dpi; vth[jj] = o.dth;
                        st_half(d_pi + jj * ldb, o.dpi);
                        st_half(d_yp + jj * ldb, o.dyp);
                        st_half(d_ys + jj * ldb, o.dys);
                    }
                }
            }
            d_pi += 4ull * ldb; d_yp += 4ull * ldb; d_ys += 4ull * ldb;
            // column sums over this warp's 32 rows (lanes): transpose-reduce of the 16 values (4 quantities x 4 columns).  Each
            // butterfly step halves the values a lane carries, 15 shuffles in all instead of 16 x 5; lanes with bit 0 clear
            // end up with the total of value index (lane >> 1) and park it for the cross-quarter sum.
            {
                float v16[16];
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) { v16[jj] = vyp[jj]; v16[4 + jj] = vys[jj]; v16[8 + jj] = vpi[jj]; v16[12 + jj] = vth[jj]; }
                float v8[8], v4[4], v2[2];
                const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4, b1 = lane & 2;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float send = b4 ? v16[i] : v16[i + 8], keep = b4 ? v16[i + 8] : v16[i];
                    v8[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float send = b3 ? v8[i] : v8[i + 4], keep = b3 ? v8[i + 4] : v8[i];
                    v4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
                }
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const float send = b2 ? v4[i] : v4[i + 2], keep = b2 ? v4[i + 2] : v4[i];
                    v2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
                }
                const float send = b1 ? v2[0] : v2[1], keep = b1 ? v2[1] : v2[0];
                float tot = keep + __shfl_xor_sync(0xffffffffu, send, 2);
                tot += __shfl_xor_sync(0xffffffffu, tot, 1);
                if ((lane & 1) == 0) {
                    const int idx = lane >> 1;  // = 8 b4 + 4 b3 + 2 b2 + b1: quantity idx >> 2, column idx & 3
                    s_col[((idx >> 2) * 4 + q) * BN + c0 + (idx & 3)] = tot;
                }
            }
        }
        asm volatile("bar.sync 1, %0;" ::"r"(EPI_THREADS) : "memory");
        for (int i = et; i < 4 * BN; i += EPI_THREADS) {
            const int qty = i / BN, gl = i - qty * BN, g = n0 + gl;
            if (g < p.G) {
                const float* s = s_col + (qty * 4) * BN + gl;
                p.colpart[((long)blockIdx.y * 4 + qty) * G + g] = s[0] + s[BN] + s[2 * BN] + s[3 * BN];
            }
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 1) {
        tc::fence_after_sync();
        tc::tmem_dealloc(tmem_z, Z_COLS);
        tc::tmem_dealloc(tmem_base, ACC_COLS);
        unsigned int smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        if (lane == 0) atomicSub(&g_resident_bwd[smid & 255], 1);
    }
}

__global__ void colpart_reduce_tc_kernel(const float* __restrict__ colpart, int nTB, int G, float* __restrict__ colsum, float mult) {
    long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (i >= 4L * G) return;
    float s = 0.0f;
    for (int t = 0; t < nTB; ++t) s += colpart[(long)t * 4 * G + i];
    colsum[i] = s * mult;
}

}  // namespace

// ptrs: the SPV_DEC_NPTR list (X, rows, -, -, -, bm, genec, lib, -, rowc, -, -, -, -, -, colpart [ceil(B/128), 4, G], -).
// d3_f16 [3 * Gp, ld_d3] (GENE-major, ld_d3 >= B a multiple of 8) receives d loss / d pi (rows 0 .. G), d loss / d y_private
// (rows Gp ..), d loss / d y_shared (rows 2 Gp ..), each / scale (i.e. the gradients of the log-likelihood), as FP16 (the A
// operand of the gradient GEMMs, which multiply by the signed scale: spv_tc_gemm_ex fmt 3); rows G .. Gp of each block and columns B .. ld_d3 are not written.  Operands as spv_dec_nb_fwd_tc.  colsum [4, G] =
// column sums of dyp, dys, dpi, dtheta (true scale).
extern "C" int spv_dec_nb_bwd_tc(int src, const void* const* ptrs, long long ldx, const void* amix_bf16, long long ld_amixb,
                                 const void* wstack_bf16, long long ld_w, int Gp, const void* zc_f16, const void* wz_f16,
                                 void* d3_f16, long long ld_d3, int B, int G, int HD, int P, int S, float scale, float* colsum, int kmix,
                                 void* stream) {
    if (!ptrs || !amix_bf16 || !wstack_bf16 || !zc_f16 || !wz_f16 || !d3_f16 || !colsum || Gp < G || B <= 0 || G <= 0 || P <= 0 || S <= 0 ||
        ld_d3 < B)
        return SPV_ERR_ARG;
    if (P + S > ZK_MAX_LATENT) return SPV_ERR_ARG;
    const int need[] = {0, 5, 6, 7, 9, 15, 18};
    for (int i : need)
        if (!ptrs[i]) return SPV_ERR_ARG;
    const int K = kmix > 0 ? kmix : HD + P + S;  // width of the mixing net's input ([hm | zz | covariates])
    if (reinterpret_cast<uintptr_t>(ptrs[9]) & 15) return SPV_ERR_ARG;  // rowc rows are read as one float4
    CUtensorMap ma, mb, mz, mzc;
    int rc = spv_make_tensor_map_bf16(&ma, amix_bf16, (unsigned long long)K, (unsigned long long)B, (unsigned long long)ld_amixb, 64, BM);
    if (rc != SPV_OK) return rc;
    rc = spv_make_tensor_map_bf16(&mb, wstack_bf16, (unsigned long long)K, (unsigned long long)G, (unsigned long long)ld_w, 64, BN);
    if (rc != SPV_OK) return rc;
    rc = spv_make_tensor_map_bf16(&mz, wz_f16, 64ull, (unsigned long long)(2 * Gp), 64ull, 64, BN);
    if (rc != SPV_OK) return rc;
    rc = spv_make_tensor_map_bf16(&mzc, zc_f16, 64ull, (unsigned long long)B, 64ull, 64, BM);
    if (rc != SPV_OK) return rc;
    NbTcBwdParams p;
    p.X = ptrs[0]; p.ldx = ldx; p.rows = (const int*)ptrs[1]; p.bm = (const float*)ptrs[5]; p.genec = (const float*)ptrs[6];
    p.lib = (const float*)ptrs[7]; p.rowc = (const float*)ptrs[9]; p.tgb = (const float2*)ptrs[18];
    p.colpart = (float*)ptrs[15]; p.dpi = reinterpret_cast<__nv_bfloat16*>(d3_f16); p.ld_dpi = ld_d3;
    // The sweep works in natural units: D3T holds the gradients of the LOG-LIKELIHOOD in fp16 - |d pi| <= 1, |d y| bounded by
    // log1p(count) + theta, all well inside fp16's normal range whatever the minibatch size - and the consumers apply the
    // signed scale: the column sums below, the gradient GEMMs through spv_tc_gemm_ex's alpha.
    if ((long long)ld_d3 >= (1ll << 28)) return SPV_ERR_ARG;  // the row pitch in bytes is a 32-bit quantity inside the kernel
    p.B = B; p.G = G; p.K = K; p.Gp = Gp; p.scale = scale;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    static bool configured[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!configured[dev & 63]) {
        if (cudaFuncSetAttribute(nb_tc_bwd_kernel<SPV_SRC_U16_LOG1P>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES) != cudaSuccess ||
            cudaFuncSetAttribute(nb_tc_bwd_kernel<SPV_SRC_F32_LOG1P>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES) != cudaSuccess)
            return SPV_ERR_LAUNCH;
        configured[dev & 63] = true;
    }
    dim3 grid((G + BN - 1) / BN, (B + BM - 1) / BM);
    if (src == SPV_SRC_U16_LOG1P) nb_tc_bwd_kernel<SPV_SRC_U16_LOG1P><<<grid, THREADS, SMEM_BYTES, st>>>(ma, mb, mz, mzc, p);
    else if (src == SPV_SRC_F32_LOG1P) nb_tc_bwd_kernel<SPV_SRC_F32_LOG1P><<<grid, THREADS, SMEM_BYTES, st>>>(ma, mb, mz, mzc, p);
    else return SPV_ERR_ARG;
    SPV_CHECK_LAUNCH();
    colpart_reduce_tc_kernel<<<(4 * G + 255) / 256, 256, 0, st>>>(p.colpart, (int)grid.y, G, colsum, scale);
    SPV_CHECK_LAUNCH();
    return SPV_OK;
}
